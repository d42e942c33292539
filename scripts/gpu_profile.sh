#!/bin/bash
# ncu launch list of the bench command + one --set full capture of the solve kernel.
tag=${1:-run}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${tag}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_projected_cg -s 1 -c 1 \
   -f -o gpurun_out/${tag}_pcg python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${tag}_pcg.log 2>&1
tail -2 gpurun_out/${tag}_pcg.log
