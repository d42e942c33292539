#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list.  Usage: scripts/gpu_check.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"
tail -3 gpurun_out/${tag}_pytest.log
cat gpurun_out/${tag}_bench.json
tail -5 gpurun_out/${tag}_bench.err
