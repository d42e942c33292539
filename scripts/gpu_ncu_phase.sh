#!/bin/bash
# ncu --set full of one phase-bench launch: scripts/gpu_ncu_phase.sh <tag> <phase> [variant]
tag=$1; phase=$2; v=$3
mkdir -p gpurun_out
[ -n "$v" ] && export IPB_LIB_PATH=$PWD/ip-nonlinear-solver_b200/ipsolver_b200/variants/lib_$v.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_projector -s 2 -c 1 \
   -f -o gpurun_out/${tag}_phase${phase} python scripts/prof_phase.py $phase 6 > gpurun_out/${tag}_ncu_phase${phase}.log 2>&1
tail -3 gpurun_out/${tag}_ncu_phase${phase}.log
ls -la gpurun_out/
