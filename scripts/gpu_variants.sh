#!/bin/bash
# phases bench for each tuning variant: scripts/gpu_variants.sh <tag> <names...>
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  echo "=== variant $v" >> gpurun_out/${tag}_variants.txt
  IPB_LIB_PATH=$PWD/ip-nonlinear-solver_b200/ipsolver_b200/variants/lib_$v.so timeout 300 python scripts/bench_phases.py 10000000 window 2>&1 | grep -v "^{" >> gpurun_out/${tag}_variants.txt
done
cat gpurun_out/${tag}_variants.txt
