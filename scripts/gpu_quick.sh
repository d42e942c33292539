#!/bin/bash
# Quick GPU pass after a kernel change: parity tests (no end-to-end runs), the phase microbench and the
# bench step at the headline size and at 1/8 of it.  Usage: scripts/gpu_quick.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${tag}_parity.txt
timeout 300 python scripts/bench_phases.py 10000000 window 20 2>&1 | grep -v "^{" | tail -11 | tee gpurun_out/${tag}_phases.txt
SIZES="10000000 1250000" scripts/gpu_small_n.sh $tag
