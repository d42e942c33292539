#!/bin/bash
tag=${1:-run}
mkdir -p gpurun_out
timeout 600 python scripts/bench_phases.py 10000000 window > gpurun_out/${tag}_phases.txt 2>&1
cat gpurun_out/${tag}_phases.txt
