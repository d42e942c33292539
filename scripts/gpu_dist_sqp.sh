#!/bin/bash
# Row-partitioned checks on N GPUs of one box.  Usage: scripts/gpu_dist_sqp.sh <tag> <nranks> [n]
tag=${1:-run}; np=${2:-2}; n=${3:-1000000}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29711 scripts/dist_check.py 200000 window > gpurun_out/${tag}_dist_check.txt 2>&1; echo "dist_check exit $?" | tee -a gpurun_out/${tag}_dist_check.txt
grep -E "OK|MISMATCH|FAILED|rel err" gpurun_out/${tag}_dist_check.txt | tail -8
run 29712 scripts/dist_sqp_check.py $n window > gpurun_out/${tag}_dist_sqp.txt 2>&1; echo "dist_sqp exit $?" | tee -a gpurun_out/${tag}_dist_sqp.txt
grep -vE "^\[|Warning|warn|W1" gpurun_out/${tag}_dist_sqp.txt | tail -15
