#!/bin/bash
# multi-GPU parity check: scripts/gpu_dist.sh <tag> <nranks> [n] [pattern]
tag=$1; np=$2; n=${3:-200000}; pat=${4:-window}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 \
   scripts/dist_check.py $n $pat > gpurun_out/${tag}_dist${np}.log 2>&1
echo "exit $?" >> gpurun_out/${tag}_dist${np}.log
grep -v "^\s*$" gpurun_out/${tag}_dist${np}.log | grep -E "Error|error|case|setup|exit|Traceback|File|ipsolver|rel err|Zc|Yb|FAILED" | head -40
