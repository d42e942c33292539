#!/bin/bash
# multi-GPU: parity check + bench line.  scripts/gpu_dist_bench.sh <tag> <nranks>
tag=$1; np=$2
[ -z "$SKIP_CHECK" ] && bash scripts/gpu_dist.sh $tag $np 400000 window
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus $np --steps 3 --warmup 3 > gpurun_out/${tag}_bench${np}.json 2> gpurun_out/${tag}_bench${np}.err
echo "bench exit $?"; cat gpurun_out/${tag}_bench${np}.json | cut -c1-700; tail -5 gpurun_out/${tag}_bench${np}.err
