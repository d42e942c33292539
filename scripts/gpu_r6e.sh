#!/bin/bash
# r6g: ring consumers spread over the schedulers (IPB_SWEEP_SPREAD) -- sweep alone and the bench step
tag=${1:-r6g}
mkdir -p gpurun_out
V=$PWD/ip-nonlinear-solver_b200/ipsolver_b200/variants
{
for v in plain_sp rot_sp plain_sp_d15 rot_sp_d15 rot_sp_d8; do
echo "== $v"
IPB_LIB_PATH=$V/lib_$v.so timeout 120 python scripts/sweep_only.py 1.25e6 0 96 128
IPB_LIB_PATH=$V/lib_$v.so timeout 120 python scripts/sweep_only.py 1e7 0 264 352 528
done
} 2>&1 | grep -v Warning | tee gpurun_out/${tag}_diag.txt
for v in plain_sp rot_sp; do for cfg in "10000000 0" "10000000 352" "1250000 0"; do
set -- $cfg
IPB_LIB_PATH=$V/lib_$v.so python bench.py --n $1 --steps 20 --warmup 3 --no-cpu --solve-n 0 --solve-config3 none --solve-config4 none --band-chunk $2 > gpurun_out/${tag}_${v}_n$1_B$2.json 2> gpurun_out/${tag}_${v}_n$1_B$2.err
python -c "
import json; d=json.load(open('gpurun_out/${tag}_${v}_n$1_B$2.json')); r=d['roofline']; print('$v', $1, $2, 'it/s %.1f ms %.3f frac %.4f' % (d['value'], d['ms_per_step'], r['frac']), d['result_check']['x_checksum'], r['timeline']['per_iteration_us']['banded sweep'])"
done; done 2>&1 | tee gpurun_out/${tag}_bench.txt
