#!/bin/bash
# r6a: rotated factor-row registers in the banded sweep -- sweep alone at the headline size and at the
# per-rank sizes for several chunk lengths, then the bench step per chunk length (x checksum at the
# default length = bit-identity check against the r3y record).
tag=${1:-r6a}
mkdir -p gpurun_out
{
python scripts/sweep_only.py 1e7 0 264 352 528
python scripts/sweep_only.py 1.25e6 0 48 96 128
python scripts/sweep_only.py 2.5e6 0 128 176
} 2>&1 | grep -v Warning | tee gpurun_out/${tag}_sweeps.txt
for cfg in "10000000 0" "10000000 352" "10000000 528" "1250000 0" "1250000 96"; do
set -- $cfg
python bench.py --n $1 --steps 20 --warmup 3 --no-cpu --solve-n 0 --solve-config3 none --solve-config4 none --band-chunk $2 > gpurun_out/${tag}_n$1_B$2.json 2> gpurun_out/${tag}_n$1_B$2.err
python -c "
import json; d=json.load(open('gpurun_out/${tag}_n$1_B$2.json')); r=d['roofline']; print($1, $2, 'it/s %.1f ms %.3f frac %.4f' % (d['value'], d['ms_per_step'], r['frac']), 'e2e %.1f' % d['e2e']['value'], d['result_check'], r['timeline']['per_iteration_us'], 'refin', r['refinements_per_launch'], 'fallbacks', r['cg_fallbacks_per_launch'], r['band'])"
done 2>&1 | tee gpurun_out/${tag}_bench.txt
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "band and not medium" 2>&1 | tail -3 | tee gpurun_out/${tag}_parity.txt
