#!/bin/bash
# Sweep the rows-per-chunk of the banded solve: scripts/gpu_chunk_sweep.sh "<n list>" "<B list>"
for n in ${1:-1250000}; do for B in ${2:-64}; do
python bench.py --n $n --steps 10 --warmup 3 --no-cpu --solve-n 0 --solve-config3 none --solve-config4 none --band-chunk $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print($n, $B, 'ms/step %.3f'%d['ms_per_step'], 'sweep %.1f us'%r['timeline']['per_iteration_us']['banded sweep'], 'refinements', r['refinements_per_launch'], 'fallbacks', r['cg_fallbacks_per_launch'], r['band']['groups'], 'feas %.1e' % d['result_check']['rel_feasibility'])"
done; done
