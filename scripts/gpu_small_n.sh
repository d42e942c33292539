#!/bin/bash
# Bench step at the per-rank sizes of an 8/4/2-GPU strong-scaling run, on ONE GPU: what a rank's
# share of the n = 1e7 problem costs without any cross-GPU exchange.  Usage: scripts/gpu_small_n.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
for n in ${SIZES:-1250000 2500000 5000000}; do
  timeout 300 python bench.py --n $n --steps 20 --warmup 5 --no-cpu --solve-n 0 --solve-config3 none \
     --solve-config4 none > gpurun_out/${tag}_n${n}.json 2> gpurun_out/${tag}_n${n}.err
  python - <<P
import json
d=json.load(open("gpurun_out/${tag}_n${n}.json"))
r=d["roofline"]
print($n, "ms/step %.3f it/s %.0f frac %.3f" % (d["ms_per_step"], d["value"], r["frac"]), r.get("timeline",{}).get("per_iteration_us"), "sweep_us", r.get("sweep_us"), d["clocks"])
P
done
