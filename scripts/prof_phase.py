"""Launch one phase-bench kernel (for ncu): python scripts/prof_phase.py <phase> <reps> [n] [pattern]"""
import sys
sys.path[:0] = ['.', 'ip-nonlinear-solver_b200']
from ipsolver_b200 import synthetic, device
from ipsolver_b200.projections import projections, set_inner_options
phase, reps = int(sys.argv[1]), int(sys.argv[2])
n = int(float(sys.argv[3])) if len(sys.argv) > 3 else 10_000_000
pattern = sys.argv[4] if len(sys.argv) > 4 else 'window'
H, A, c, b = synthetic.make_qp(n, pattern, 1e4, 0)
ctx = device.Context.default()
set_inner_options(aat_mode=2)
Z, LS, Y = projections(A)
proj = Z.projector
for _ in range(3):
    ms = device.bench_phase(proj, phase, reps)
print("phase", phase, "reps", reps, "ms", ms)
