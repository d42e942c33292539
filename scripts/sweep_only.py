"""Time of the banded sweep alone (ipb_bench_phase 9: repeated inside one launch, a grid barrier
after each).  Usage: python scripts/sweep_only.py n [chunk_rows ...]"""
import sys
sys.path[:0] = ['.', 'ip-nonlinear-solver_b200']
from ipsolver_b200 import synthetic, device  # noqa: E402
from ipsolver_b200.projections import projections, set_inner_options  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
chunks = [int(v) for v in sys.argv[2:]] or [0]
H, A, c, b = synthetic.make_qp(n, 'window', 1e4, 0)
for B in chunks:
    set_inner_options(band=(0, B, 0))
    Z, LS, Y = projections(A)
    proj = Z.projector
    device.bench_phase(proj, 9, 3)
    us = min(device.bench_phase(proj, 9, 20) for _ in range(3)) / 20 * 1e3
    info = proj.band_info
    steps = 2 * info["chunk_rows"] + 3 * info["overlap"]
    print("n=%d m=%d chunk_rows=%d groups=%d: sweep %.1f us (%.0f ns per step of a chunk)"
          % (n, n // 2, info["chunk_rows"], info["groups"], us, 1e3 * us / steps), flush=True)
