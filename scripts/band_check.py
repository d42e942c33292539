"""Banded direct solve of (A A')^-1 on the GPU: accuracy against the CG path,
sweep timings for several chunk sizes, and a forced-iteration projected-CG run.
Usage: python scripts/band_check.py [n] [--sweep-sizes 64,128,256]"""
import json
import sys
import time

sys.path[:0] = ['.', 'ip-nonlinear-solver_b200']
import numpy as np  # noqa: E402
from ipsolver_b200 import synthetic, device  # noqa: E402
from ipsolver_b200.qp_subproblem import HessianOperator  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
sizes = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0").split(",")]
PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
m = n // 2
H, A, c, b = synthetic.make_qp(n, 'window', 1e4, 0)
ctx = device.Context.default()
dA = device.DeviceCSR(ctx, A, True)
proj = device.DeviceProjector(ctx, dA, aat_mode=2)
print("band info", proj.band_info, "nnzM", proj.aat_nnz, flush=True)
Hd = HessianOperator(n, [H]).device
out = {}
# reference result with CG only
proj.set_band(1)
t0 = time.perf_counter(); z_cg = proj.null_space(c); t_cg = time.perf_counter() - t0
i_cg = proj.last_info
v_cg = proj.least_squares(c); y_cg = proj.row_space(b)
print("CG    Z: inner %d solves %d refin %d orth %.2e kernel %.2f ms" % (
    i_cg.inner_iters, i_cg.n_solves, i_cg.n_refinements, i_cg.orth, i_cg.kernel_ms), flush=True)
dc, db, dx = device.DeviceVector(ctx, c), device.DeviceVector(ctx, b), device.DeviceVector(ctx, length=n)
for B in sizes:
    proj.set_band(0, B, 0)
    bi = proj.band_info
    z = proj.null_space(c); iz = proj.last_info
    v = proj.least_squares(c); il = proj.last_info
    y = proj.row_space(b); iy = proj.last_info
    rel = lambda a, r: float(np.linalg.norm(a - r) / np.linalg.norm(r))  # noqa: E731
    print("band B=%d %s" % (B, bi))
    print("  Z : inner %d refin %d fallbacks %d orth %.2e kernel %.3f ms  rel diff vs CG %.2e  ||A z||/||z|| %.2e" % (
        iz.inner_iters, iz.n_refinements, iz.n_fallbacks, iz.orth, iz.kernel_ms, rel(z, z_cg),
        np.linalg.norm(A.dot(z)) / np.linalg.norm(z)))
    print("  LS: inner %d fallbacks %d kernel %.3f ms rel diff %.2e" % (il.inner_iters, il.n_fallbacks, il.kernel_ms, rel(v, v_cg)))
    print("  Y : inner %d fallbacks %d kernel %.3f ms rel diff %.2e" % (iy.inner_iters, iy.n_fallbacks, iy.kernel_ms, rel(y, y_cg)))
    device.bench_phase(proj, 9, 3)
    reps = 20
    ms = min(device.bench_phase(proj, 9, reps) for _ in range(3)) / reps
    steps = bi["chunk_rows"] + 2 * bi["overlap"]
    moved = bi["groups"] * (2 * steps - bi["overlap"]) * 2048 + bi["groups"] * 32 * (steps - bi["overlap"]) * 16 + 8 * m * (steps / bi["chunk_rows"]) + 8 * m
    ideal = m * (2 * 64 + 32)
    print("  sweep %.1f us: moved %.3f GB -> %.0f GB/s (%.1f%% of peak); non-redundant bytes %.3f GB -> %.0f GB/s" % (
        1e3 * ms, moved / 1e9, moved / ms / 1e6, 100 * moved / ms / 1e6 / PEAK, ideal / 1e9, ideal / ms / 1e6), flush=True)
    for it in (8,):
        device.projected_cg_dev(proj, Hd, dc, db, dx, tol=1e-30, max_iter=it)
        info = device.projected_cg_dev(proj, Hd, dc, db, dx, tol=1e-30, max_iter=it)
        print("  projected_cg %d its: kernel %.3f ms -> %.1f it/s; inner %d solves %d refin %d max_orth %.2e" % (
            info.niter, info.kernel_ms, info.niter / info.kernel_ms * 1e3, info.inner_iters, info.n_solves,
            info.n_refinements, info.max_orth), flush=True)
        out["B%d" % B] = dict(sweep_us=1e3 * ms, pcg_ms=info.kernel_ms, it_s=info.niter / info.kernel_ms * 1e3)
    xb = dx.download()
proj.set_band(1)
device.projected_cg_dev(proj, Hd, dc, db, dx, tol=1e-30, max_iter=8)
info = device.projected_cg_dev(proj, Hd, dc, db, dx, tol=1e-30, max_iter=8)
xc = dx.download()
print("CG-only projected_cg: kernel %.3f ms inner %d; x rel diff band vs CG %.2e" % (
    info.kernel_ms, info.inner_iters, np.linalg.norm(xb - xc) / np.linalg.norm(xc)))
print(json.dumps(out))
# ---- phase timeline of one 8-iteration launch (banded path, default chunk size)
proj.set_band(0, int(sys.argv[3]) if len(sys.argv) > 3 else 0, 0)
device.projected_cg_dev(proj, Hd, dc, db, dx, tol=1e-30, max_iter=8)
ctx.trace(4000)
info = device.projected_cg_dev(proj, Hd, dc, db, dx, tol=1e-30, max_iter=8)
t = ctx.trace_read().astype(np.int64)
ctx.trace(0)
d = np.diff(t) / 1e3
print("timeline: %d barriers, kernel %.3f ms, sum of phases %.3f ms" % (len(t), info.kernel_ms, d.sum() / 1e3))
print("phases (us):", " ".join("%.0f" % v for v in d), "| fused H p:", int(info.reserved))
