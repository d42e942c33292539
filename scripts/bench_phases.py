"""Phase-level timings of the inner iteration on the GPU (measurement aid).

Each phase is repeated inside ONE persistent launch with a grid barrier after
every repetition (``ipb_bench_phase``), so the figure is what the phase costs
inside the solve kernels.  Bytes are the compulsory traffic of the phase.
Usage: python scripts/bench_phases.py [n] [pattern] [reps]
"""
import json
import sys

sys.path[:0] = ['.', 'ip-nonlinear-solver_b200']
import numpy as np  # noqa: E402
from ipsolver_b200 import synthetic, device  # noqa: E402
from ipsolver_b200.projections import projections, set_inner_options  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
pattern = sys.argv[2] if len(sys.argv) > 2 else 'window'
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
H, A, c, b = synthetic.make_qp(n, pattern, 1e4, 0)
m = n // 2
ctx = device.Context.default()
set_inner_options(aat_mode=2)
Z, LS, Y = projections(A)
proj = Z.projector
nnzM = proj.aat_nnz
B = lambda nnz, r, cc: 12 * nnz + 4 * (r + 1) + 8 * cc + 8 * r  # noqa: E731
phases = [
    ("update (9 m-vector passes)", 0, 72 * m),
    ("t = A' r", 1, B(A.nnz, n, m)),
    ("w = A t, w.r", 2, B(A.nnz, m, n) + 8 * m),
    ("w = M r, w.r", 3, B(nnzM, m, m)),
    ("barrier only", 4, 0),
    ("M r copy pipeline only", 5, 12 * nnzM + 4 * (m + 1) + 8 * m),
    ("M r row arithmetic only", 8, 0),
]
import torch  # noqa: E402
_a = torch.empty(1 << 27, dtype=torch.float64, device='cuda')
_b = torch.empty_like(_a)
_e0, _e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_best = 1e9
for _ in range(6):
    _e0.record(); _b.copy_(_a); _e1.record(); torch.cuda.synchronize()
    _best = min(_best, _e0.elapsed_time(_e1))
print("copy bandwidth of this box, this process: %.0f GB/s (2 x 1 GiB / %.3f ms)" % (2 * _a.numel() * 8 / _best / 1e6, _best))
del _a, _b
import subprocess  # noqa: E402
print(subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu', '--format=csv,noheader'], capture_output=True, text=True).stdout.strip())
print("n=%d m=%d nnzA=%d nnzM=%d grid=%s peak=%.0f GB/s" % (n, m, A.nnz, nnzM, ctx.grid(), PEAK))
out = {}
for name, ph, nbytes in phases:
    device.bench_phase(proj, ph, 5)
    ms = min(device.bench_phase(proj, ph, reps) for _ in range(4)) / reps     # best of 4
    gbs = nbytes / ms / 1e6
    out[name] = dict(us=1e3 * ms, GBs=gbs, frac=gbs / PEAK)
    print("%-28s %9.2f us  %8.1f GB/s  %5.1f%% of measured peak" % (name, 1e3 * ms, gbs, 100 * gbs / PEAK))
dc = device.DeviceVector(ctx, c)
for mode, label in ((1, "inner iteration, A then A'"), (2, "inner iteration, explicit M")):
    proj.set_aat_mode(mode)
    device.bench_inner_iterations(proj, dc, 20)
    it = 60
    ms = min(device.bench_inner_iterations(proj, dc, it) for _ in range(3)) / (it + 1)
    nb = 80 * m + (B(nnzM, m, m) if mode == 2 else B(A.nnz, n, m) + B(A.nnz, m, n))
    print("%-28s %9.2f us  %8.1f GB/s  %5.1f%%" % (label, 1e3 * ms, nb / ms / 1e6, 100 * nb / ms / 1e6 / PEAK))
    out[label] = dict(us=1e3 * ms, GBs=nb / ms / 1e6)
dH = device.DeviceCSR(ctx, H, False)
dn1, dn2 = device.DeviceVector(ctx, c), device.DeviceVector(ctx, length=n)
f = lambda: ctx.lib.ipb_csr_matvec_dev(ctx.h, dH.h, dn1.h, dn2.h)  # noqa: E731
f(); f()
ctx.timer_start()
for _ in range(20):
    f()
ms = ctx.timer_stop() / 20
print("%-28s %9.2f us  %8.1f GB/s  %5.1f%%  (separate launches)" % ("H p", 1e3 * ms, B(H.nnz, n, n) / ms / 1e6, 100 * B(H.nnz, n, n) / ms / 1e6 / PEAK))
print(json.dumps(out))
