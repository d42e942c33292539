"""Multi-GPU parity check, run under torchrun (one rank per GPU):
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/dist_check.py [n] [pattern]
Every rank solves its block of the row-partitioned problem; rank 0 also solves
the whole problem on its own GPU (single-GPU path) and compares."""
import os
import sys
import time

sys.path[:0] = ['.', 'ip-nonlinear-solver_b200']
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from ipsolver_b200 import synthetic, device, distributed as D  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
size = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gloo = dist.new_group(backend="gloo")
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000
pattern = sys.argv[2] if len(sys.argv) > 2 else 'window'
if pattern == 'hashed':
    # every rank generates only a window around its own rows; rank 0 alone also
    # builds the whole problem, for the single-GPU comparison
    H, A, c, b = synthetic.make_qp_hashed(n, 1e4, 0)
    t0 = time.perf_counter()
    S = D.DistSolver(None, None, rank, size, group=gloo, device=local,
                     generated=dict(n=n, kappa=1e4, seed=0))
else:
    H, A, c, b = synthetic.make_qp(n, pattern, 1e4, 0)
    t0 = time.perf_counter()
    S = D.DistSolver(H, A, rank, size, group=gloo, device=local)
lbg, ubg = synthetic.box_bounds(n, 1.5)
if rank == 0:
    print("setup %.2f s; ghosts n-space %s m-space %s" % (
        time.perf_counter() - t0, [len(g) for g in S.parts["ghosts"][0]],
        [len(g) for g in S.parts["ghosts"][1]]), flush=True)
dc, db = S.vector(S.local(c, 0)), S.vector(S.local(b, 1), space=1)
dx = S.vector(space=0)
dlb, dub = S.vector(S.local(lbg, 0)), S.vector(S.local(ubg, 0))
# ---- stage 1: products with halo exchange, and the three projector maps
import ctypes as C  # noqa: E402
lib, ctx = S.ctx.lib, S.ctx
dy_m, dy_n = S.vector(space=1), S.vector(space=0)
ctx.check(lib.ipb_csr_matvec_dev(ctx.h, S.A_raw.h, dc.h, dy_m.h))
Ax = S.gather(dy_m.download())
ctx.check(lib.ipb_csr_matvec_dev(ctx.h, S.At_raw.h, db.h, dy_n.h))
Atb = S.gather(dy_n.download())
ctx.check(lib.ipb_hess_matvec_dev(ctx.h, S.H.h, dc.h, dy_n.h))
Hc = S.gather(dy_n.download())
pinfo = device.ProjInfo()
ctx.check(lib.ipb_proj_null_space_dev(ctx.h, S.proj.h, dc.h, dy_n.h, C.byref(pinfo)))
Zc = S.gather(dy_n.download())
ctx.check(lib.ipb_proj_row_space_dev(ctx.h, S.proj.h, db.h, dy_n.h, C.byref(pinfo)))
Yb = S.gather(dy_n.download())
if rank == 0:
    rel = lambda a_, b_: np.linalg.norm(a_ - b_) / max(np.linalg.norm(b_), 1e-300)  # noqa: E731
    print("A x   rel err %.2e" % rel(Ax, A.dot(c)))
    print("A' b  rel err %.2e" % rel(Atb, A.T.dot(b)))
    print("H c   rel err %.2e" % rel(Hc, H.dot(c)))
    print("Z c : ||A Zc||/||Zc|| %.2e   inner its %d" % (np.linalg.norm(A.dot(Zc)) / np.linalg.norm(Zc), pinfo.inner_iters))
    print("Y b : ||A Yb - b||/||b|| %.2e" % (np.linalg.norm(A.dot(Yb) - b) / np.linalg.norm(b)), flush=True)
cases = [dict(tol=1e-30, max_iter=6), dict(), dict(trust_radius=20.0 * np.sqrt(n / 4000.0), lb=True)]
results = []
for kw in cases:
    kw = dict(kw)
    if kw.pop("lb", False):
        kw.update(lb=dlb, ub=dub)
    dist.barrier()
    try:
        info = S.projected_cg(dc, db, dx, **kw)
    except Exception as exc:
        i = S.proj.last_pcg_info
        print("rank", rank, "case", list(kw), "FAILED:", exc, "| niter", i.niter, "stop", i.stop_cond,
              "status", i.status, "inner", i.inner_iters, "solves", i.n_solves, "refin", i.n_refinements,
              "nH", i.n_spmv_h, "max_orth", i.max_orth, "tol", i.tol_used, flush=True)
        info = i
    x = S.gather(dx.download())
    results.append((x, int(info.niter), int(info.stop_cond), int(info.hits_boundary),
                    int(info.inner_iters), float(info.kernel_ms)))
dist.barrier()
ok = True
if rank == 0:
    from ipsolver_b200.projections import projections
    from ipsolver_b200.qp_subproblem import projected_cg, HessianOperator
    Z, LS, Y = projections(A)
    Hop = HessianOperator(n, [H])
    for kw, (x, niter, stop, hits, inner, ms) in zip(cases, results):
        kw = dict(kw)
        if kw.pop("lb", False):
            kw.update(lb=lbg, ub=ubg)
        xs, inf = projected_cg(Hop, c, Z, Y, b, **kw)
        err = np.linalg.norm(x - xs) / np.linalg.norm(xs)
        feas = np.linalg.norm(A.dot(x) + b) / np.linalg.norm(b)
        good = (abs(niter - inf["niter"]) <= 1 and stop == inf["stop_cond"]
                and bool(hits) == inf["hits_boundary"] and err < 1e-8)
        ok = ok and good
        print("case %s: dist niter %d stop %d hits %d inner %d (%.2f ms) | single %s | rel err %.2e feas %.1e %s"
              % (kw.keys(), niter, stop, hits, inner, ms, inf, err, feas, "OK" if good else "MISMATCH"),
              flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
