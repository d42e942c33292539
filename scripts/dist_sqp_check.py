"""Row-partitioned SQP check, run under torchrun (one rank per GPU):
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
       scripts/dist_sqp_check.py [n] [pattern]
Every rank runs ``minimize_constrained(method='equality_constrained_sqp')`` with
``DistSolver.hot_path()``: the host loop is replicated, each device step works
on the rank's own blocks (dogleg, projected CG, H d, A d, LS c across the
GPUs).  Rank 0 also solves the problem on its own GPU alone and compares; at
n = 1e6 (BASELINE config 2) the result is also held against the record of the
unmodified reference (tests/golden/e2e_large.npz: 27 / 80 / 40)."""
import os
import sys
import time

# torchrun exports OMP_NUM_THREADS=1; the SQP run ends by xtol, where the HOST scalars (numpy dot /
# norm) decide the tail, and their summation order depends on the BLAS thread count: use the count
# the reference record was made with (8), before numpy loads its BLAS
os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = os.environ.get("IPB_HOST_THREADS", "8")

sys.path[:0] = ['.', 'ip-nonlinear-solver_b200', 'tests']
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import ipsolver_b200 as ipb  # noqa: E402
from ipsolver_b200 import synthetic, distributed as D  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
size = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gloo = dist.new_group(backend="gloo")
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
pattern = sys.argv[2] if len(sys.argv) > 2 else 'window'
H, A, q, b = synthetic.make_qp(n, pattern, 10.0, seed=0)
t0 = time.perf_counter()
S = D.DistSolver(H, A, rank, size, group=gloo, device=local)
t_setup = time.perf_counter() - t0
args = (lambda x: 0.5 * x.dot(H.dot(x)) + q.dot(x), np.zeros(n), lambda x: H.dot(x) + q,
        lambda x: H)
rel = lambda a_, b_: np.linalg.norm(a_ - b_) / max(np.linalg.norm(b_), 1e-300)  # noqa: E731
ok = True

# ---- the helpers of the second-order correction: Y.dot and box_intersections across the ranks
hp = S.hot_path(H, A)
Z, LS, Y = hp.projections(A)
rng = np.random.default_rng(5)
bt = rng.standard_normal(n // 2)
y = Y.dot(bt)
d0 = rng.standard_normal(n)
lb, ub = synthetic.box_bounds(n, 1.5)
ta, tb, hit = hp.box_intersections(0.1 * d0, y, lb, ub)
if rank == 0:
    r_y = np.linalg.norm(A.dot(y) - bt) / np.linalg.norm(bt)
    print("Y b across %d ranks: ||A y - b||/||b|| = %.2e" % (size, r_y), flush=True)
    from ipsolver_b200.qp_subproblem import box_intersections
    ta1, tb1, hit1 = box_intersections(0.1 * d0, y, lb, ub)
    good = r_y < 1e-10 and hit == hit1 and abs(ta - ta1) <= 1e-12 * max(1, abs(ta1)) \
        and abs(tb - tb1) <= 1e-12 * max(1, abs(tb1))
    print("box_intersections dist (%.15g, %.15g, %s) | single (%.15g, %.15g, %s) %s"
          % (ta, tb, hit, ta1, tb1, hit1, "OK" if good else "MISMATCH"), flush=True)
    ok = ok and good

# ---- the SQP run
dist.barrier()
t0 = time.perf_counter()
res = ipb.minimize_constrained(*args, ipb.LinearConstraint(A, ("equals", b)),
                               method='equality_constrained_sqp', hot_path=hp)
t_dist = time.perf_counter() - t0
dist.barrier()
if rank == 0:
    print("setup %.2f s; %d ranks: niter %d cg %d nfev %d status %d opt %.6e fun %.12e viol %.2e in %.2f s"
          % (t_setup, size, res.niter, res.cg_niter, res.nfev, res.status, res.optimality, res.fun,
             res.constr_violation, t_dist), flush=True)
    t0 = time.perf_counter()
    one = ipb.minimize_constrained(*args, ipb.LinearConstraint(A, ("equals", b)),
                                   method='equality_constrained_sqp')
    t_one = time.perf_counter() - t0
    # the ranks' block-wise reductions round differently from one GPU's (1e-16): north_star's
    # +-1 on the outer counts; an extra outer iteration brings its own CG iterations along
    dn_it = abs(res.niter - one.niter)
    same = (dn_it <= 1 and abs(res.nfev - one.nfev) <= 1 and res.status == one.status
            and abs(res.cg_niter - one.cg_niter) <= 1 + 20 * dn_it)
    ex, ev = rel(res.x, one.x), rel(res.v, one.v)
    good = same and ex < 1e-8 and ev < 1e-8 and abs(res.fun - one.fun) <= 1e-10 * abs(one.fun)
    print("single GPU: niter %d cg %d nfev %d status %d opt %.6e fun %.12e in %.2f s | x rel %.2e v rel %.2e %s"
          % (one.niter, one.cg_niter, one.nfev, one.status, one.optimality, one.fun, t_one, ex, ev,
             "OK" if good else "MISMATCH"), flush=True)
    ok = ok and good
    if n == 1_000_000 and pattern == 'window':
        import e2e_problems
        g = np.load(os.path.join('tests', 'golden', 'e2e_large.npz'))
        name = "eqp_window_1000000"
        niter, cg, status, nfev, opt, viol, fval, radius, t_ref = g[name + "/scalars"]
        s = e2e_problems.LARGE_STRIDE
        ex = rel(res.x[::s], g[name + "/xs"])
        good = (abs(res.niter - niter) <= 1 and abs(res.cg_niter - cg) <= 1 and abs(res.nfev - nfev) <= 1
                and res.status == status and abs(res.fun - fval) <= 1e-8 * abs(fval) and ex < 1e-8)
        print("reference record: niter %d cg %d nfev %d status %d opt %.6e fun %.12e (%.0f s) | x sample rel %.2e %s"
              % (niter, cg, nfev, status, opt, fval, t_ref, ex, "OK" if good else "MISMATCH"), flush=True)
        ok = ok and good
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
