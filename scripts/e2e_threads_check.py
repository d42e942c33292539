"""Config 2 (n = 1e6 SQP) on one GPU under the host BLAS thread count given by
OMP_NUM_THREADS: the run ends by xtol, i.e. its tail is decided by round-off in
the HOST scalars (numpy dot / norm), whose summation order depends on the
thread count.  Usage: OMP_NUM_THREADS=k python scripts/e2e_threads_check.py"""
import os
import sys
sys.path[:0] = ['.', 'ip-nonlinear-solver_b200', 'tests']
import numpy as np  # noqa: E402
import e2e_problems  # noqa: E402
import ipsolver_b200 as ipb  # noqa: E402
from threadpoolctl import threadpool_info  # noqa: E402

api = dict(minimize_constrained=ipb.minimize_constrained, LinearConstraint=ipb.LinearConstraint,
           BoxConstraint=ipb.BoxConstraint, NonlinearConstraint=ipb.NonlinearConstraint)
res = e2e_problems.LARGE_PROBLEMS["eqp_window_1000000"](api)()
print("OMP_NUM_THREADS=%s blas threads %s cpus %d: niter %d cg %d nfev %d status %d opt %.6e fun %.12e"
      % (os.environ.get("OMP_NUM_THREADS"), [p.get("num_threads") for p in threadpool_info()],
         os.cpu_count(), res.niter, res.cg_niter, res.nfev, res.status, res.optimality, res.fun))
