"""The UNMODIFIED reference (baseline/_ref) on BASELINE config 2 (n = 1e6) under
the host BLAS thread count given by OMP_NUM_THREADS -- the CPU counterpart of
scripts/e2e_threads_check.py.  Usage: OMP_NUM_THREADS=k python scripts/ref_threads_check.py [n]"""
import os
import sys
import time
sys.path[:0] = ['.', 'ip-nonlinear-solver_b200', 'tests']
import ref_harness  # noqa: E402
ipsolver = ref_harness.import_reference()
import e2e_problems  # noqa: E402
from threadpoolctl import threadpool_info  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
api = dict(minimize_constrained=ipsolver.minimize_constrained,
           LinearConstraint=ipsolver.LinearConstraint, BoxConstraint=ipsolver.BoxConstraint,
           NonlinearConstraint=ipsolver.NonlinearConstraint)
t0 = time.time()
res = e2e_problems._eqp(n, "window")(api)()
print("reference n=%d OMP_NUM_THREADS=%s blas threads %s cpus %d: niter %d cg %d nfev %d status %d opt %.6e fun %.12e in %.0f s"
      % (n, os.environ.get("OMP_NUM_THREADS"), [p.get("num_threads") for p in threadpool_info()],
         os.cpu_count(), res.niter, res.cg_niter, res.nfev, res.status, res.optimality, res.fun,
         time.time() - t0), flush=True)
