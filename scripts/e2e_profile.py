"""Where the time of the BASELINE config-2 solve (n = 1e6 SQP) goes on the GPU box."""
import cProfile
import pstats
import sys
import time

sys.path[:0] = ['.', 'ip-nonlinear-solver_b200', 'tests']
import numpy as np  # noqa: E402
import ipsolver_b200 as ipb  # noqa: E402
import e2e_problems  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "eqp_window_1000000"
api = dict(minimize_constrained=ipb.minimize_constrained, LinearConstraint=ipb.LinearConstraint,
           BoxConstraint=ipb.BoxConstraint, NonlinearConstraint=ipb.NonlinearConstraint)
prob = e2e_problems.LARGE_PROBLEMS.get(name) or e2e_problems.PROBLEMS[name]
run = prob(api)
t0 = time.perf_counter(); res = run(); t1 = time.perf_counter()
print("first run %.2f s (includes context creation, uploads): niter %d cg %d" % (t1 - t0, res.niter, res.cg_niter))
run = prob(api)
pr = cProfile.Profile()
t0 = time.perf_counter(); pr.enable(); res = run(); pr.disable(); t1 = time.perf_counter()
print("second run %.2f s: niter %d cg %d nfev %d status %d opt %.3e" % (
    t1 - t0, res.niter, res.cg_niter, res.nfev, res.status, res.optimality))
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
