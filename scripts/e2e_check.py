"""End-to-end runs of ipsolver_b200.minimize_constrained on the GPU vs the golden results of the reference."""
import os, sys, time
sys.path[:0] = ['.', 'ip-nonlinear-solver_b200', 'tests']
import numpy as np
import e2e_problems
import ipsolver_b200 as ipb
gold = np.load('tests/golden/e2e.npz')
api = dict(minimize_constrained=ipb.minimize_constrained, LinearConstraint=ipb.LinearConstraint,
           BoxConstraint=ipb.BoxConstraint, NonlinearConstraint=ipb.NonlinearConstraint)
names = sys.argv[1:] or list(e2e_problems.PROBLEMS)
for name in names:
    g = gold[name + "/scalars"]
    t0 = time.time()
    try:
        res = e2e_problems.PROBLEMS[name](api)()
    except Exception as exc:
        print("%-26s FAILED %r" % (name, exc), flush=True)
        continue
    xg, vg = gold[name + "/x"], gold[name + "/v"]
    ex = np.linalg.norm(res.x - xg) / np.linalg.norm(xg)
    ev = np.linalg.norm(res.v - vg) / max(np.linalg.norm(vg), 1e-300)
    print("%-26s niter %d/%d cg %d/%d status %d/%d nfev %d/%d opt %.2e/%.2e viol %.1e fun rel %.1e  x rel %.1e  v rel %.1e  %.1fs"
          % (name, res.niter, g[0], res.cg_niter, g[1], res.status, g[2], res.nfev, g[3], res.optimality, g[4],
             res.constr_violation, abs(res.fun - g[6]) / abs(g[6]), ex, ev, time.time() - t0), flush=True)
