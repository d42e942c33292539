"""Every end-to-end problem of tests/e2e_problems.py on the device against the
golden record of the reference: counts, optimality, x / v errors."""
import os
import sys

sys.path[:0] = ['.', 'ip-nonlinear-solver_b200', 'tests']
import numpy as np  # noqa: E402
import ipsolver_b200 as ipb  # noqa: E402
import e2e_problems  # noqa: E402

api = dict(minimize_constrained=ipb.minimize_constrained, LinearConstraint=ipb.LinearConstraint,
           BoxConstraint=ipb.BoxConstraint, NonlinearConstraint=ipb.NonlinearConstraint)
gold = np.load(os.path.join('tests', 'golden', 'e2e.npz'))
rel = lambda a, b: np.linalg.norm(np.asarray(a) - b) / max(np.linalg.norm(b), 1e-300)  # noqa: E731
for name, build in e2e_problems.PROBLEMS.items():
    niter, cg, status, nfev, opt, viol, fval, radius = gold[name + "/scalars"]
    res = build(api)()
    print("%-26s dev %4d/%5d/%4d st %d opt %.3e | ref %4d/%5d/%4d st %d opt %.3e | x %.1e v %.1e fun %.1e" % (
        name, res.niter, res.cg_niter, res.nfev, res.status, res.optimality, niter, cg, nfev, status, opt,
        rel(res.x, gold[name + "/x"]), rel(res.v, gold[name + "/v"]) if len(res.v) else 0.0,
        abs(res.fun - fval) / max(abs(fval), 1e-300)), flush=True)
