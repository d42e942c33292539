"""Golden records of the BASELINE.json configs at their stated sizes, FROM THE
REFERENCE ITSELF (build container only; ``/root/reference`` does not travel):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_large.py NAME [NAME ...]

For every ``e2e_problems.LARGE_PROBLEMS[NAME]`` it runs the unmodified
reference ``minimize_constrained`` and merges a compact record into
``tests/golden/e2e_large.npz``:

  NAME/scalars  niter, cg_niter, status, nfev, optimality, constr_violation,
                fun, trust_radius, execution_time [s]
  NAME/xs, NAME/vs      x[::997], v[::997]
  NAME/norms    ||x||_2, ||x||_inf, ||v||_2, ||v||_inf, n, len(v)

(the full vectors are 8 MB+ each; the strided sample plus the norms pin the
solution without putting megabytes of doubles into the history).
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "ip-nonlinear-solver_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.dont_write_bytecode = True

from ipsolver import (minimize_constrained, LinearConstraint,  # noqa: E402
                      BoxConstraint, NonlinearConstraint)
import e2e_problems  # noqa: E402

warnings.simplefilter("ignore")


def record(res):
    x, v = np.asarray(res.x, float), np.asarray(res.v, float)
    s = e2e_problems.LARGE_STRIDE
    return {
        "scalars": np.array([res.niter, res.cg_niter, res.status, res.nfev, res.optimality,
                             res.constr_violation, res.fun, res.trust_radius,
                             res.execution_time], dtype=float),
        "xs": x[::s].copy(), "vs": v[::s].copy(),
        "norms": np.array([np.linalg.norm(x), np.abs(x).max(), np.linalg.norm(v),
                           np.abs(v).max() if v.size else 0.0, x.size, v.size]),
    }


def main(names):
    path = os.environ.get("GOLDEN_OUT") or os.path.join(HERE, "e2e_large.npz")
    api = dict(minimize_constrained=minimize_constrained, LinearConstraint=LinearConstraint,
               BoxConstraint=BoxConstraint, NonlinearConstraint=NonlinearConstraint)
    for name in names:
        res = e2e_problems.LARGE_PROBLEMS[name](api)()
        rec = record(res)
        print(name, "niter", res.niter, "cg", res.cg_niter, "status", res.status, "nfev", res.nfev,
              "opt %.6e" % res.optimality, "viol %.3e" % res.constr_violation,
              "fun %.12e" % res.fun, "t %.1fs" % res.execution_time, flush=True)
        # merge under a lock-free read-modify-write (one writer at a time is the rule)
        out = dict(np.load(path)) if os.path.exists(path) else {}
        for k, a in rec.items():
            out[name + "/" + k] = a
        np.savez_compressed(path, **out)


if __name__ == "__main__":
    main(sys.argv[1:])
