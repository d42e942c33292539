"""Generate the golden vectors under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (``/root/reference`` does not exist on the
GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports the unmodified reference (``ipsolver``) from ``/root/reference``,
runs its ``projections`` / ``projected_cg`` / ``modified_dogleg`` /
intersection routines and ``minimize_constrained`` on the inputs defined in
``tests/cases.py`` and ``ipsolver_b200.synthetic`` and stores inputs-free
outputs (the inputs are regenerated deterministically by the tests).

Files written:
  kat.json            small known-answer cases (full-precision floats)
  synth_pcg.npz       synthetic n=2000 projector + projected-CG outputs
  e2e.npz             end-to-end ``minimize_constrained`` results
"""
import json
import os
import sys
import warnings

import numpy as np
import scipy.sparse as sps

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "ip-nonlinear-solver_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.dont_write_bytecode = True

from ipsolver._large_scale_constrained.projections import (  # noqa: E402
    projections, orthogonality)
from ipsolver._large_scale_constrained.qp_subproblem import (  # noqa: E402
    projected_cg, modified_dogleg, sphere_intersections, box_intersections,
    box_sphere_intersections)
from ipsolver import (minimize_constrained, LinearConstraint,  # noqa: E402
                      BoxConstraint, NonlinearConstraint)

import cases  # noqa: E402
from ipsolver_b200 import synthetic  # noqa: E402
import e2e_problems  # noqa: E402

warnings.simplefilter("ignore")


def f(x):
    return float(x)


def vec(x):
    return [float(v) for v in np.asarray(x, dtype=float).ravel()]


def kat():
    out = {"pcg": {}, "dogleg": {}, "proj38": {}, "ball": [], "box": [],
           "box_ball": []}
    for name, (H, A, c, b, kw, exc) in cases.PCG_CASES.items():
        Hs, As = sps.csc_matrix(H), sps.csc_matrix(A)
        Z, _, Y = projections(As)
        c, b = np.array(c), np.array(b)
        if exc is not None:
            try:
                projected_cg(Hs, c, Z, Y, b, **kw)
                raise AssertionError(name + " should raise")
            except exc as e:
                out["pcg"][name] = {"raises": type(e).__name__, "msg": str(e)}
            continue
        x, info = projected_cg(Hs, c, Z, Y, b, **kw)
        rec = {"x": vec(x), "niter": int(info["niter"]),
               "stop_cond": int(info["stop_cond"]),
               "hits_boundary": bool(info["hits_boundary"])}
        if "allvecs" in info:
            rec["allvecs"] = [vec(v) for v in info["allvecs"]]
        out["pcg"][name] = rec
    for name, (A, b, radius, lb, ub) in cases.DOGLEG_CASES.items():
        A = np.array(A, dtype=float)
        _, _, Y = projections(A)
        x = modified_dogleg(A, Y, np.array(b), radius, np.array(lb, float),
                            np.array(ub, float))
        out["dogleg"][name] = vec(x)
    A = sps.csc_matrix(np.array(cases.A38, dtype=float))
    Z, LS, Y = projections(A)
    out["proj38"]["Z"] = [vec(Z.matvec(np.array(p, float))) for p in cases.P38_N]
    out["proj38"]["LS"] = [vec(LS.matvec(np.array(p, float))) for p in cases.P38_N]
    out["proj38"]["Y"] = [vec(Y.matvec(np.array(p, float))) for p in cases.P38_M]
    out["proj38"]["orth"] = [f(orthogonality(A, np.array(p, float)))
                             for p in cases.P38_N]
    for z, d, radius, line in cases.BALL_CASES:
        ta, tb, hit = sphere_intersections(np.array(z, float),
                                           np.array(d, float), radius, line)
        out["ball"].append([f(ta), f(tb), bool(hit)])
    for z, d, lb, ub, line in cases.BOX_CASES:
        args = [np.array(v, float) for v in (z, d, lb, ub)]
        ta, tb, hit = box_intersections(*args, entire_line=line)
        out["box"].append([f(ta), f(tb), bool(hit)])
        for radius in (2.5, np.inf):
            ta, tb, hit = box_sphere_intersections(*args, radius, line)
            out["box_ball"].append([f(ta), f(tb), bool(hit)])
    with open(os.path.join(HERE, "kat.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print("kat.json:", {k: len(v) for k, v in out.items()})


def synth():
    out = {}
    n = cases.SYNTH_N
    for name, (pattern, kappa, kw, tau) in cases.SYNTH_RUNS.items():
        H, A, c, b = synthetic.make_qp(n, pattern, kappa, seed=0)
        Z, LS, Y = projections(A)
        kw = dict(kw)
        if tau is not None:
            kw["lb"], kw["ub"] = synthetic.box_bounds(n, tau)
        x, info = projected_cg(H, c, Z, Y, b, **kw)
        out[name + "/x"] = x
        out[name + "/info"] = np.array([info["niter"], info["stop_cond"],
                                        int(info["hits_boundary"])])
        print(name, info, "||x||=%.4f ||Y b||=%.4f" % (
            np.linalg.norm(x), np.linalg.norm(Y.dot(-b))))
    for pattern in ("window", "rand"):
        H, A, c, b = synthetic.make_qp(n, pattern, 1e4, seed=0)
        Z, LS, Y = projections(A)
        out[pattern + "/Zc"] = Z.dot(c)
        out[pattern + "/LSc"] = LS.dot(c)
        out[pattern + "/Yb"] = Y.dot(b)
        out[pattern + "/orth_Zc"] = np.array(orthogonality(A, Z.dot(c)))
        lb, ub = synthetic.box_bounds(n, 0.3)
        for radius in (4.0, 40.0):
            out[pattern + "/dogleg_r%g" % radius] = modified_dogleg(
                A, Y, b, radius, lb, ub)
        out[pattern + "/dogleg_free"] = modified_dogleg(
            A, Y, b, 6.0, np.full(n, -np.inf), np.full(n, np.inf))
    np.savez_compressed(os.path.join(HERE, "synth_pcg.npz"), **out)
    print("synth_pcg.npz:", len(out), "arrays")


def e2e(only=None):
    path = os.path.join(HERE, "e2e.npz")
    out = dict(np.load(path)) if (only and os.path.exists(path)) else {}
    for name, build in e2e_problems.PROBLEMS.items():
        if only and name not in only:
            continue
        prob = build(dict(minimize_constrained=minimize_constrained,
                          LinearConstraint=LinearConstraint,
                          BoxConstraint=BoxConstraint,
                          NonlinearConstraint=NonlinearConstraint))
        res = prob()
        out[name + "/x"] = np.asarray(res.x)
        out[name + "/v"] = np.asarray(res.v)
        out[name + "/scalars"] = np.array(
            [res.niter, res.cg_niter, res.status, res.nfev, res.optimality,
             res.constr_violation, res.fun, res.trust_radius], dtype=float)
        print(name, "niter", res.niter, "cg", res.cg_niter, "status",
              res.status, "nfev", res.nfev, "opt %.3e" % res.optimality,
              "viol %.3e" % res.constr_violation, "fun %.10f" % res.fun,
              "t %.2fs" % res.execution_time)
    np.savez_compressed(os.path.join(HERE, "e2e.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["kat", "synth", "e2e"]
    if which[0] == "e2e" and len(which) > 1:
        e2e(only=which[1:])          # add / refresh the named problems only
    else:
        for w in which:
            {"kat": kat, "synth": synth, "e2e": e2e}[w]()
