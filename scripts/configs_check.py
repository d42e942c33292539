"""BASELINE.json configs 3 and 4 on the GPU (scaled to what the host-side callbacks allow in a
minute or two): the barrier path (tr_interior_point on a sparse QP with linear inequalities and a
box) and Thomson/Elec with a dense Hessian block (GEMV-heavy projected_cg)."""
import json
import sys
import time
sys.path[:0] = ['.', 'ip-nonlinear-solver_b200', 'tests']
import numpy as np
import e2e_problems
import ipsolver_b200 as ipb
from ipsolver_b200.device import Context

api = dict(minimize_constrained=ipb.minimize_constrained, LinearConstraint=ipb.LinearConstraint,
           BoxConstraint=ipb.BoxConstraint, NonlinearConstraint=ipb.NonlinearConstraint)
runs = [("config 3: tr_interior_point, sparse QP n=%d + linear ineq + box", e2e_problems._ipqp, int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000),
        ("config 4: Thomson/Elec N=%d equality constraints, dense Hessian, SQP",
         lambda N: e2e_problems._thomson(N, "equals", "equality_constrained_sqp"), int(sys.argv[2]) if len(sys.argv) > 2 else 1000)]
for label, build, size in runs:
    ctx = Context.default()
    l0 = ctx.launches
    t0 = time.time()
    res = build(size)(api)()
    dt = time.time() - t0
    print(json.dumps({"run": label % size, "seconds": round(dt, 2), "niter": int(res.niter),
                      "cg_niter": int(res.cg_niter), "nfev": int(res.nfev), "status": int(res.status),
                      "optimality": float(res.optimality), "constr_violation": float(res.constr_violation),
                      "fun": float(res.fun), "kernel_launches": int(ctx.launches - l0)}), flush=True)
