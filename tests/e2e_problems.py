"""End-to-end problems, written against an *API namespace* so the very same
problem can be handed to the reference (``ipsolver``, when generating golden
vectors in the build container) and to the B200 drop-in (``ipsolver_b200``).

``PROBLEMS[name](api)`` returns a zero-argument callable that runs
``api['minimize_constrained']`` and returns its ``OptimizeResult``.
"""
import os
import sys

import numpy as np
import scipy.sparse as sps

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(_ROOT, "ip-nonlinear-solver_b200"))
from ipsolver_b200 import synthetic  # noqa: E402


class ThomsonSphere:
    """N unit charges on the unit sphere, Coulomb energy (COPS problem 2).

    Variables are stacked ``[x; y; z]`` (3N).  The objective Hessian is a
    dense ``3N x 3N`` array and the constraint Hessian a dense block-diagonal,
    as in the reference's ``Elec`` test problem (reference
    ``ipsolver/tests/test_minimized_constrained.py:194-309``), which is what
    BASELINE.json configs 1 and 4 name.
    """

    def __init__(self, n_charges, seed=0):
        self.N = n_charges
        rng = np.random.RandomState(seed)
        phi = rng.uniform(0, 2 * np.pi, n_charges)
        theta = rng.uniform(-np.pi, np.pi, n_charges)
        self.x0 = np.hstack((np.cos(theta) * np.cos(phi),
                             np.cos(theta) * np.sin(phi), np.sin(theta)))

    def _deltas(self, p):
        q = p.reshape(3, self.N)
        return q[:, :, None] - q[:, None, :]          # (3, N, N)

    def fun(self, p):
        d = self._deltas(p)
        with np.errstate(divide='ignore'):
            inv = (d[0]**2 + d[1]**2 + d[2]**2) ** -0.5
        np.fill_diagonal(inv, 0.0)
        return 0.5 * np.sum(inv)

    def grad(self, p):
        d = self._deltas(p)
        with np.errstate(divide='ignore'):
            inv3 = (d[0]**2 + d[1]**2 + d[2]**2) ** -1.5
        np.fill_diagonal(inv3, 0.0)
        return np.hstack([-np.sum(d[k] * inv3, axis=1) for k in range(3)])

    def hess(self, p):
        d = self._deltas(p)
        dist = (d[0]**2 + d[1]**2 + d[2]**2) ** 0.5
        with np.errstate(divide='ignore'):
            inv3 = dist ** -3
            inv5 = dist ** -5
        np.fill_diagonal(inv3, 0.0)
        np.fill_diagonal(inv5, 0.0)
        blocks = [[None] * 3 for _ in range(3)]
        for a in range(3):
            for b in range(a, 3):
                blk = -3 * d[a] * d[b] * inv5
                if a == b:
                    blk = inv3 + blk
                np.fill_diagonal(blk, 0.0)
                np.fill_diagonal(blk, -np.sum(blk, axis=1))
                blocks[a][b] = blocks[b][a] = blk
        return np.block(blocks)

    def constraint(self, api, kind):
        N = self.N

        def c(p):
            q = p.reshape(3, N)
            return q[0]**2 + q[1]**2 + q[2]**2 - 1

        def jac(p):
            idx = np.arange(N)
            return sps.csr_matrix(
                (2 * p, (np.tile(idx, 3), np.arange(3 * N))), shape=(N, 3 * N))

        def hess(p, v):
            return np.diag(np.tile(2 * v, 3))

        return api['NonlinearConstraint'](c, (kind,), jac, hess)


def _eqp(n, pattern, kappa=10.0, **method_opts):
    def build(api):
        H, A, q, b = synthetic.make_qp(n, pattern, kappa, seed=0)

        def run():
            return api['minimize_constrained'](
                lambda x: 0.5 * x.dot(H.dot(x)) + q.dot(x), np.zeros(n),
                lambda x: H.dot(x) + q, lambda x: H,
                api['LinearConstraint'](A, ("equals", b)),
                method='equality_constrained_sqp',
                **method_opts)
        return run
    return build


def _ipqp(n, pattern="window", kappa=10.0):
    def build(api):
        H, A, q, b = synthetic.make_qp(n, pattern, kappa, seed=0)
        Ai = synthetic.sparse_jacobian(n, n // 4, pattern, seed=7)
        # pivots of the inequality rows on odd columns, clear of A's pivots
        Ai = sps.csr_matrix((Ai.data, (Ai.indices + 1) % n, Ai.indptr),
                            shape=Ai.shape)
        Ai.sort_indices()
        bi = 1.0 + np.random.RandomState(8).uniform(size=n // 4)

        def run():
            return api['minimize_constrained'](
                lambda x: 0.5 * x.dot(H.dot(x)) + q.dot(x), np.zeros(n),
                lambda x: H.dot(x) + q, lambda x: H,
                [api['LinearConstraint'](A, ("equals", b)),
                 api['LinearConstraint'](Ai, ("less", bi)),
                 api['BoxConstraint'](("interval", -10, 10))],
                method='tr_interior_point')
        return run
    return build


def _thomson(N, kind, method):
    def build(api):
        prob = ThomsonSphere(N)

        def run():
            return api['minimize_constrained'](
                prob.fun, prob.x0, prob.grad, prob.hess,
                prob.constraint(api, kind), method=method)
        return run
    return build


# ---- small classics (Nocedal & Wright): equality, inequality + bounds, no
# constraints at all (m = 0: Z = I), linear inequality, linear eq + ineq
def _maratos(api):
    """min 2(x^2 + y^2 - 1) - x  s.t.  x^2 + y^2 = 1 (ex. 15.4); x* = (1, 0)."""
    th = np.pi / 3
    x0 = np.array([np.cos(th), np.sin(th)])
    con = api['NonlinearConstraint'](
        lambda x: [x[0]**2 + x[1]**2], ("equals", 1),
        lambda x: [[2 * x[0], 2 * x[1]]], lambda x, v: 2 * v[0] * np.eye(2))
    return lambda: api['minimize_constrained'](
        lambda x: 2 * (x[0]**2 + x[1]**2 - 1) - x[0], x0,
        lambda x: np.array([4 * x[0] - 1, 4 * x[1]]), lambda x: 4 * np.eye(2), con)


def _hyperbolic(api):
    """min 1/2 (x-2)^2 + 1/2 (y-1/2)^2  s.t.  1/(x+1) - y >= 1/4, x, y >= 0
    (ex. 19.1); x* = (1.952823, 0.088659)."""
    def c_hess(x, v):
        return 2 * v[0] * np.array([[1 / (x[0] + 1)**3, 0], [0, 0]])
    con = [api['NonlinearConstraint'](lambda x: 1 / (x[0] + 1) - x[1], ("greater", 1 / 4),
                                      lambda x: [[-1 / (x[0] + 1)**2, -1]], c_hess),
           api['BoxConstraint'](("greater",))]
    return lambda: api['minimize_constrained'](
        lambda x: 1 / 2 * (x[0] - 2)**2 + 1 / 2 * (x[1] - 1 / 2)**2, np.array([0.0, 0.0]),
        lambda x: np.array([x[0] - 2, x[1] - 1 / 2]), lambda x: np.eye(2), con)


def _rosen_fun(x):
    return 100 * (x[1] - x[0]**2)**2 + (1 - x[0])**2


def _rosen_grad(x):
    return np.array([-400 * x[0] * (x[1] - x[0]**2) - 2 * (1 - x[0]), 200 * (x[1] - x[0]**2)])


def _rosen_hess(x):
    return np.array([[1200 * x[0]**2 - 400 * x[1] + 2, -400 * x[0]], [-400 * x[0], 200.0]])


def _rosenbrock(constraints):
    def build(api):
        x0 = np.array([-1.2, 1.0]) if constraints == "none" else np.array([-1.0, -0.5])
        if constraints == "none":
            con = ()
        elif constraints == "ineq":
            con = api['LinearConstraint'](np.array([[1.0, 2.0]]), ("less", 1))
        else:
            con = [api['LinearConstraint'](np.array([[1.0, 2.0]]), ("less", 1)),
                   api['LinearConstraint'](np.array([[2.0, 1.0]]), ("equals", 1))]
        return lambda: api['minimize_constrained'](_rosen_fun, x0, _rosen_grad, _rosen_hess, con)
    return build


def _rosenbrock_box_enforced(api):
    """Rosenbrock in a box that must stay feasible at every iterate
    (``enforce_feasibility``): the infeasible start is moved inside and the
    slack of those rows follows the constraint (N&W p. 576)."""
    con = api['BoxConstraint'](("interval", [-0.5, -0.5], [0.4, 0.3]), enforce_feasibility=True)
    return lambda: api['minimize_constrained'](_rosen_fun, np.array([-1.2, 1.0]), _rosen_grad,
                                               _rosen_hess, con)


PROBLEMS = {
    "rosenbrock_box_enforced": _rosenbrock_box_enforced,
    "maratos": _maratos,
    "hyperbolic_ineq": _hyperbolic,
    "rosenbrock_unconstrained": _rosenbrock("none"),
    "rosenbrock_ineq": _rosenbrock("ineq"),
    "rosenbrock_eq_ineq": _rosenbrock("eqineq"),
    "eqp_window_2000": _eqp(2000, "window"),
    "eqp_rand_2000": _eqp(2000, "rand"),
    "eqp_window_20000": _eqp(20000, "window"),
    "eqp_window_2000_gtol6": _eqp(2000, "window", gtol=1e-6),
    "eqp_rand_2000_gtol6": _eqp(2000, "rand", gtol=1e-6),
    "eqp_window_20000_gtol6": _eqp(20000, "window", gtol=1e-6),
    "ipqp_window_1000": _ipqp(1000),
    "ipqp_window_4000": _ipqp(4000),
    "thomson_20_equals": _thomson(20, "equals", "equality_constrained_sqp"),
    "thomson_20_less": _thomson(20, "less", "tr_interior_point"),
}

# BASELINE.json configs 2-4 at their stated sizes (golden records: a strided
# sample of x / v plus norms, tests/golden/make_golden_large.py -> e2e_large.npz)
LARGE_PROBLEMS = {
    "eqp_window_1000000": _eqp(1000000, "window"),
    "ipqp_window_1000000": _ipqp(1000000),
    "ipqp_window_100000": _ipqp(100000),
    "thomson_2000_equals": _thomson(2000, "equals", "equality_constrained_sqp"),
    "thomson_500_equals": _thomson(500, "equals", "equality_constrained_sqp"),
}
LARGE_STRIDE = 997       # sample x[::LARGE_STRIDE], v[::LARGE_STRIDE]
