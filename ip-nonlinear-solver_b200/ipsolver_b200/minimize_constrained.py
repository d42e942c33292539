"""``minimize_constrained`` and the constraint classes -- the thin API shell
above the two outer loops, so that a user of the reference can run the B200
path end to end without the reference installed.

This is NOT a re-implementation of the reference's modelling layer
(``_constraints.py``, ``_canonical_constraint.py``, ``_numdiff.py`` are out of
scope): only what the hot path's callers need is here --
the public call signature (reference ``_minimize_constrained.py:96-100``), the
three constraint containers with the ``kind`` vocabulary (``("equals", c)``,
``("less", ub)``, ``("greater", lb)``, ``("interval", lb, ub)``), the split into
``c_ineq(x) <= 0`` / ``c_eq(x) = 0`` in the reference's row order (lower-bound
rows, then upper-bound rows, per constraint; ``_canonical_constraint.py:283-360``)
and the stopping rules (``_minimize_constrained.py:460-502``).  ``hess`` returns a
scipy sparse matrix or an ndarray (the Lagrangian Hessian is then applied on the
device) or a ``LinearOperator`` (an opaque host callback: the CG recurrence is
then driven from the host, projections stay on the device); the reference's
finite-difference modes (``hess='2-point'`` ..., ``_numdiff.py``) are not built.
"""
import time
import warnings

import numpy as np
import scipy.sparse as sps
from scipy.optimize import OptimizeResult
from scipy.sparse.linalg import LinearOperator

from .equality_constrained_sqp import equality_constrained_sqp, device_hot_path
from .tr_interior_point import tr_interior_point

__all__ = ['minimize_constrained', 'NonlinearConstraint', 'LinearConstraint', 'BoxConstraint']

MESSAGES = {
    0: "The maximum number of function evaluations is exceeded.",
    1: "`gtol` termination condition is satisfied.",
    2: "`xtol` termination condition is satisfied.",
    3: "`callback` function requested termination",
}


class NonlinearConstraint:
    """``fun(x)`` constrained by ``kind``; ``jac(x)`` its Jacobian (sparse or
    dense), ``hess(x, v)`` the Hessian of ``v . fun`` (matrix) or None."""

    def __init__(self, fun, kind, jac, hess=None, enforce_feasibility=False):
        self.fun, self.kind, self.jac, self.hess = fun, kind, jac, hess
        self.enforce_feasibility = enforce_feasibility


class LinearConstraint:
    """``A x`` constrained by ``kind``."""

    def __init__(self, A, kind, enforce_feasibility=False):
        self.A, self.kind, self.enforce_feasibility = A, kind, enforce_feasibility


class BoxConstraint:
    """``x`` itself constrained by ``kind``."""

    def __init__(self, kind, enforce_feasibility=False):
        self.kind, self.enforce_feasibility = kind, enforce_feasibility


def _bounds_of(kind, m):
    """``kind`` -> (lb, ub) arrays of length m."""
    if isinstance(kind, str):
        kind = (kind,)
    if not isinstance(kind, (tuple, list)) or len(kind) == 0:
        raise ValueError("The parameter `kind` should be a non-empty tuple, list or string.")
    word, args = kind[0], list(kind[1:])
    if word not in ("greater", "less", "equals", "interval"):
        raise ValueError("Keyword `%s` not available." % word)
    if len(args) != 2 if word == "interval" else len(args) > 1:
        raise ValueError("Invalid `kind` format.")
    if not args:
        args = [0]
    vals = [np.broadcast_to(np.asarray(a, dtype=float), (m,)).copy() for a in args]
    if word == "equals":
        return vals[0], vals[0].copy()
    if word == "greater":
        return vals[0], np.full(m, np.inf)
    if word == "less":
        return np.full(m, -np.inf), vals[0]
    if (vals[0] > vals[1]).any():
        raise ValueError("lb[i] > ub[i].")
    return vals[0], vals[1]


def _inside_box(box, x0, rel=0.01, absol=0.01):
    """Entries of x0 whose bounds are to be kept feasible throughout are moved a
    little inside them (by ``absol``, or ``rel`` of the interval if that is
    less) -- reference ``_reinforce_box_constraint`` (_constraints.py:449-477)."""
    m = x0.shape[0]
    enforce = np.broadcast_to(np.asarray(box.enforce_feasibility, dtype=bool), (m,))
    lb, ub = _bounds_of(box.kind, m)
    if not (enforce & ((x0 < lb) | (x0 > ub))).any():
        return x0
    with np.errstate(invalid='ignore'):
        low = np.minimum(lb + absol, lb + rel * (ub - lb))
        up = np.maximum(ub - absol, ub - rel * (ub - lb))
    x = x0.copy()
    sel = enforce & np.isfinite(lb)
    x[sel] = np.maximum(x[sel], low[sel])
    sel = enforce & np.isfinite(ub)
    x[sel] = np.minimum(x[sel], up[sel])
    return x


class _Rows:
    """One user constraint split into equality and inequality rows."""

    def __init__(self, fun, jac, hess, kind, enforce, x0, sparse_jacobian, constant_jac=False):
        # constant_jac: linear / box constraints -- the Jacobian never changes
        # (reference _canonical_constraint.py:251-265 re-assembles it every time)
        self.constant_jac = bool(constant_jac)
        f0 = np.atleast_1d(fun(x0))
        self.m = f0.shape[0]
        J0 = jac(x0)
        self.sparse = bool(sparse_jacobian or (sparse_jacobian is None and sps.issparse(J0)))
        self.fun, self._jac, self.hess = fun, jac, hess
        lb, ub = _bounds_of(kind, self.m)
        fixed = (lb == ub) & np.isfinite(lb)
        low, up = ~fixed & np.isfinite(lb), ~fixed & np.isfinite(ub)
        idx = np.arange(self.m)
        self.eq, self.val_eq = idx[fixed], lb[fixed]
        self.ineq = np.hstack((idx[low], idx[up]))
        self.val_ineq = np.hstack((lb[low], ub[up]))
        self.sign = np.hstack((-np.ones(low.sum()), np.ones(up.sum())))
        enforce = np.broadcast_to(np.asarray(enforce, dtype=bool), (self.m,))
        self.enforce = enforce[self.ineq]
        if (enforce & ((f0 < lb) | (f0 > ub))).any():
            raise ValueError("Unfeasible initial point. Either set ``enforce_feasibility=False`` "
                             "or choose a new feasible initial point ``x0``.")

    def jac(self, x):
        J = self._jac(x)
        if self.sparse:
            return sps.csr_matrix(J)
        return J.toarray() if sps.issparse(J) else np.atleast_2d(J)

    def values(self, x):
        c = np.atleast_1d(self.fun(x))
        return self.sign * (c[self.ineq] - self.val_ineq), c[self.eq] - self.val_eq

    def jacobians(self, x):
        J = self.jac(x)
        n = J.shape[1]
        if len(self.ineq) == 0 and len(self.eq) == self.m:
            # every row is an equality row, in order: no row selection needed
            return (sps.csr_matrix((0, n)) if self.sparse else np.empty((0, n))), J
        if self.sparse:
            J_eq = J[self.eq, :] if len(self.eq) else sps.csr_matrix((0, n))
            J_in = (sps.diags(self.sign) @ J[self.ineq, :]).tocsr() if len(self.ineq) \
                else sps.csr_matrix((0, n))
        else:
            J_eq = J[self.eq, :] if len(self.eq) else np.empty((0, n))
            J_in = J[self.ineq, :] * self.sign[:, None] if len(self.ineq) else np.empty((0, n))
        return J_in, J_eq

    def hessian(self, x, v_eq, v_ineq):
        v = np.zeros(self.m)
        if len(v_eq):
            v[self.eq] += v_eq
        if len(v_ineq):
            pos = self.sign == 1
            v[self.ineq[pos]] += v_ineq[pos]
            v[self.ineq[~pos]] -= v_ineq[~pos]
        return self.hess(x, v)


def _rows_of(constraint, x0, sparse_jacobian):
    if isinstance(constraint, NonlinearConstraint):
        return _Rows(constraint.fun, constraint.jac, constraint.hess, constraint.kind,
                     constraint.enforce_feasibility, x0, sparse_jacobian)
    if isinstance(constraint, LinearConstraint):
        A = constraint.A
        if sparse_jacobian or (sparse_jacobian is None and sps.issparse(A)):
            A = sps.csr_matrix(A)
        else:
            A = A.toarray() if sps.issparse(A) else np.atleast_2d(A)
        return _Rows(lambda x: A.dot(x), lambda x: A, None, constraint.kind,
                     constraint.enforce_feasibility, x0, sparse_jacobian, constant_jac=True)
    if isinstance(constraint, BoxConstraint):
        n = x0.shape[0]
        eye = sps.eye(n).tocsr() if (sparse_jacobian or sparse_jacobian is None) else np.eye(n)
        return _Rows(lambda x: x, lambda x: eye, None, constraint.kind,
                     constraint.enforce_feasibility, x0, sparse_jacobian, constant_jac=True)
    raise ValueError("Unknown Constraint type.")


class _Canonical:
    """All constraints stacked: ``constr(x) -> (c_ineq, c_eq)``, ``jac(x) ->
    (J_ineq, J_eq)``, ``hess_terms(x, v_eq, v_ineq) -> [matrices]``."""

    def __init__(self, blocks, n_vars, sparse_jacobian):
        self.blocks, self.n_vars = blocks, n_vars
        self.n_eq = sum(len(b.eq) for b in blocks)
        self.n_ineq = sum(len(b.ineq) for b in blocks)
        self.sparse = (any(b.sparse for b in blocks) if blocks
                       else bool(sparse_jacobian or sparse_jacobian is None))
        self.enforce_feasibility = (np.hstack([b.enforce for b in blocks]).astype(bool)
                                    if blocks else np.empty(0, dtype=bool))
        # all constraints linear: the stacked Jacobian is assembled once and the SAME
        # objects are handed out every time, so projections() recognises them and
        # neither uploads nor refreshes anything
        self._constant = bool(blocks) and all(b.constant_jac for b in blocks)
        self._jac_cache = None

    def constr(self, x):
        if not self.blocks:
            return np.empty(0), np.empty(0)
        parts = [b.values(x) for b in self.blocks]
        return np.hstack([p[0] for p in parts]), np.hstack([p[1] for p in parts])

    def jac(self, x):
        if self._constant and self._jac_cache is not None:
            return self._jac_cache
        out = self._assemble_jac(x)
        if self._constant:
            self._jac_cache = out
        return out

    def _assemble_jac(self, x):
        n = self.n_vars
        if not self.blocks:
            empty = sps.csr_matrix((0, n)) if self.sparse else np.empty((0, n))
            return empty, empty
        parts = [b.jacobians(x) for b in self.blocks]
        if self.sparse:
            return (sps.vstack([sps.csr_matrix(p[0]) for p in parts], format="csr"),
                    sps.vstack([sps.csr_matrix(p[1]) for p in parts], format="csr"))
        dense = lambda M: M.toarray() if sps.issparse(M) else np.atleast_2d(M)   # noqa: E731
        return np.vstack([dense(p[0]) for p in parts]), np.vstack([dense(p[1]) for p in parts])

    def hess_terms(self, x, v_eq, v_ineq):
        terms, ie, ii = [], 0, 0
        for b in self.blocks:
            ne, ni = len(b.eq), len(b.ineq)
            if b.hess is not None:
                terms.append(b.hessian(x, v_eq[ie:ie + ne], v_ineq[ii:ii + ni]))
            ie += ne
            ii += ni
        return terms


def minimize_constrained(fun, x0, grad, hess, constraints=(), method=None, xtol=1e-8,
                         gtol=1e-8, sparse_jacobian=None, options={}, callback=None,
                         max_iter=1000, verbose=0, hot_path=None):
    """Reference ``minimize_constrained`` (``_minimize_constrained.py:96``) for
    Hessians given as matrices.  ``hot_path`` supplies the L1 functions the
    loops call (default: the single-GPU device path;
    ``DistSolver.hot_path()`` for a row-partitioned run; the CPU tests pass
    the oracle's routines) -- it is not part of the reference's signature."""
    hp = hot_path or device_hot_path()
    x0 = np.atleast_1d(x0).astype(float)
    n_vars = x0.size
    if not callable(hess):
        raise TypeError("ipsolver_b200.minimize_constrained needs `hess` as a callable returning "
                        "a scipy sparse matrix or an ndarray; finite-difference Hessians are host "
                        "callbacks outside the device path.")
    # as in the reference, f0 and g0 are taken at the x0 the caller passed, BEFORE
    # a box with enforce_feasibility moves it (_minimize_constrained.py:378-379 vs
    # :431-432) -- a quirk that shapes the first iteration and is kept for parity
    f0 = fun(x0)
    g0 = np.atleast_1d(grad(x0))
    if isinstance(constraints, (NonlinearConstraint, LinearConstraint, BoxConstraint)):
        constraints = [constraints]
    for con in constraints:                      # reference _constraints.py:315-344
        if isinstance(con, BoxConstraint):
            x_in = _inside_box(con, x0)
            if not np.array_equal(x_in, x0):
                warnings.warn("The initial point was changed in order to stay inside box "
                              "constraints.")
                x0 = x_in
    canon = _Canonical([_rows_of(c, x0, sparse_jacobian) for c in constraints], n_vars,
                       sparse_jacobian)
    c_ineq0, c_eq0 = canon.constr(x0)
    J_ineq0, J_eq0 = canon.jac(x0)

    def lagr_hess(x, v_eq=np.empty(0), v_ineq=np.empty(0)):
        H0 = hess(x)
        terms = canon.hess_terms(x, v_eq, v_ineq)
        if isinstance(H0, LinearOperator) or any(isinstance(t, LinearOperator) for t in terms):
            # an opaque term (reference _canonical_constraint.py:131-135): the sum is a
            # host callback; projected_cg then drives its loop from the host and only
            # the projections run on the device
            parts = [H0] + terms
            return LinearOperator((len(x), len(x)), dtype=np.float64,
                                  matvec=lambda p: sum(np.asarray(t.dot(p)).reshape(-1)
                                                       for t in parts))
        H0 = sps.csr_matrix(H0) if sps.issparse(H0) else np.atleast_2d(np.asarray(H0))
        return hp.hessian(len(x), [H0] + terms)

    state = OptimizeResult(niter=0, nfev=1, ngev=1, ncev=1, njev=1, nhev=0, cg_niter=0, cg_info={})
    options = dict(options)
    if options.get("return_all", False):
        state.allvecs, state.allmult = [], []
    if method is None:
        method = 'equality_constrained_sqp' if canon.n_ineq == 0 else 'tr_interior_point'
    barrier_tol = options.pop("barrier_tol", 1e-8)

    def stop_criteria(state):
        state.status = None
        if callback is not None and callback(state):
            state.status = 3
        elif state.optimality < gtol and state.constr_violation < gtol:
            state.status = 1
        elif state.trust_radius < xtol and (method != 'tr_interior_point'
                                            or state.barrier_parameter < barrier_tol):
            state.status = 2
        elif state.niter > max_iter:
            state.status = 0
        return state.status in (0, 1, 2, 3)

    t0 = time.time()
    if method == 'equality_constrained_sqp':
        if canon.n_ineq > 0:
            raise ValueError("'equality_constrained_sqp' does not support inequality constraints.")
        result = equality_constrained_sqp(
            lambda x: (fun(x), canon.constr(x)[1]),
            lambda x: (np.atleast_1d(grad(x)), canon.jac(x)[1]),
            lagr_hess, x0, f0, g0, c_eq0, J_eq0, stop_criteria, state,
            hot_path=hp, **options)
    elif method == 'tr_interior_point':
        result = tr_interior_point(
            fun, lambda x: np.atleast_1d(grad(x)), lagr_hess, n_vars, canon.n_ineq, canon.n_eq,
            canon.constr, canon.jac, x0, f0, g0, c_ineq0, J_ineq0, c_eq0, J_eq0, stop_criteria,
            canon.enforce_feasibility, xtol, state, hot_path=hp, **options)
    else:
        raise ValueError("Unknown optimization ``method``.")
    result.execution_time = time.time() - t0
    result.method = method
    result.message = MESSAGES[result.status]
    if verbose >= 1:
        print(result.message)
        print("Number of iterations: {0}, function evaluations: {1}, CG iterations: {2}, "
              "optimality: {3:.2e}, constraint violation: {4:.2e}, execution time: {5:4.2} s."
              .format(result.niter, result.nfev, result.cg_niter, result.optimality,
                      result.constr_violation, result.execution_time))
    return result
