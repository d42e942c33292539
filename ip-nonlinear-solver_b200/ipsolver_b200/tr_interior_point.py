"""Trust-region barrier interior-point loop -- caller of the device hot path.

Host control logic with the interface of the reference's
``ipsolver/_large_scale_constrained/tr_interior_point.py`` (``BarrierSubproblem``
:24-251, ``tr_interior_point`` :254-355), re-stated from Byrd-Hribar-Nocedal
(1999) with the reference's constants.  What it hands to the hot path:

  * the barrier Jacobian ``[[J_eq, 0], [J_ineq, diag(s)]]`` as ONE CSR matrix
    whose pattern never changes -- a new iterate only changes values, so the
    device copy is refreshed in place (``projections`` caches the pattern);
  * the scaled Lagrangian Hessian ``blockdiag(Hx, S Hs S)`` as a
    :class:`HessianOperator` carrying the terms of ``Hx`` and the diagonal tail,
    so ``H.p`` runs as SpMV/GEMV + a diagonal block on the device (the reference
    builds an opaque closure, :222-241);
  * finite lower bounds ``-0.995`` on the slack block of the step (:306-308),
    which make the box logic of ``projected_cg`` live.
"""
import numpy as np
import scipy.sparse as sps
from scipy.sparse.linalg import LinearOperator

from .equality_constrained_sqp import equality_constrained_sqp

__all__ = ['tr_interior_point', 'BarrierSubproblem']

TAU_BOUNDARY = 0.995         # fraction-to-the-boundary, (3.18) of BHN
THETA_DECAY = 0.2            # decay of barrier parameter and subproblem tolerance
RADIUS_RESTART = 5           # trust-radius enlargement between barrier problems


class BarrierSubproblem:
    """``min fun(x) - mu*sum(log s)  s.t.  c_eq(x) = 0, c_ineq(x) + s = 0`` in
    the variables ``z = [x; s]``."""

    def __init__(self, x0, s0, fun, grad, lagr_hess, n_vars, n_ineq, n_eq, constr, jac,
                 barrier_parameter, tolerance, enforce_feasibility, global_stop_criteria,
                 xtol, fun0, grad0, constr_ineq0, jac_ineq0, constr_eq0, jac_eq0):
        self.n_vars, self.n_ineq, self.n_eq = n_vars, n_ineq, n_eq
        self.x0, self.s0 = x0, s0
        self.fun, self.grad, self.lagr_hess, self.constr, self.jac = fun, grad, lagr_hess, constr, jac
        self.barrier_parameter, self.tolerance = barrier_parameter, tolerance
        self.enforce_feasibility = enforce_feasibility
        self.global_stop_criteria, self.xtol = global_stop_criteria, xtol
        self._pattern = None
        self.fun0 = self._barrier_value(fun0, constr_ineq0, s0)
        self.grad0 = self._barrier_gradient(grad0)
        self.constr0 = np.hstack((constr_eq0, constr_ineq0 + s0))
        self.jac0 = self._barrier_jacobian(jac_eq0, jac_ineq0, s0)

    def update(self, barrier_parameter, tolerance):
        self.barrier_parameter, self.tolerance = barrier_parameter, tolerance

    def get_variables(self, z):
        return z[:self.n_vars]

    def get_slack(self, z):
        return z[self.n_vars:self.n_vars + self.n_ineq]

    # ---- value, gradient, constraints, Jacobian of the barrier problem
    def _barrier_value(self, f, c_ineq, s):
        # rows with enforce_feasibility keep s = -c_ineq (N&W p. 576); done in
        # place on the caller's z, as the reference does (:92)
        s[self.enforce_feasibility] = -c_ineq[self.enforce_feasibility]
        if s.shape[0] <= 4096:
            # scalar logs, then np.sum: the arithmetic of the reference (:93-95),
            # kept for small problems where runs are compared count for count
            logs = [np.log(si) if si > 0 else -np.inf for si in s]
        else:
            # the per-element Python loop is O(n_ineq) interpreter work per function
            # evaluation; vectorised, equal up to the last ulp of log
            with np.errstate(divide='ignore', invalid='ignore'):
                logs = np.where(s > 0, np.log(np.where(s > 0, s, 1.0)), -np.inf)
        return f - self.barrier_parameter * np.sum(logs)

    def _barrier_gradient(self, g):
        return np.hstack((g, -self.barrier_parameter * np.ones(self.n_ineq)))

    def function_and_constraints(self, z):
        x, s = self.get_variables(z), self.get_slack(z)
        f = self.fun(x)
        c_ineq, c_eq = self.constr(x)
        return self._barrier_value(f, c_ineq, s), np.hstack((c_eq, c_ineq + s))

    def gradient_and_jacobian(self, z):
        x, s = self.get_variables(z), self.get_slack(z)
        g = self.grad(x)
        J_ineq, J_eq = self.jac(x)
        return self._barrier_gradient(g), self._barrier_jacobian(J_eq, J_ineq, s)

    def _barrier_jacobian(self, J_eq, J_ineq, s):
        if self.n_ineq == 0:
            return J_eq
        if not (sps.issparse(J_eq) or sps.issparse(J_ineq)):
            J_eq, J_ineq = np.atleast_2d(J_eq), np.atleast_2d(J_ineq)
            top = np.hstack((J_eq, np.zeros((self.n_eq, self.n_ineq))))
            return np.vstack((top, np.hstack((J_ineq, np.diag(s)))))
        return self._sparse_jacobian(sps.csr_matrix(J_eq), sps.csr_matrix(J_ineq), s)

    def _sparse_jacobian(self, J_eq, J_ineq, s):
        """CSR of ``[[J_eq, 0], [J_ineq, diag(s)]]``: the rows of ``J_ineq`` each
        get one more entry (the slack, last in its row since its column index is
        the largest).  Index arrays are built once per pattern."""
        stacked = sps.vstack([J_eq, J_ineq], "csr")
        key = (stacked.nnz, stacked.shape)
        if self._pattern is None or self._pattern[0] != key \
                or not np.array_equal(self._pattern[1], stacked.indices) \
                or not np.array_equal(self._pattern[5], stacked.indptr):
            extra = np.concatenate((np.zeros(self.n_eq, dtype=np.int64),
                                    np.arange(self.n_ineq + 1, dtype=np.int64)))
            indptr = stacked.indptr.astype(np.int64) + extra
            nnz = stacked.nnz + self.n_ineq
            slack_pos = indptr[-self.n_ineq:] - 1
            is_slack = np.zeros(nnz, dtype=bool)
            is_slack[slack_pos] = True
            indices = np.empty(nnz, dtype=np.int64)
            indices[is_slack] = self.n_vars + np.arange(self.n_ineq)
            indices[~is_slack] = stacked.indices
            self._pattern = (key, stacked.indices.copy(), indptr, indices, is_slack,
                             stacked.indptr.copy())
        _, _, indptr, indices, is_slack, _ = self._pattern
        data = np.empty(indices.shape[0])
        data[is_slack] = s
        data[~is_slack] = stacked.data
        return sps.csr_matrix((data, indices, indptr),
                              (self.n_eq + self.n_ineq, self.n_vars + self.n_ineq))

    # ---- scaling and Hessian
    def scaling(self, z):
        diag = np.hstack((np.ones(self.n_vars), self.get_slack(z)))
        n = self.n_vars + self.n_ineq
        return LinearOperator((n, n), matvec=lambda vec: diag * vec, dtype=np.float64)

    def slack_hessian(self, z, v):
        """``S Hs S``: primal-dual form ``v*s`` where the multiplier is positive,
        primal form ``mu`` elsewhere (reference :206-220)."""
        s = self.get_slack(z)
        v_ineq = v[-self.n_ineq:]
        return np.where(v_ineq > 0, v_ineq * s, self.barrier_parameter)

    def lagrangian_hessian(self, z, v):
        x = self.get_variables(z)
        Hx = self.lagr_hess(x, v[:self.n_eq], v[self.n_eq:self.n_eq + self.n_ineq])
        if self.n_ineq == 0:
            return Hx
        tail = self.slack_hessian(z, v)
        if hasattr(Hx, "with_diag_tail"):        # HessianOperator: stays on the device
            return Hx.with_diag_tail(tail)
        # opaque operator (host loop only, e.g. the CPU tests with the oracle)
        n = self.n_vars + self.n_ineq
        return LinearOperator(
            (n, n), dtype=np.float64,
            matvec=lambda vec: np.hstack((Hx.dot(vec[:self.n_vars]), tail * vec[self.n_vars:])))

    def stop_criteria(self, state):
        """(2.3) of BHN (reference :243-251)."""
        return ((state.optimality < self.tolerance and state.constr_violation < self.tolerance)
                or self.global_stop_criteria(state) or state.trust_radius < self.xtol)


def tr_interior_point(fun, grad, lagr_hess, n_vars, n_ineq, n_eq, constr, jac, x0, fun0, grad0,
                      constr_ineq0, jac_ineq0, constr_eq0, jac_eq0, stop_criteria,
                      enforce_feasibility, xtol, state, initial_barrier_parameter=0.1,
                      initial_tolerance=0.1, initial_penalty=1.0, initial_trust_radius=1.0,
                      return_all=False, factorization_method=None, hot_path=None):
    """minimize fun(x) s.t. c_ineq(x) <= 0, c_eq(x) = 0 by a sequence of barrier
    subproblems, each solved with the trust-region SQP (reference :254-355)."""
    if enforce_feasibility is None:
        enforce_feasibility = np.zeros(n_ineq, bool)
    state.barrier_parameter, state.tolerance = initial_barrier_parameter, initial_tolerance
    state.trust_radius, state.penalty = initial_trust_radius, initial_penalty
    state.optimality = state.constr_violation = np.inf
    s0 = np.maximum(-1.5 * constr_ineq0, np.ones(n_ineq))            # :294
    sub = BarrierSubproblem(x0, s0, fun, grad, lagr_hess, n_vars, n_ineq, n_eq, constr, jac,
                            state.barrier_parameter, state.tolerance, enforce_feasibility,
                            stop_criteria, xtol, fun0, grad0, constr_ineq0, jac_ineq0,
                            constr_eq0, jac_eq0)
    z = np.hstack((x0, s0))
    f_z, c_z, g_z, J_z = sub.fun0, sub.constr0, sub.grad0, sub.jac0
    step_lb = np.hstack((np.full(n_vars, -np.inf), np.full(n_ineq, -TAU_BOUNDARY)))
    step_ub = np.full(n_vars + n_ineq, np.inf)
    first = True
    while True:
        if not first:
            state.trust_radius = max(initial_trust_radius, RADIUS_RESTART * state.trust_radius)
            state.barrier_parameter *= THETA_DECAY
            state.tolerance *= THETA_DECAY
        first = False
        sub.update(state.barrier_parameter, state.tolerance)
        state = equality_constrained_sqp(
            sub.function_and_constraints, sub.gradient_and_jacobian, sub.lagrangian_hessian,
            z, f_z, g_z, c_z, J_z, sub.stop_criteria, state, step_lb, step_ub,
            initial_penalty, state.trust_radius, sub.scaling, return_all,
            factorization_method, hot_path=hot_path)
        z = state.x
        if stop_criteria(state):
            break
        f_z, c_z = sub.function_and_constraints(z)
        g_z, J_z = sub.gradient_and_jacobian(z)
    state.x, state.s = sub.get_variables(z), sub.get_slack(z)
    if return_all:
        zs = state.allvecs
        state.allvecs = [sub.get_variables(zz) for zz in zs]
        state.allslack = [sub.get_slack(zz) for zz in zs]
    return state
