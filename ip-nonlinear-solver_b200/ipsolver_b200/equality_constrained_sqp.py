"""Byrd-Omojokun trust-region SQP -- the CALLER of the device hot path.

Host control logic with the interface of the reference's
``ipsolver/_large_scale_constrained/equality_constrained_sqp.py:18-27`` (same
arguments, same ``state`` fields and counters), re-stated from the algorithm
(Lalee-Nocedal-Plantenga 1998; Byrd-Hribar-Nocedal 1999, sec. 3) with the
reference's constants so that iteration counts can be compared.  Every linear
algebra step -- ``projections``, ``LS.dot``, ``modified_dogleg``,
``projected_cg``, ``Y.dot``, ``box_intersections``, ``A.dot`` / ``A.T.dot`` --
goes to the device
through ``ipsolver_b200``; user callbacks (``fun_and_constr``,
``grad_and_jac``, ``lagr_hess``) run on the host between those calls.

``hot_path`` is the object that supplies those functions: by default the
single-GPU device path below; ``distributed.DistSolver.hot_path()`` for a
row-partitioned run (every rank runs this loop on the same host vectors, each
device step works on the rank's own blocks); the CPU tests pass the oracle's
routines to check the host logic against the reference's golden results
without a GPU.
"""
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sps
from numpy.linalg import norm

__all__ = ['equality_constrained_sqp']

# algorithm constants (reference equality_constrained_sqp.py:50-60)
RHO_PENALTY = 0.3            # rho of (3.51) in Byrd-Hribar-Nocedal
ETA_ACCEPT = 1e-8            # sufficient-reduction ratio
RATIO_LARGE, RATIO_MID = 0.9, 0.3
GROW_LARGE, GROW_MID = 7.0, 2.0
SHRINK_MAX, SHRINK_MIN = 0.5, 0.1
SOC_RATIO = 0.1
ZETA_NORMAL = 0.8            # share of the trust region given to the normal step
BOX_NORMAL = 0.5


def device_hot_path():
    from . import projections as _p, qp_subproblem as _q

    def jac_dot(A, Z, x):
        """``A x`` with the device copy of A that ``Z`` was built from (same
        row sums, bit for bit, as scipy's csr_matvec)."""
        return Z.projector.A.matvec(x) if A.shape[0] else np.zeros(0)

    def jac_tdot(A, Z, y):
        return Z.projector.A.rmatvec(y) if A.shape[0] else np.zeros(A.shape[1])

    def fused_step(H, Z, c, b, radius, lb, ub, bounds_changed):
        """Normal + tangential step on device-resident vectors:
        ``c`` / ``b`` None = unchanged since the last call.  The host only sees
        ``dn`` (to set the tangential radius exactly as the reference does,
        :127) and ``d, H d, A d`` (for the scalars of :138-140)."""
        proj = Z.projector
        dn = proj.sqp_normal_step(c, b, ZETA_NORMAL * radius, lb, ub, bounds_changed, BOX_NORMAL)
        radius_t = np.sqrt(radius**2 - np.linalg.norm(dn)**2)
        d, Hd, Ad, cg_info = proj.sqp_tangential_step(_q.as_device_hessian(H, proj.n), radius_t)
        return dn, d, Hd, Ad, cg_info

    def fused_multipliers(Z, c):
        return Z.projector.sqp_multipliers(c)

    return SimpleNamespace(projections=_p.projections, projected_cg=_q.projected_cg,
                           modified_dogleg=_q.modified_dogleg,
                           box_intersections=_q.box_intersections,
                           hessian=_q.HessianOperator, jac_dot=jac_dot, jac_tdot=jac_tdot,
                           fused_step=fused_step, fused_multipliers=fused_multipliers)


def _identity_scaling(x):
    return sps.eye(np.shape(x)[0])


def _next_radius(radius, ratio, step_norm):
    """Trust-radius update (reference :196-212; (3.55) of BHN and N&W p. 696)."""
    if ratio >= RATIO_LARGE:
        return max(GROW_LARGE * step_norm, radius)
    if ratio >= RATIO_MID:
        return max(GROW_MID * step_norm, radius)
    if ratio < ETA_ACCEPT:
        shrink = (1 - ETA_ACCEPT) / (1 - ratio)
        proposal = shrink * step_norm
        if proposal >= SHRINK_MAX * radius:
            return radius * SHRINK_MAX
        if proposal >= SHRINK_MIN * radius:
            return proposal
        return radius * SHRINK_MIN
    return radius


def _record(state, x, v, f, c, b, A, hp, Z, optimality=None):
    state.x, state.v, state.fun, state.grad, state.constr, state.jac = x, v, f, c, b, A
    state.optimality = (norm(c + hp.jac_tdot(A, Z, v), np.inf) if optimality is None
                        else optimality)
    state.constr_violation = norm(b, np.inf) if len(b) > 0 else 0


def equality_constrained_sqp(fun_and_constr, grad_and_jac, lagr_hess,
                             x0, fun0, grad0, constr0, jac0, stop_criteria, state,
                             trust_lb=None, trust_ub=None, initial_penalty=1.0,
                             initial_trust_radius=1.0, scaling=_identity_scaling,
                             return_all=False, factorization_method=None,
                             hot_path=None):
    """minimize fun(x) subject to constr(x) = 0 (reference :18-252)."""
    hp = hot_path or device_hot_path()
    n = np.shape(x0)[0]
    lb = np.full(n, -np.inf) if trust_lb is None else trust_lb
    ub = np.full(n, np.inf) if trust_ub is None else trust_ub

    x = np.copy(x0)
    radius, penalty = initial_trust_radius, initial_penalty
    f, c, b, A = fun0, grad0, constr0, jac0
    S = scaling(x)
    Z, LS, Y = hp.projections(A, factorization_method)      # :81
    # device hot path with matrix Hessians: the vectors of an iteration stay in HBM
    # (one upload of c per accepted step, one download of d per iteration)
    fused = hasattr(hp, "fused_step")
    c_up = b_up = bounds_up = True           # what the device copy still lacks
    bounded = bool(np.isfinite(lb).any() or np.isfinite(ub).any())
    if fused:
        v, opt0 = hp.fused_multipliers(Z, c)
        c_up = False
        _record(state, x, v, f, c, b, A, hp, Z, opt0)
    else:
        v = -LS.dot(c)                                       # least-squares multipliers :83
        _record(state, x, v, f, c, b, A, hp, Z)
    state.niter += 1
    state.trust_radius, state.penalty = radius, penalty
    if return_all:
        state.allvecs += [np.copy(x)]
        state.allmult += [np.copy(v)]

    need_hessian = True
    while not stop_criteria(state):
        if need_hessian:
            H = lagr_hess(x, v)
            state.nhev += 1
        use_fused = fused and (hp.can_fuse(H) if hasattr(hp, "can_fuse")
                               else getattr(H, "device", None) is not None)
        if use_fused:
            dn, d, Hd, Ad, cg_info = hp.fused_step(H, Z, c if c_up else None, b if b_up else None,
                                                   radius, lb if bounded else None,
                                                   ub if bounded else None, bounds_up)
            c_up = b_up = bounds_up = False
            dt = d - dn
            norm_dn, norm_dt, norm_d = norm(dn), norm(dt), norm(d)
            q_model = 1 / 2 * Hd.dot(d) + c.T.dot(d)
            norm_lin = norm(Ad + b)
        else:
            # normal step: reduce ||A d + b|| inside a shrunken region (:113)
            dn = hp.modified_dogleg(A, Y, b, ZETA_NORMAL * radius, BOX_NORMAL * lb, BOX_NORMAL * ub)
            # tangential step: trust-region EQP in the null space of A (:125-132)
            c_t = H.dot(dn) + c
            radius_t = np.sqrt(radius**2 - np.linalg.norm(dn)**2)
            dt, cg_info = hp.projected_cg(H, c_t, Z, Y, np.zeros_like(b), radius_t, lb - dn, ub - dn)
            d = dn + dt
            norm_dn, norm_dt, norm_d = norm(dn), norm(dt), norm(d)
            # model and linearised constraint at the step (:138-140)
            q_model = 1 / 2 * (H.dot(d)).dot(d) + c.T.dot(d)
            norm_lin = norm(hp.jac_dot(A, Z, d) + b)
        vpred = max(1e-16, norm(b) - norm_lin)               # (3.52), kept positive
        penalty_before = penalty
        if q_model > 0:
            penalty = max(penalty, q_model / ((1 - RHO_PENALTY) * vpred))
        predicted = -q_model + penalty * vpred
        merit = f + penalty * norm(b)
        # trial point (:158-169)
        x_trial = x + S.dot(d)
        f_trial, b_trial = fun_and_constr(x_trial)
        state.nfev += 1
        state.ncev += 1
        ratio = (merit - (f_trial + penalty * norm(b_trial))) / predicted
        # second-order correction (:172-193)
        if ratio < ETA_ACCEPT and norm_dn <= SOC_RATIO * norm_dt:
            y = -Y.dot(b_trial)
            _, t, hit = hp.box_intersections(d, y, lb, ub)
            x_soc = x + S.dot(d + t * y)
            f_soc, b_soc = fun_and_constr(x_soc)
            state.nfev += 1
            state.ncev += 1
            ratio_soc = (merit - (f_soc + penalty * norm(b_soc))) / predicted
            if hit and ratio_soc >= ETA_ACCEPT:
                x_trial, f_trial, b_trial, ratio = x_soc, f_soc, b_soc, ratio_soc
        radius = _next_radius(radius, ratio, norm_d)
        state.niter += 1
        if ratio >= ETA_ACCEPT:                              # accept (:216-238)
            x, f, b = x_trial, f_trial, b_trial
            c, A = grad_and_jac(x)
            S = scaling(x)
            state.ngev += 1
            state.njev += 1
            Z, LS, Y = hp.projections(A)                     # method dropped here, as in :225
            b_up = True
            if fused:
                v, opt_new = hp.fused_multipliers(Z, c)
                c_up = False
                _record(state, x, v, f, c, b, A, hp, Z, opt_new)
            else:
                v = -LS.dot(c)
                _record(state, x, v, f, c, b, A, hp, Z)
            need_hessian = True
        else:                                                # reject (:240-242)
            penalty = penalty_before
            need_hessian = False
        state.trust_radius, state.penalty = radius, penalty
        state.cg_niter += cg_info["niter"]
        state.cg_info = cg_info
        if return_all:
            state.allvecs.append(np.copy(x))
            state.allmult.append(np.copy(v))
    return state
