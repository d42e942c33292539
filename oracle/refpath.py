"""CPU restatement of the reference hot path (numpy + scipy, single thread).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Every function names the reference lines it follows (paths relative to
``/root/reference/ipsolver/_large_scale_constrained/``).  The arithmetic is
kept in the reference's evaluation order (same numpy calls, same operand
order) so that results are bit-identical to the reference on the same
machine; ``tests/test_oracle_vs_golden.py`` holds it to that.

Quirks that are reproduced on purpose:
  * the CG residual that is stored is the *projected* one (``r = g_next``);
  * squared norms are ``norm(v)**2`` (sqrt, then square), not ``v.v``;
  * ``max_iter`` is capped at ``n - m``;
  * the "last feasible point" starts as uninitialised memory in the
    reference (``np.empty_like``); the oracle uses zeros (what SciPy's
    descendant does) -- the value is only observable when the loop ends
    outside the box without ever having found a feasible point.
"""
from math import copysign

import numpy as np
import scipy.linalg
import scipy.sparse as sps
import scipy.sparse.linalg as spla
from numpy.linalg import norm

_TINY = 1e-25  # qp_subproblem.py:496


# --------------------------------------------------------------------------
# segment / line intersections
# --------------------------------------------------------------------------
def segment_ball(z, d, radius, entire_line=False):
    """Interval of t with ``||z + t d|| <= radius``.  qp_subproblem.py:66-149."""
    if norm(d) == 0:                                   # :99-100
        return 0, 0, False
    if np.isinf(radius):                               # :102-110
        if entire_line:
            return -np.inf, np.inf, True
        return 0, 1, True
    qa = np.dot(d, d)                                  # :112-115
    qb = 2 * np.dot(z, d)
    qc = np.dot(z, z) - radius**2
    disc = qb*qb - 4*qa*qc
    if disc < 0:                                       # :116-118
        return 0, 0, False
    root = np.sqrt(disc)
    # cancellation-free pair of roots, :128-131
    aux = qb + copysign(root, qb)
    t_lo = -aux / (2*qa)
    t_hi = -2*qc / aux
    t_lo, t_hi = sorted([t_lo, t_hi])
    if entire_line:                                    # :133-134
        return t_lo, t_hi, True
    if t_hi < 0 or t_lo > 1:                           # :138-141
        return 0, 0, False
    return max(0, t_lo), min(1, t_hi), True            # :143-147


def segment_box(z, d, lb, ub, entire_line=False):
    """Interval of t with ``lb <= z + t d <= ub``.  qp_subproblem.py:152-234."""
    z = np.asarray(z)
    d = np.asarray(d)
    lb = np.asarray(lb)
    ub = np.asarray(ub)
    if norm(d) == 0:                                   # :194-195
        return 0, 0, False
    still = (d == 0)                                   # :198-203
    if (z[still] < lb[still]).any() or (z[still] > ub[still]).any():
        return 0, 0, False
    moving = np.logical_not(still)                     # :205-209
    z, d, lb, ub = z[moving], d[moving], lb[moving], ub[moving]
    t_at_lb = (lb - z) / d                             # :212-216
    t_at_ub = (ub - z) / d
    t_lo = max(np.minimum(t_at_lb, t_at_ub))
    t_hi = min(np.maximum(t_at_lb, t_at_ub))
    hit = bool(t_lo <= t_hi)                           # :219-222
    if not entire_line:                                # :224-232
        if t_hi < 0 or t_lo > 1:
            return 0, 0, False
        t_lo = max(0, t_lo)
        t_hi = min(1, t_hi)
    return t_lo, t_hi, hit


def segment_box_ball(z, d, lb, ub, radius, entire_line=False):
    """Intersection of the two intervals above.  qp_subproblem.py:237-303."""
    lo_b, hi_b, hit_b = segment_box(z, d, lb, ub, entire_line)
    lo_s, hi_s, hit_s = segment_ball(z, d, radius, entire_line)
    t_lo = np.maximum(lo_b, lo_s)
    t_hi = np.minimum(hi_b, hi_s)
    hit = bool(hit_b and hit_s and t_lo <= t_hi)
    return t_lo, t_hi, hit


def in_box(x, lb, ub):
    """qp_subproblem.py:306-308."""
    return (lb <= x).all() and (x <= ub).all()


def clip_to_box(x, lb, ub):
    """qp_subproblem.py:310-317."""
    return np.minimum(np.maximum(x, lb), ub)


# --------------------------------------------------------------------------
# normal step
# --------------------------------------------------------------------------
def dogleg_normal_step(A, Y, b, radius, lb, ub):
    """Approximate minimiser of ``||A x + b||`` in ball-and-box.

    qp_subproblem.py:320-413.
    """
    newton = -Y.dot(b)                                 # :368
    if in_box(newton, lb, ub) and norm(newton) <= radius:   # :370-373
        return newton
    g = A.T.dot(b)                                     # :376
    Ag = A.dot(g)                                      # :379
    cauchy = -np.dot(g, g) / np.dot(Ag, Ag) * g        # :380
    origin = np.zeros_like(cauchy)
    # segment cauchy -> newton, :386-391
    step = newton - cauchy
    _, t, hit = segment_box_ball(cauchy, step, lb, ub, radius)
    if hit:
        cand1 = cauchy + t*step
    else:
        # segment origin -> cauchy, :395-399
        _, t, _ = segment_box_ball(origin, cauchy, lb, ub, radius)
        cand1 = origin + t*cauchy
    # segment origin -> newton, :403-407
    _, t, _ = segment_box_ball(origin, newton, lb, ub, radius)
    cand2 = origin + t*newton
    if norm(A.dot(cand1) + b) < norm(A.dot(cand2) + b):     # :410-413
        return cand1
    return cand2


# --------------------------------------------------------------------------
# projected CG (CG-Steihaug in the null space of A)
# --------------------------------------------------------------------------
def eqp_projected_cg(H, c, Z, Y, b, trust_radius=np.inf, lb=None, ub=None,
                     tol=None, max_iter=None, max_infeasible_iter=None,
                     return_all=False):
    """min 1/2 x'Hx + c'x  s.t.  Ax + b = 0, ||x|| <= radius, lb <= x <= ub.

    qp_subproblem.py:416-643.  Returns ``(x, info)``.
    """
    n, = np.shape(c)
    m, = np.shape(b)

    x = Y.dot(-b)                                      # :502
    r = Z.dot(H.dot(x) + c)                            # :503
    g = Z.dot(r)                                       # :504
    p = -g                                             # :505
    trail = [x] if return_all else None
    Hp = H.dot(p)                                      # :511
    rg = norm(g)**2                                    # :512

    room = trust_radius - norm(x)                      # :515-526
    if room < 0:
        raise ValueError("Trust region problem does not have a solution.")
    if room < _TINY:
        info = {'niter': 0, 'stop_cond': 2, 'hits_boundary': True}
        if return_all:
            trail.append(x)
            info['allvecs'] = trail
        return x, info

    if tol is None:                                    # :529-530
        tol = max(min(0.01 * np.sqrt(rg), 0.1 * rg), _TINY)
    if lb is None:                                     # :532-535
        lb = np.full(n, -np.inf)
    if ub is None:
        ub = np.full(n, np.inf)
    if max_iter is None:                               # :537-539
        max_iter = n - m
    max_iter = min(max_iter, n - m)
    if max_infeasible_iter is None:                    # :541-542
        max_infeasible_iter = n - m

    on_boundary = False
    reason = 1
    outside = 0
    # the reference leaves this uninitialised (:547); zeros here.
    x_feasible = np.zeros_like(x)
    k = 0
    for _ in range(max_iter):
        if rg < tol:                                   # :551-553
            reason = 4
            break
        k += 1
        curv = Hp.dot(p)                               # :556
        if curv <= 0:                                  # :558-576
            if np.isinf(trust_radius):
                raise ValueError("Negative curvature not allowed "
                                 "for unrestrited problems.")
            _, alpha, hit = segment_box_ball(x, p, lb, ub, trust_radius,
                                             entire_line=True)
            if hit:
                x = x + alpha*p
            x = clip_to_box(x, lb, ub)
            reason = 3
            on_boundary = True
            break

        alpha = rg / curv                              # :579-580
        x_new = x + alpha*p

        if np.linalg.norm(x_new) >= trust_radius:      # :583-596
            _, theta, hit = segment_box_ball(x, alpha*p, lb, ub, trust_radius)
            if hit:
                x = x + theta*alpha*p
            x = clip_to_box(x, lb, ub)
            reason = 2
            on_boundary = True
            break

        if in_box(x_new, lb, ub):                      # :599-602
            outside = 0
        else:
            outside += 1
        if outside > 0:                                # :604-613
            _, theta, hit = segment_box_ball(x, alpha*p, lb, ub, trust_radius)
            if hit:
                x_feasible = x + theta*alpha*p
                x_feasible = clip_to_box(x_feasible, lb, ub)
                outside = 0
        if outside > max_infeasible_iter:              # :615-616
            break
        if return_all:
            trail.append(x_new)

        r_new = r + alpha*Hp                           # :622
        g_new = Z.dot(r_new)                           # :624
        rg_new = norm(g_new)**2                        # :626
        beta = rg_new / rg                             # :627
        p = -g_new + beta*p                            # :628
        x = x_new                                      # :630-634
        g = g_new
        r = g_new
        rg = norm(g)**2
        Hp = H.dot(p)

    if not in_box(x, lb, ub):                          # :636-638
        x = x_feasible
        on_boundary = True
    info = {'niter': k, 'stop_cond': reason, 'hits_boundary': on_boundary}
    if return_all:
        info['allvecs'] = trail
    return x, info


def kkt_direct_solve(H, c, A, b):
    """Direct KKT solve used by the reference tests as a comparator.

    qp_subproblem.py:21-63.
    """
    n, = np.shape(c)
    m, = np.shape(b)
    K = sps.csc_matrix(sps.bmat([[H, A.T], [A, None]]))
    sol = spla.splu(K).solve(np.hstack([-c, -b]))
    return sol[:n], -sol[n:n+m]


# --------------------------------------------------------------------------
# projectors
# --------------------------------------------------------------------------
def orthogonality_measure(A, g):
    """``||A g|| / (||A||_F ||g||)``; 0 when either norm is 0.

    projections.py:23-55.
    """
    ng = np.linalg.norm(g)
    if sps.issparse(A):
        nA = spla.norm(A, ord='fro')
    else:
        nA = np.linalg.norm(A, ord='fro')
    if ng == 0 or nA == 0:
        return 0
    return np.linalg.norm(A.dot(g)) / (nA*ng)


def _refine_normal(A, z, solve_aat, orth_tol, max_refin):
    """Refinement loop shared by the AA' based variants.  projections.py:71-78."""
    k = 0
    while orthogonality_measure(A, z) > orth_tol:
        if k >= max_refin:
            break
        v = solve_aat(A.dot(z))
        z = z - A.T.dot(v)
        k += 1
    return z


def _normal_equation_maps(A, m, n, orth_tol, max_refin, tol, solve_aat=None):
    """projections.py:58-90 with the AA' solve injected.

    The reference uses CHOLMOD (``cholesky_AAt``, :62), absent from this image;
    the default here is SuperLU on ``A A'`` -- same maps, different round-off.
    """
    if solve_aat is None:
        solve_aat = spla.factorized(sps.csc_matrix(A.dot(A.T)))

    def null_space(x):
        v = solve_aat(A.dot(x))
        z = x - A.T.dot(v)
        return _refine_normal(A, z, solve_aat, orth_tol, max_refin)

    def least_squares(x):
        return solve_aat(A.dot(x))

    def row_space(x):
        return A.T.dot(solve_aat(x))

    return null_space, least_squares, row_space


def _augmented_system_maps(A, m, n, orth_tol, max_refin, tol):
    """projections.py:93-172 (SuperLU on [[I, A'], [A, 0]])."""
    K = sps.csc_matrix(sps.bmat([[sps.eye(n), A.T], [A, None]]))
    try:
        solve = spla.factorized(K)
    except RuntimeError:
        return _svd_maps(A.toarray(), m, n, orth_tol, max_refin, tol)

    def null_space(x):
        rhs = np.hstack([x, np.zeros(m)])              # :117
        sol = solve(rhs)                               # :120
        z = sol[:n]
        k = 0
        while orthogonality_measure(A, z) > orth_tol:  # :126-139
            if k >= max_refin:
                break
            resid = rhs - K.dot(sol)
            sol += solve(resid)
            z = sol[:n]
            k += 1
        return z

    def least_squares(x):
        rhs = np.hstack([x, np.zeros(m)])              # :151-156
        return solve(rhs)[n:m+n]

    def row_space(x):
        rhs = np.hstack([np.zeros(n), x])              # :165-170
        return solve(rhs)[:n]

    return null_space, least_squares, row_space


def _qr_maps(A, m, n, orth_tol, max_refin, tol):
    """projections.py:175-233 (dense, pivoted QR of A')."""
    Q, R, P = scipy.linalg.qr(A.T, pivoting=True, mode='economic')
    if np.linalg.norm(R[-1, :], np.inf) < tol:
        return _svd_maps(A, m, n, orth_tol, max_refin, tol)

    def apply_pinv_t(x, out):
        w = scipy.linalg.solve_triangular(R, Q.T.dot(x), lower=False)
        out[P] = w
        return out

    def null_space(x):
        v = apply_pinv_t(x, np.zeros(m))
        z = x - A.T.dot(v)
        k = 0
        while orthogonality_measure(A, z) > orth_tol:
            if k >= max_refin:
                break
            v = apply_pinv_t(z, v)
            z = z - A.T.dot(v)
            k += 1
        return z

    def least_squares(x):
        return apply_pinv_t(x, np.zeros(m))

    def row_space(x):
        w = scipy.linalg.solve_triangular(R, x[P], lower=False, trans='T')
        return Q.dot(w)

    return null_space, least_squares, row_space


def _svd_maps(A, m, n, orth_tol, max_refin, tol):
    """projections.py:236-287 (dense, thin SVD with small values dropped)."""
    U, s, Vt = scipy.linalg.svd(A, full_matrices=False)
    keep = s > tol
    U, Vt, s = U[:, keep], Vt[keep, :], s[keep]

    def pinv_t(x):
        return U.dot(1/s*Vt.dot(x))

    def null_space(x):
        z = x - A.T.dot(pinv_t(x))
        k = 0
        while orthogonality_measure(A, z) > orth_tol:
            if k >= max_refin:
                break
            z = z - A.T.dot(pinv_t(z))
            k += 1
        return z

    def least_squares(x):
        return pinv_t(x)

    def row_space(x):
        return Vt.T.dot(1/s*U.T.dot(x))

    return null_space, least_squares, row_space


def build_projectors(A, method=None, orth_tol=1e-12, max_refin=3, tol=1e-15,
                     solve_aat=None):
    """``(Z, LS, Y)`` LinearOperators for A.  projections.py:290-406.

    ``NormalEquation`` does not degrade to ``AugmentedSystem`` here (the
    reference does when scikit-sparse is missing, :377-382): it runs the
    normal-equation maps with SuperLU (or ``solve_aat``) in CHOLMOD's place.
    """
    m, n = np.shape(A)
    if m*n == 0:                                       # :368-369
        A = sps.csc_matrix(A)
    if sps.issparse(A):                                # :372-387
        if method is None:
            method = "AugmentedSystem"
        if method not in ("NormalEquation", "AugmentedSystem"):
            raise ValueError("Method not allowed for sparse matrix.")
    else:
        if method is None:
            method = "QRFactorization"
        if method not in ("QRFactorization", "SVDFactorization"):
            raise ValueError("Method not allowed for dense array.")
    if method == "NormalEquation":
        maps = _normal_equation_maps(A, m, n, orth_tol, max_refin, tol,
                                     solve_aat)
    elif method == "AugmentedSystem":
        maps = _augmented_system_maps(A, m, n, orth_tol, max_refin, tol)
    elif method == "QRFactorization":
        maps = _qr_maps(A, m, n, orth_tol, max_refin, tol)
    else:
        maps = _svd_maps(A, m, n, orth_tol, max_refin, tol)
    null_space, least_squares, row_space = maps
    # the reference omits dtype (:402-404), which makes scipy probe each map
    # with an int8 zero vector; same here so call counts agree.
    Z = spla.LinearOperator((n, n), null_space)
    LS = spla.LinearOperator((m, n), least_squares)
    Y = spla.LinearOperator((n, m), row_space)
    return Z, LS, Y


# --------------------------------------------------------------------------
# host model of the Lagrangian Hessian operators and the whole hot-path bundle
# --------------------------------------------------------------------------
class HostLagrangianHessian(spla.LinearOperator):
    """What the reference's Hessian closures compute
    (``ipsolver/_canonical_constraint.py:119-139``: the products of the terms
    added up one after the other; ``tr_interior_point.py:222-241``: plus the
    diagonal slack block)."""

    def __init__(self, n, terms, diag_tail=None):
        self.terms, self.diag_tail = list(terms), diag_tail
        self.n_x = n
        n_all = n + (0 if diag_tail is None else len(diag_tail))
        super().__init__(dtype=np.float64, shape=(n_all, n_all))

    def with_diag_tail(self, tail):
        return HostLagrangianHessian(self.n_x, self.terms, np.asarray(tail, dtype=float))

    def _matvec(self, p):
        p = np.asarray(p, dtype=float).ravel()
        px = p[:self.n_x]
        out = np.zeros_like(px)
        for h in self.terms:
            out += h.dot(px)
        if self.diag_tail is None:
            return out
        return np.hstack((out, self.diag_tail * p[self.n_x:]))


def host_hot_path():
    """The oracle's routines bundled the way the outer loops of
    ``ipsolver_b200`` take them (``hot_path=``): used by the CPU tests and by
    the CPU-baseline legs of ``bench.py`` only."""
    from types import SimpleNamespace
    return SimpleNamespace(
        projections=lambda A, method=None: build_projectors(A, method),
        projected_cg=eqp_projected_cg, modified_dogleg=dogleg_normal_step,
        box_intersections=segment_box, hessian=HostLagrangianHessian,
        # the Jacobian products of the SQP call sites (reference
        # equality_constrained_sqp.py:86,140,238): scipy's csr_matvec
        jac_dot=lambda A, Z, x: A.dot(x), jac_tdot=lambda A, Z, y: A.T.dot(y))
