"""Device-backed trust-region EQP routines -- the reference's
``ipsolver/_large_scale_constrained/qp_subproblem.py`` interface (same names,
argument meaning and error behaviour), each call one persistent CUDA kernel.

``projected_cg`` runs the whole CG-Steihaug loop (H products, null-space
projections with their ``(A A')^-1`` solves, reductions, sphere/box
intersections, termination logic) on the device; only ``x`` and the ``info``
scalars return.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sps
from scipy.sparse.linalg import LinearOperator

from .device import Context, DeviceCSR, DeviceHessian
from ._cabi import dptr, f64
from .projections import DeviceMap, _fingerprint

__all__ = [
    'sphere_intersections', 'box_intersections', 'box_sphere_intersections',
    'inside_box_boundaries', 'reinforce_box_boundaries', 'modified_dogleg',
    'projected_cg', 'HessianOperator', 'as_device_hessian',
]


# ------------------------------------------------------------------ Hessian
class HessianOperator(LinearOperator):
    """Lagrangian Hessian as a LinearOperator that *carries its terms* so it
    can be applied on the device (the reference's operator is an opaque
    closure, ``_canonical_constraint.py:119-139``; SURVEY.md finding 6).

        H v = [ (sum of sparse terms + sum of dense terms) v_x ; diag_tail * v_s ]

    ``terms`` are scipy sparse matrices / ndarrays of shape (n_x, n_x);
    ``diag_tail`` holds the ``S Hs S`` diagonal of the barrier subproblem
    (``tr_interior_point.py:206-241``).  Sparse terms are summed into one CSR
    and dense terms into one array on the host before the upload, so the
    device does one SpMV and/or one GEMV per product.
    """

    def __init__(self, n, terms=(), diag_tail=None, ctx=None):
        self.n = int(n)
        self.diag_tail = None if diag_tail is None else f64(diag_tail)
        self.n_x = self.n - (0 if self.diag_tail is None else self.diag_tail.shape[0])
        super().__init__(dtype=np.float64, shape=(self.n, self.n))
        sparse_sum, dense_sum = None, None
        for t in terms:
            if sps.issparse(t):
                t = sps.csr_matrix(t, dtype=np.float64)
                sparse_sum = t if sparse_sum is None else sparse_sum + t
            else:
                t = np.atleast_2d(np.asarray(t, dtype=np.float64))
                dense_sum = t if dense_sum is None else dense_sum + t
        self.sparse = sparse_sum
        self.dense = dense_sum
        self._ctx = ctx
        self._dev = None
        self._csr = None
        self._token = object()

    def with_diag_tail(self, diag_tail):
        """``blockdiag(self, diag(diag_tail))`` -- the scaled barrier Hessian
        (``tr_interior_point.py:222-241``); the matrix terms are shared."""
        if self.diag_tail is not None:
            raise ValueError("operator already has a diagonal tail")
        out = HessianOperator.__new__(HessianOperator)
        out.diag_tail = f64(diag_tail)
        out.n_x = self.n
        out.n = self.n + out.diag_tail.shape[0]
        LinearOperator.__init__(out, dtype=np.float64, shape=(out.n, out.n))
        out.sparse, out.dense, out._ctx, out._dev = self.sparse, self.dense, self._ctx, None
        out._csr, out._token = None, object()
        return out

    @property
    def device(self):
        """Device operator.  The CSR term is cached per sparsity pattern on the
        context: a new Hessian with the pattern of the previous one (every outer
        iteration of the SQP / barrier loops) only re-uploads its values instead
        of rebuilding row blocks and descriptors.  The cached matrix remembers
        which operator wrote its values last: an operator that finds another
        one's values there uploads its own again before it is applied, so two
        live operators with one pattern never act for each other."""
        if self._dev is None:
            ctx = self._ctx or Context.default()
            csr = None
            if self.sparse is not None:
                S = self.sparse
                cache = ctx.__dict__.setdefault("_hess_csr_cache", {})
                key = (S.shape, S.nnz)
                csr = cache.pop(key, None)
                fp = _fingerprint(S)
                if csr is not None and getattr(csr, "host_arrays", None) is not None \
                        and csr.host_arrays[0] is S.data and csr.host_arrays[1] is S.indices \
                        and csr.host_arrays[2] is S.indptr and csr.host_fp == fp:
                    pass        # the very same arrays as last time (a constant Hessian): nothing to do
                else:
                    if not S.has_canonical_format:
                        S = S.copy()
                        S.sum_duplicates()
                        self.sparse = S
                    if csr is not None and csr.same_pattern(S):
                        csr.update_values(S.data)
                    else:
                        csr = DeviceCSR(ctx, S, with_transpose=False)
                    csr.host_arrays, csr.host_fp = (S.data, S.indices, S.indptr), _fingerprint(S)
                cache[key] = csr                               # most recently used last
                while len(cache) > 4:
                    cache.pop(next(iter(cache)))
                csr.last_writer = self._token
            self._csr = csr
            self._dev = DeviceHessian(ctx, self.n, self.n_x, csr=csr,
                                      dense=self.dense, diag_tail=self.diag_tail)
        elif self._csr is not None and self._csr.last_writer is not self._token:
            self._csr.update_values(self.sparse.data)
            self._csr.host_arrays = (self.sparse.data, self.sparse.indices, self.sparse.indptr)
            self._csr.host_fp = _fingerprint(self.sparse)
            self._csr.last_writer = self._token
        return self._dev

    def _matvec(self, x):
        return self.device.matvec(np.asarray(x, dtype=np.float64).reshape(-1))


def as_device_hessian(H, n):
    """Device Hessian for anything with an explicit matrix form, else None
    (an opaque operator: ``projected_cg`` then runs its loop on the host with
    the device projectors)."""
    if isinstance(H, HessianOperator):
        return H.device
    if isinstance(H, DeviceHessian):
        return H
    if sps.issparse(H) or isinstance(H, np.ndarray):
        return HessianOperator(n, [H]).device
    if isinstance(H, LinearOperator) or callable(getattr(H, "dot", None)):
        return None
    raise TypeError("projected_cg: H must be a matrix or a LinearOperator, got %s"
                    % type(H).__name__)


# ------------------------------------------------------------------ intersections
def _intersections(z, d, lb, ub, trust_radius, entire_line):
    ctx = Context.default()
    z = f64(z)
    d = f64(d, z.shape[0])
    if lb is not None or ub is not None:
        n = z.shape[0]
        lb = np.full(n, -np.inf) if lb is None else f64(lb, n)
        ub = np.full(n, np.inf) if ub is None else f64(ub, n)
    ta, tb, hit = C.c_double(), C.c_double(), C.c_int()
    ctx.check(ctx.lib.ipb_box_sphere_intersections(
        ctx.h, z.shape[0], dptr(z), dptr(d), dptr(lb), dptr(ub),
        float(trust_radius), int(bool(entire_line)),
        C.byref(ta), C.byref(tb), C.byref(hit)))
    return ta.value, tb.value, bool(hit.value)


def sphere_intersections(z, d, trust_radius, entire_line=False):
    """Reference ``qp_subproblem.py:66-149``."""
    return _intersections(z, d, None, None, trust_radius, entire_line)


def box_intersections(z, d, lb, ub, entire_line=False):
    """Reference ``qp_subproblem.py:152-234``."""
    return _intersections(z, d, lb, ub, np.inf, entire_line)


def box_sphere_intersections(z, d, lb, ub, trust_radius, entire_line=False,
                             extra_info=False):
    """Reference ``qp_subproblem.py:237-303``."""
    ta, tb, hit = _intersections(z, d, lb, ub, trust_radius, entire_line)
    if not extra_info:
        return ta, tb, hit
    ta_s, tb_s, hit_s = _intersections(z, d, None, None, trust_radius, entire_line)
    ta_b, tb_b, hit_b = _intersections(z, d, lb, ub, np.inf, entire_line)
    return (ta, tb, hit, {'ta': ta_s, 'tb': tb_s, 'intersect': hit_s},
            {'ta': ta_b, 'tb': tb_b, 'intersect': hit_b})


def inside_box_boundaries(x, lb, ub):
    """Reference ``qp_subproblem.py:306-308`` (host-side scalar test used by
    the outer loop on vectors it already holds)."""
    return bool((lb <= x).all() and (x <= ub).all())


def reinforce_box_boundaries(x, lb, ub):
    """Reference ``qp_subproblem.py:310-317``."""
    return np.minimum(np.maximum(x, lb), ub)


# ------------------------------------------------------------------ solvers
def _projector_of(*ops):
    binding = None
    for op in ops:
        if not isinstance(op, DeviceMap):
            raise TypeError(
                "Z and Y must come from ipsolver_b200.projections.projections() "
                "(got %s); there is no host fallback." % type(op).__name__)
        if binding is None:
            binding = op.binding
        elif op.binding is not binding:
            raise ValueError("Z and Y belong to different projections() calls")
    return binding.ensure()


def modified_dogleg(A, Y, b, trust_radius, lb, ub):
    """Reference ``qp_subproblem.py:320-413``; ``A`` must be the matrix ``Y``
    was built from (it is taken from the device projector)."""
    proj = _projector_of(Y)
    return proj.modified_dogleg(b, trust_radius, lb, ub)


def projected_cg(H, c, Z, Y, b, trust_radius=np.inf, lb=None, ub=None,
                 tol=None, max_iter=None, max_infeasible_iter=None,
                 return_all=False):
    """Reference ``qp_subproblem.py:416-643``.  Raises ``ValueError`` in the
    two situations the reference does (:517, :560)."""
    proj = _projector_of(Z, Y)
    c = np.asarray(c, dtype=np.float64)
    Hd = as_device_hessian(H, c.shape[0])
    if Hd is None:
        return _projected_cg_opaque(H, c, Z, Y, np.asarray(b, dtype=np.float64), trust_radius,
                                    lb, ub, tol, max_iter, max_infeasible_iter, return_all)
    return proj.projected_cg(Hd, c, b, trust_radius, lb, ub, tol, max_iter,
                             max_infeasible_iter, return_all)


def _projected_cg_opaque(H, c, Z, Y, b, trust_radius, lb, ub, tol, max_iter,
                         max_infeasible_iter, return_all):
    """``projected_cg`` for a Hessian that is only a callback (a user
    ``LinearOperator``, a finite-difference Hessian -- reference
    ``_minimize_constrained.py:404-406,417-419``): ``H.dot`` has to run on the
    host, so the CG-Steihaug recurrence (Gould, Hribar & Nocedal 2001, Alg. 6.2,
    with the sphere / box truncation of reference ``qp_subproblem.py:416-643``)
    is driven from here, one iteration per ``H.dot``, while every projection
    ``Z.dot`` / ``Y.dot`` (the inner solves with A A') and every intersection
    still runs on the device."""
    n, m = c.shape[0], b.shape[0]
    lb = np.full(n, -np.inf) if lb is None else np.asarray(lb, dtype=np.float64)
    ub = np.full(n, np.inf) if ub is None else np.asarray(ub, dtype=np.float64)
    x = Y.dot(-b)
    r = Z.dot(H.dot(x) + c)
    g = Z.dot(r)
    p = -g
    trail = [x] if return_all else None
    Hp = H.dot(p)
    gg = np.linalg.norm(g) ** 2
    slack = trust_radius - np.linalg.norm(x)
    if slack < -8.0 * np.finfo(float).eps * trust_radius:
        raise ValueError("Trust region problem does not have a solution.")
    info = {'niter': 0, 'stop_cond': 1, 'hits_boundary': False}
    if slack < 1e-25:                       # x = Y(-b) already on the boundary
        info.update(stop_cond=2, hits_boundary=True)
        if return_all:
            trail.append(x)
            info['allvecs'] = trail
        return x, info
    if tol is None:
        tol = max(min(0.01 * np.sqrt(gg), 0.1 * gg), 1e-25)
    cap = n - m
    n_iter = cap if max_iter is None else min(max_iter, cap)
    patience = cap if max_infeasible_iter is None else max_infeasible_iter
    outside_for = 0
    fallback = np.zeros_like(x)             # last point known to satisfy the bounds
    for _ in range(max(n_iter, 0)):
        if gg < tol:
            info['stop_cond'] = 4
            break
        info['niter'] += 1
        curv = Hp.dot(p)
        if curv <= 0:                       # follow p to the boundary of the region
            if np.isinf(trust_radius):
                raise ValueError("Negative curvature not allowed for unrestrited problems.")
            _, t_far, ok = box_sphere_intersections(x, p, lb, ub, trust_radius, entire_line=True)
            if ok:
                x = x + t_far * p
            x = reinforce_box_boundaries(x, lb, ub)
            info.update(stop_cond=3, hits_boundary=True)
            break
        alpha = gg / curv
        x_new = x + alpha * p
        if np.linalg.norm(x_new) >= trust_radius:      # leaves the ball: stop on it
            _, frac, ok = box_sphere_intersections(x, alpha * p, lb, ub, trust_radius)
            if ok:
                x = x + frac * alpha * p
            x = reinforce_box_boundaries(x, lb, ub)
            info.update(stop_cond=2, hits_boundary=True)
            break
        outside_for = 0 if inside_box_boundaries(x_new, lb, ub) else outside_for + 1
        if outside_for > 0:                 # remember where the segment leaves the box
            _, frac, ok = box_sphere_intersections(x, alpha * p, lb, ub, trust_radius)
            if ok:
                fallback = reinforce_box_boundaries(x + frac * alpha * p, lb, ub)
                outside_for = 0
        if outside_for > patience:
            break
        if return_all:
            trail.append(x_new)
        g_new = Z.dot(r + alpha * Hp)
        gg_new = np.linalg.norm(g_new) ** 2
        p = -g_new + (gg_new / gg) * p
        x, r = x_new, g_new                 # the projected residual is the one kept
        gg = np.linalg.norm(g_new) ** 2
        Hp = H.dot(p)
    if not inside_box_boundaries(x, lb, ub):
        x = fallback
        info['hits_boundary'] = True
    if return_all:
        info['allvecs'] = trail
    return x, info
