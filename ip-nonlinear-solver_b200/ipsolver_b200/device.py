"""Thin object layer over the C ABI: context, device CSR matrices, Hessian
operator, projector and device vectors.  Host code above this module never
sees ctypes."""
import ctypes as C
import os
import weakref

import numpy as np
import scipy.sparse as sps

from . import _cabi
from ._cabi import (DeviceLibraryError, PcgInfo, ProjInfo, dptr, f64,
                    IPB_ERR_TR_INFEASIBLE, IPB_ERR_NEG_CURVATURE,
                    IPB_ERR_INNER_SOLVE)


class InnerSolveError(RuntimeError):
    """No solver for ``(A A') v = rhs`` converged: the banded direct solve, the CG
    and the least-squares CG for singular ``A A'`` all failed."""


class Context:
    """One per process and GPU.  ``Context.default()`` picks LOCAL_RANK."""
    _default = None

    def __init__(self, device=None):
        self.lib = _cabi.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        st = self.lib.ipb_ctx_create(int(device), C.byref(h))
        self.h = h
        if st != 0:
            msg = self.lib.ipb_last_error(h).decode() if h else "no context"
            raise DeviceLibraryError(
                "cannot create a CUDA context on device %d: %s "
                "(there is no CPU fallback)" % (device, msg))
        self.device = int(device)
        self._pattern_cache = {}

    @classmethod
    def default(cls):
        if cls._default is None:
            cls._default = cls()
        return cls._default

    # ---- pinned result buffers, recycled when the numpy array is collected
    PIN_MIN = 1 << 15          # doubles; smaller results use ordinary numpy memory

    def empty(self, n):
        """Uninitialised float64 array of length n for a device result: page-locked
        (recycled through a per-context pool) when large."""
        n = int(n)
        if n < self.PIN_MIN:
            return np.empty(n)
        pool = self.__dict__.setdefault("_pin_pool", {})
        free = pool.setdefault(n, [])
        if free:
            ptr = free.pop()
        else:
            p = C.c_void_p()
            self.check(self.lib.ipb_host_alloc(self.h, 8 * n, C.byref(p)))
            ptr = p.value
        arr = np.ctypeslib.as_array((C.c_double * n).from_address(ptr))
        weakref.finalize(arr, self._unpin, ptr, n)
        return arr

    def _unpin(self, ptr, n):
        free = self._pin_pool.setdefault(n, [])
        if len(free) < 4:
            free.append(ptr)
        else:
            self.lib.ipb_host_free(self.h, C.c_void_p(ptr))

    def check(self, st):
        if st == 0:
            return
        msg = self.lib.ipb_last_error(self.h).decode()
        if st in (IPB_ERR_TR_INFEASIBLE, IPB_ERR_NEG_CURVATURE):
            raise ValueError(msg)        # same type as qp_subproblem.py:517,560
        if st == IPB_ERR_INNER_SOLVE:
            raise InnerSolveError(msg)
        raise RuntimeError("ipsolver_b200 error %d: %s" % (st, msg))

    @property
    def launches(self):
        return int(self.lib.ipb_ctx_launches(self.h))

    def grid(self):
        b, t, s = C.c_int(), C.c_int(), C.c_int()
        self.check(self.lib.ipb_ctx_grid(self.h, C.byref(b), C.byref(t), C.byref(s)))
        return b.value, t.value, s.value

    def sync(self):
        self.check(self.lib.ipb_ctx_sync(self.h))

    def trace(self, capacity):
        """Record the time of every grid barrier of the following launches
        (0: off); read with :meth:`trace_read`."""
        self._trace_cap = int(capacity)
        self.check(self.lib.ipb_trace_enable(self.h, int(capacity)))

    def trace_read(self):
        """Barrier times (ns, numpy uint64) of the last launch."""
        cap = getattr(self, "_trace_cap", 0)
        buf = np.zeros(max(cap, 1), dtype=np.uint64)
        cnt = C.c_int()
        self.check(self.lib.ipb_trace_read(
            self.h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), cap, C.byref(cnt)))
        return buf[:min(cnt.value, cap)]

    def timer_start(self):
        self.check(self.lib.ipb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self.check(self.lib.ipb_timer_stop(self.h, C.byref(ms)))
        return ms.value


class DeviceVector:
    def __init__(self, ctx, data=None, length=None):
        self.ctx = ctx
        if data is not None:
            data = f64(data)
            length = data.shape[0]
        self.n = int(length)
        h = C.c_void_p()
        ctx.check(ctx.lib.ipb_vec_create(ctx.h, self.n, C.byref(h)))
        self.h = h
        weakref.finalize(self, ctx.lib.ipb_vec_destroy, ctx.h, h)
        if data is not None:
            self.upload(data)

    def upload(self, data):
        data = f64(data, self.n)
        self.ctx.check(self.ctx.lib.ipb_vec_upload(self.ctx.h, self.h, dptr(data)))

    def download(self, out=None):
        if out is None:
            out = np.empty(self.n)
        self.ctx.check(self.ctx.lib.ipb_vec_download(self.ctx.h, self.h, dptr(out)))
        return out

    @property
    def ptr(self):
        return self.ctx.lib.ipb_vec_ptr(self.h)


class DeviceCSR:
    """A scipy CSR matrix resident in HBM (values fp64, indices int32)."""

    def __init__(self, ctx, M, with_transpose=True, local_columns=False, cols=None):
        """``local_columns``: the column indices are a rank's local numbering
        (own block at 0, ghost entries negative / past the own length; see
        ``distributed.py``) and are taken as they are; ``cols`` is then the
        own length of the column space."""
        M = sps.csr_matrix(M)
        if not local_columns and not M.has_canonical_format:
            M = M.copy()
            M.sum_duplicates()
        self.ctx = ctx
        self.shape = M.shape if cols is None else (M.shape[0], int(cols))
        self.nnz = int(M.nnz)
        self.with_transpose = bool(with_transpose)
        indptr, indices = M.indptr, M.indices
        if indptr.dtype != indices.dtype or indptr.dtype not in (np.int32, np.int64):
            big = indptr.dtype == np.int64 or indices.dtype == np.int64
            dt = np.int64 if big else np.int32
            indptr, indices = indptr.astype(dt), indices.astype(dt)
        indptr = np.ascontiguousarray(indptr)
        indices = np.ascontiguousarray(indices)
        data = f64(M.data)
        h = C.c_void_p()
        ctx.check(ctx.lib.ipb_csr_create(
            ctx.h, self.shape[0], self.shape[1], self.nnz,
            indptr.ctypes.data_as(C.c_void_p), indices.ctypes.data_as(C.c_void_p),
            indptr.dtype.itemsize, dptr(data), int(with_transpose), C.byref(h)))
        self.h = h
        self._pattern = (indptr, indices)
        weakref.finalize(self, ctx.lib.ipb_csr_destroy, ctx.h, h)

    def same_pattern(self, M):
        ip, ix = self._pattern
        return (M.shape == self.shape and M.nnz == self.nnz
                and M.has_canonical_format
                and (M.indptr is ip or np.array_equal(M.indptr, ip))
                and (M.indices is ix or np.array_equal(M.indices, ix)))

    def update_values(self, data):
        data = f64(data, self.nnz)
        self.ctx.check(self.ctx.lib.ipb_csr_update_values(self.ctx.h, self.h, dptr(data)))

    def matvec(self, x):
        x = f64(x, self.shape[1])
        y = np.empty(self.shape[0])
        self.ctx.check(self.ctx.lib.ipb_csr_matvec(self.ctx.h, self.h, dptr(x), dptr(y)))
        return y

    def rmatvec(self, y):
        y = f64(y, self.shape[0])
        x = np.empty(self.shape[1])
        self.ctx.check(self.ctx.lib.ipb_csr_rmatvec(self.ctx.h, self.h, dptr(y), dptr(x)))
        return x

    def fro_norm(self):
        out = C.c_double()
        self.ctx.check(self.ctx.lib.ipb_csr_fro_norm(self.ctx.h, self.h, C.byref(out)))
        return out.value


class DeviceHessian:
    """``H v = [(S + D) v_x ; tail * v_s]`` on the device."""

    def __init__(self, ctx, n, n_x=None, csr=None, dense=None, diag_tail=None):
        self.ctx = ctx
        self.n = int(n)
        self.n_x = self.n if n_x is None else int(n_x)
        h = C.c_void_p()
        ctx.check(ctx.lib.ipb_hess_create(ctx.h, self.n, self.n_x, C.byref(h)))
        self.h = h
        weakref.finalize(self, ctx.lib.ipb_hess_destroy, ctx.h, h)
        self.csr = None
        if csr is not None:
            self.set_csr(csr)
        if dense is not None:
            self.set_dense(dense)
        if diag_tail is not None:
            self.set_diag_tail(diag_tail)

    def set_csr(self, csr):
        if not isinstance(csr, DeviceCSR):
            csr = DeviceCSR(self.ctx, csr, with_transpose=False)
        self.csr = csr                       # keep alive (borrowed by the C side)
        self.ctx.check(self.ctx.lib.ipb_hess_set_csr(self.ctx.h, self.h, csr.h))

    def set_dense(self, dense):
        d = np.ascontiguousarray(dense, dtype=np.float64)
        if d.shape != (self.n_x, self.n_x):
            raise ValueError("dense Hessian block must be n_x by n_x")
        self.ctx.check(self.ctx.lib.ipb_hess_set_dense(self.ctx.h, self.h, dptr(d.reshape(-1))))

    def set_diag_tail(self, diag):
        d = f64(diag, self.n - self.n_x)
        self.ctx.check(self.ctx.lib.ipb_hess_set_diag_tail(self.ctx.h, self.h, dptr(d)))

    def matvec(self, x):
        x = f64(x, self.n)
        y = np.empty(self.n)
        self.ctx.check(self.ctx.lib.ipb_hess_matvec(self.ctx.h, self.h, dptr(x), dptr(y)))
        return y


class DeviceProjector:
    """Z, LS, Y of ``projections.py:290`` with (A A')^-1 applied on the device
    (banded direct solve or CG).  Holds the device copy of A (and A')."""

    # row-partitioned ranks must all take the same "is there a box" decision: a rank whose
    # own block happens to have no finite bound still runs the box passes (distributed.py)
    keep_given_bounds = False

    def __init__(self, ctx, A, orth_tol=1e-12, max_refin=3, inner_rtol=1e-13,
                 inner_maxit=1000, aat_mode=0, band=(0, 0, 0)):
        self.ctx = ctx
        self.A = A if isinstance(A, DeviceCSR) else DeviceCSR(ctx, A, True)
        self.m, self.n = self.A.shape
        band = tuple(int(v) for v in band)
        self.params = (float(orth_tol), int(max_refin), float(inner_rtol),
                       int(inner_maxit), int(aat_mode), band)
        h = C.c_void_p()
        ctx.check(ctx.lib.ipb_proj_create(ctx.h, self.A.h, *self.params[:4], C.byref(h)))
        self.h = h
        weakref.finalize(self, ctx.lib.ipb_proj_destroy, ctx.h, h)
        if aat_mode:
            self.set_aat_mode(aat_mode)
        if band != (0, 0, 0):
            self.set_band(*band)
        self.last_info = None
        self.totals = dict(inner_iters=0, n_solves=0, n_refinements=0, applies=0,
                           n_fallbacks=0)

    def refresh(self):
        self.ctx.check(self.ctx.lib.ipb_proj_refresh(self.ctx.h, self.h))

    def set_aat_mode(self, mode):
        """0: explicit A A' when it moves fewer bytes (default), 1: always
        A then A', 2: always the explicit product."""
        self.ctx.check(self.ctx.lib.ipb_proj_set_aat_mode(self.ctx.h, self.h, int(mode)))

    def set_band(self, mode=0, chunk_rows=0, overlap=0):
        """Banded direct solve of ``(A A')``: mode 0 use it when ``A A'`` is
        banded (default), 1 never (CG only); chunk_rows / overlap 0 = defaults."""
        self.ctx.check(self.ctx.lib.ipb_proj_set_band(
            self.ctx.h, self.h, int(mode), int(chunk_rows), int(overlap)))

    @property
    def band_info(self):
        """dict(in_use, bandwidth, chunk_rows, overlap, groups, trusted)."""
        out = (C.c_int32 * 6)()
        self.ctx.check(self.ctx.lib.ipb_proj_band_info(self.ctx.h, self.h, out))
        return dict(zip(("in_use", "bandwidth", "chunk_rows", "overlap", "groups", "trusted"),
                        [int(v) for v in out]))

    @property
    def aat_nnz(self):
        """nnz of the explicit A A' held in HBM, or -1 when it is not held."""
        return int(self.ctx.lib.ipb_proj_aat_nnz(self.h))

    def _note(self, info):
        self.last_info = info
        self.totals["inner_iters"] += info.inner_iters
        self.totals["n_solves"] += info.n_solves
        self.totals["n_refinements"] += info.n_refinements
        self.totals["n_fallbacks"] += info.n_fallbacks
        self.totals["applies"] += 1

    def null_space(self, x):
        x = f64(x, self.n)
        z = self.ctx.empty(self.n)
        info = ProjInfo()
        self.ctx.check(self.ctx.lib.ipb_proj_null_space(
            self.ctx.h, self.h, dptr(x), dptr(z), C.byref(info)))
        self._note(info)
        return z

    def least_squares(self, x):
        x = f64(x, self.n)
        v = self.ctx.empty(self.m)
        info = ProjInfo()
        self.ctx.check(self.ctx.lib.ipb_proj_least_squares(
            self.ctx.h, self.h, dptr(x), dptr(v), C.byref(info)))
        self._note(info)
        return v

    def row_space(self, b):
        b = f64(b, self.m)
        y = self.ctx.empty(self.n)
        info = ProjInfo()
        self.ctx.check(self.ctx.lib.ipb_proj_row_space(
            self.ctx.h, self.h, dptr(b), dptr(y), C.byref(info)))
        self._note(info)
        return y

    def orthogonality(self, g):
        g = f64(g, self.n)
        out = C.c_double()
        self.ctx.check(self.ctx.lib.ipb_orthogonality(self.ctx.h, self.h, dptr(g), C.byref(out)))
        return out.value

    def modified_dogleg(self, b, trust_radius, lb=None, ub=None):
        b = f64(b, self.m)
        lb, ub = _bounds(lb, ub, self.n)
        x = self.ctx.empty(self.n)
        info = ProjInfo()
        self.ctx.check(self.ctx.lib.ipb_modified_dogleg(
            self.ctx.h, self.h, dptr(b), float(trust_radius), dptr(lb), dptr(ub),
            dptr(x), C.byref(info)))
        self._note(info)
        return x

    def sqp_normal_step(self, c, b, normal_radius, lb=None, ub=None, bounds_changed=True,
                        box_factor=0.5):
        """``dn = modified_dogleg(...)`` with ``c`` / ``b`` / bounds kept resident
        for :meth:`sqp_tangential_step` (``None``: unchanged since the last call)."""
        c = None if c is None else f64(c, self.n)
        b = None if b is None else f64(b, self.m)
        if self.keep_given_bounds and lb is not None and ub is not None:
            lb, ub = f64(lb, self.n), f64(ub, self.n)     # the box decision is the caller's (global)
        else:
            lb, ub = _bounds(lb, ub, self.n)
        dn = self.ctx.empty(self.n)
        info = ProjInfo()
        self.ctx.check(self.ctx.lib.ipb_sqp_normal_step(
            self.ctx.h, self.h, dptr(c), dptr(b), float(normal_radius), dptr(lb), dptr(ub),
            int(bool(bounds_changed)), float(box_factor), dptr(dn), C.byref(info)))
        self._note(info)
        return dn

    def sqp_tangential_step(self, H, tangential_radius):
        """``d = dn + projected_cg(H, H dn + c, ..., radius_t, lb - dn, ub - dn)``
        on the vectors the normal step left in HBM; returns ``d, H d, A d, info``."""
        d, Hd = self.ctx.empty(self.n), self.ctx.empty(self.n)
        Ad = self.ctx.empty(self.m)
        info = PcgInfo()
        st = self.ctx.lib.ipb_sqp_tangential_step(
            self.ctx.h, H.h, self.h, float(tangential_radius), dptr(d), dptr(Hd),
            dptr(Ad) if self.m else None, C.byref(info))
        self.last_pcg_info = info
        self.ctx.check(st)
        self.totals["inner_iters"] += info.inner_iters
        self.totals["n_solves"] += info.n_solves
        self.totals["n_refinements"] += info.n_refinements
        out = {'niter': int(info.niter), 'stop_cond': int(info.stop_cond),
               'hits_boundary': bool(info.hits_boundary)}
        return d, Hd, Ad, out

    def sqp_multipliers(self, c):
        """``v = -LS c`` and ``||c + A'v||_inf``; ``c`` stays resident for the
        following :meth:`sqp_step` calls."""
        c = f64(c, self.n)
        v = self.ctx.empty(self.m)
        opt = C.c_double()
        info = ProjInfo()
        self.ctx.check(self.ctx.lib.ipb_sqp_multipliers(
            self.ctx.h, self.h, dptr(c), dptr(v), C.byref(opt), C.byref(info)))
        if self.m:
            self._note(info)
        return v, opt.value

    def projected_cg(self, H, c, b, trust_radius=np.inf, lb=None, ub=None,
                     tol=None, max_iter=None, max_infeasible_iter=None,
                     return_all=False):
        c = f64(c, self.n)
        b = f64(b, self.m)
        lb, ub = _bounds(lb, ub, self.n)
        x = self.ctx.empty(self.n)
        info = PcgInfo()
        cap = 0
        allvecs = None
        if return_all:
            nm = max(self.n - self.m, 0)
            cap = (nm if max_iter is None else min(max_iter, nm)) + 2
            if cap * self.n * 8 > (1 << 32):
                raise MemoryError("return_all would need %.1f GB" % (cap * self.n * 8 / 1e9))
            allvecs = np.empty(cap * self.n)
        st = self.ctx.lib.ipb_projected_cg(
            self.ctx.h, H.h, self.h, dptr(c), dptr(b), float(trust_radius),
            dptr(lb), dptr(ub), -1.0 if tol is None else float(tol),
            -1 if max_iter is None else int(max_iter),
            -1 if max_infeasible_iter is None else int(max_infeasible_iter),
            dptr(x), dptr(allvecs), cap, C.byref(info))
        self.last_pcg_info = info
        self.ctx.check(st)
        self.totals["inner_iters"] += info.inner_iters
        self.totals["n_solves"] += info.n_solves
        self.totals["n_refinements"] += info.n_refinements
        out = {'niter': int(info.niter), 'stop_cond': int(info.stop_cond),
               'hits_boundary': bool(info.hits_boundary)}
        if return_all:
            k = min(info.n_allvecs, cap)
            out['allvecs'] = [allvecs[i*self.n:(i+1)*self.n].copy() for i in range(k)]
        return x, out


def projected_cg_dev(proj, H, c, b, x, trust_radius=np.inf, lb=None, ub=None,
                     tol=None, max_iter=None, max_infeasible_iter=None):
    """Device-resident variant: c, b, lb, ub, x are DeviceVectors; only the
    info struct crosses PCIe."""
    ctx = proj.ctx
    info = PcgInfo()
    st = ctx.lib.ipb_projected_cg_dev(
        ctx.h, H.h, proj.h, c.h, b.h, float(trust_radius),
        lb.h if lb is not None else None, ub.h if ub is not None else None,
        -1.0 if tol is None else float(tol),
        -1 if max_iter is None else int(max_iter),
        -1 if max_infeasible_iter is None else int(max_infeasible_iter),
        x.h, C.byref(info))
    proj.last_pcg_info = info
    ctx.check(st)
    return info


def bench_inner_iterations(proj, x, iters):
    ms = C.c_double()
    proj.ctx.check(proj.ctx.lib.ipb_bench_inner_iterations(
        proj.ctx.h, proj.h, x.h, int(iters), C.byref(ms)))
    return ms.value


def bench_phase(proj, phase, reps):
    ms = C.c_double()
    proj.ctx.check(proj.ctx.lib.ipb_bench_phase(
        proj.ctx.h, proj.h, int(phase), int(reps), C.byref(ms)))
    return ms.value


def _bounds(lb, ub, n):
    """None when no bound is finite (the kernels then skip the box passes)."""
    if lb is None and ub is None:
        return None, None
    lb = np.full(n, -np.inf) if lb is None else f64(lb, n)
    ub = np.full(n, np.inf) if ub is None else f64(ub, n)
    if (lb == -np.inf).all() and (ub == np.inf).all():
        return None, None
    return lb, ub
