"""ctypes binding of ``libipsolver_b200.so`` (C ABI: ``include/ipsolver_b200.h``).

This is the only module that touches the shared library.  There is no CPU
fallback: if the library is missing or no CUDA device is usable, importing
the device layer raises and every product entry point fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# IPB_LIB_PATH selects another build of the same library (tuning variants
# compiled with different block-shape constants, scripts/build_variants.sh)
LIB_PATH = os.environ.get("IPB_LIB_PATH") or os.path.join(_HERE, "libipsolver_b200.so")

IPB_OK = 0
IPB_ERR_INVALID = 1
IPB_ERR_CUDA = 2
IPB_ERR_TR_INFEASIBLE = 3
IPB_ERR_NEG_CURVATURE = 4
IPB_ERR_INNER_SOLVE = 5
IPB_ERR_NOMEM = 6
IPB_ERR_COMM = 7


class PcgInfo(C.Structure):
    _fields_ = [("niter", C.c_int64), ("stop_cond", C.c_int32),
                ("hits_boundary", C.c_int32), ("status", C.c_int32),
                ("n_allvecs", C.c_int32), ("inner_iters", C.c_int64),
                ("n_solves", C.c_int64), ("n_refinements", C.c_int64),
                ("n_spmv_h", C.c_int64), ("max_orth", C.c_double),
                ("tol_used", C.c_double), ("kernel_ms", C.c_double),
                ("n_sweeps", C.c_int64), ("n_fallbacks", C.c_int32),
                ("reserved", C.c_int32)]


class ProjInfo(C.Structure):
    _fields_ = [("inner_iters", C.c_int64), ("n_solves", C.c_int64),
                ("n_refinements", C.c_int64), ("status", C.c_int32),
                ("n_fallbacks", C.c_int32), ("orth", C.c_double),
                ("kernel_ms", C.c_double)]


_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I64 = C.c_int64

# name -> (restype, argtypes); must list every symbol of include/ipsolver_b200.h
PROTOTYPES = {
    "ipb_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "ipb_ctx_destroy": (C.c_int, [_P]),
    "ipb_last_error": (C.c_char_p, [_P]),
    "ipb_ctx_grid": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int),
                               C.POINTER(C.c_int)]),
    "ipb_ctx_launches": (_I64, [_P]),
    "ipb_ctx_sync": (C.c_int, [_P]),
    "ipb_csr_create": (C.c_int, [_P, _I64, _I64, _I64, _P, _P, C.c_int, _D,
                                 C.c_int, C.POINTER(_P)]),
    "ipb_csr_update_values": (C.c_int, [_P, _P, _D]),
    "ipb_csr_destroy": (C.c_int, [_P, _P]),
    "ipb_csr_matvec": (C.c_int, [_P, _P, _D, _D]),
    "ipb_csr_rmatvec": (C.c_int, [_P, _P, _D, _D]),
    "ipb_csr_matvec_dev": (C.c_int, [_P, _P, _P, _P]),
    "ipb_csr_rmatvec_dev": (C.c_int, [_P, _P, _P, _P]),
    "ipb_csr_fro_norm": (C.c_int, [_P, _P, _D]),
    "ipb_hess_create": (C.c_int, [_P, _I64, _I64, C.POINTER(_P)]),
    "ipb_hess_set_csr": (C.c_int, [_P, _P, _P]),
    "ipb_hess_set_dense": (C.c_int, [_P, _P, _D]),
    "ipb_hess_set_diag_tail": (C.c_int, [_P, _P, _D]),
    "ipb_hess_destroy": (C.c_int, [_P, _P]),
    "ipb_hess_matvec": (C.c_int, [_P, _P, _D, _D]),
    "ipb_hess_matvec_dev": (C.c_int, [_P, _P, _P, _P]),
    "ipb_proj_create": (C.c_int, [_P, _P, C.c_double, C.c_int, C.c_double,
                                  C.c_int, C.POINTER(_P)]),
    "ipb_proj_refresh": (C.c_int, [_P, _P]),
    "ipb_proj_set_aat_mode": (C.c_int, [_P, _P, C.c_int]),
    "ipb_proj_aat_nnz": (_I64, [_P]),
    "ipb_proj_set_band": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int]),
    "ipb_proj_band_info": (C.c_int, [_P, _P, C.POINTER(C.c_int32)]),
    "ipb_bench_phase": (C.c_int, [_P, _P, C.c_int, C.c_int, _D]),
    "ipb_proj_destroy": (C.c_int, [_P, _P]),
    "ipb_proj_null_space": (C.c_int, [_P, _P, _D, _D, C.POINTER(ProjInfo)]),
    "ipb_proj_least_squares": (C.c_int, [_P, _P, _D, _D, C.POINTER(ProjInfo)]),
    "ipb_proj_row_space": (C.c_int, [_P, _P, _D, _D, C.POINTER(ProjInfo)]),
    "ipb_proj_null_space_dev": (C.c_int, [_P, _P, _P, _P, C.POINTER(ProjInfo)]),
    "ipb_proj_least_squares_dev": (C.c_int, [_P, _P, _P, _P, C.POINTER(ProjInfo)]),
    "ipb_proj_row_space_dev": (C.c_int, [_P, _P, _P, _P, C.POINTER(ProjInfo)]),
    "ipb_orthogonality": (C.c_int, [_P, _P, _D, _D]),
    "ipb_projected_cg": (C.c_int, [_P, _P, _P, _D, _D, C.c_double, _D, _D,
                                   C.c_double, _I64, _I64, _D, _D, _I64,
                                   C.POINTER(PcgInfo)]),
    "ipb_projected_cg_dev": (C.c_int, [_P, _P, _P, _P, _P, C.c_double, _P, _P,
                                       C.c_double, _I64, _I64, _P,
                                       C.POINTER(PcgInfo)]),
    "ipb_modified_dogleg": (C.c_int, [_P, _P, _D, C.c_double, _D, _D, _D,
                                      C.POINTER(ProjInfo)]),
    "ipb_sqp_normal_step": (C.c_int, [_P, _P, _D, _D, C.c_double, _D, _D, C.c_int, C.c_double,
                                      _D, C.POINTER(ProjInfo)]),
    "ipb_sqp_tangential_step": (C.c_int, [_P, _P, _P, C.c_double, _D, _D, _D,
                                          C.POINTER(PcgInfo)]),
    "ipb_sqp_multipliers": (C.c_int, [_P, _P, _D, _D, _D, C.POINTER(ProjInfo)]),
    "ipb_box_sphere_intersections": (C.c_int, [_P, _I64, _D, _D, _D, _D,
                                               C.c_double, C.c_int, _D, _D,
                                               C.POINTER(C.c_int)]),
    "ipb_box_sphere_intersections_dev": (C.c_int, [_P, _P, _P, _P, _P, C.c_double, C.c_int, _D, _D,
                                                   C.POINTER(C.c_int)]),
    "ipb_comm_init": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_I64), C.POINTER(_I64),
                                C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64),
                                C.POINTER(_I64), _I64, _I64, _P]),
    "ipb_comm_connect": (C.c_int, [_P, _P]),
    "ipb_comm_set_send_list": (C.c_int, [_P, C.c_int, _I64, _P]),
    "ipb_csr_set_space": (C.c_int, [_P, _P, C.c_int]),
    "ipb_proj_create_parts": (C.c_int, [_P, _P, _P, _P, _P, _P, _D, C.c_double, C.c_double,
                                        C.c_int, C.c_double, C.c_int, C.POINTER(_P)]),
    "ipb_proj_set_band_parts": (C.c_int, [_P, _P, _P, _I64, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ipb_vec_create": (C.c_int, [_P, _I64, C.POINTER(_P)]),
    "ipb_vec_destroy": (C.c_int, [_P, _P]),
    "ipb_vec_upload": (C.c_int, [_P, _P, _D]),
    "ipb_vec_download": (C.c_int, [_P, _P, _D]),
    "ipb_vec_len": (_I64, [_P]),
    "ipb_vec_ptr": (_P, [_P]),
    "ipb_host_alloc": (C.c_int, [_P, _I64, C.POINTER(_P)]),
    "ipb_host_free": (C.c_int, [_P, _P]),
    "ipb_trace_enable": (C.c_int, [_P, C.c_int]),
    "ipb_trace_read": (C.c_int, [_P, C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_int)]),
    "ipb_timer_start": (C.c_int, [_P]),
    "ipb_timer_stop": (C.c_int, [_P, _D]),
    "ipb_bench_inner_iterations": (C.c_int, [_P, _P, _P, C.c_int, _D]),
}

_lib = None


class DeviceLibraryError(RuntimeError):
    """The CUDA library is missing, or no usable B200 is visible."""


def load():
    """Load the shared library and attach prototypes (no CUDA call yet)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DeviceLibraryError(
            "libipsolver_b200.so not found at %s -- build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError if a symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dptr(a):
    """``double*`` of a C-contiguous float64 array (or NULL for None)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_D)


def f64(a, n=None):
    out = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    if n is not None and out.shape[0] != n:
        raise ValueError("expected a vector of length %d, got %d"
                         % (n, out.shape[0]))
    return out
