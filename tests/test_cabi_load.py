"""CPU-side checks of the C-ABI boundary: the library loads, exports every
symbol declared in include/ipsolver_b200.h, the ctypes prototypes cover the
same set, and the host mirrors validate arguments like the reference.  No
compute call is made (there is no GPU on the build box)."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ipsolver_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ipb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_functions():
    syms = declared_symbols()
    assert "ipb_projected_cg" in syms and "ipb_proj_null_space" in syms
    assert len(syms) >= 40


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    entry.build()
    from ipsolver_b200 import _cabi
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), "missing export: " + name


def test_prototypes_cover_header():
    from ipsolver_b200 import _cabi
    assert sorted(_cabi.PROTOTYPES) == declared_symbols()


def test_info_struct_layout_matches_header():
    from ipsolver_b200 import _cabi
    assert ctypes.sizeof(_cabi.PcgInfo) == 8 + 4 * 4 + 4 * 8 + 3 * 8 + 8 + 2 * 4
    assert ctypes.sizeof(_cabi.ProjInfo) == 3 * 8 + 2 * 4 + 2 * 8


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from ipsolver_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.DeviceLibraryError):
        _cabi.load()


def test_no_gpu_fails_loudly_not_silently():
    """Without a CUDA device the product raises; it never computes on the CPU."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from ipsolver_b200 import device, projections
    device.Context._default = None
    with pytest.raises(device.DeviceLibraryError):
        projections.projections(sps.csr_matrix(np.eye(2, 4)))


def test_method_validation_without_device():
    """Reference projections.py:372-387: wrong family raises ValueError before
    anything touches the device."""
    from ipsolver_b200.projections import projections
    A = np.eye(3, 8)
    with pytest.raises(ValueError):
        projections(A, "AugmentedSystem")
    with pytest.raises(ValueError):
        projections(sps.csr_matrix(A), "QRFactorization")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ip-nonlinear-solver_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
