"""Pin the CPU oracle against vectors produced by the reference itself
(tests/golden/make_golden.py) and against the reference tests' known answers.

CPU only.  The oracle restates the reference's numpy call sequence, so the
comparison is bit-for-bit where the same LAPACK/SuperLU build is used; the
assertions allow 1e-13 relative so the suite also passes on a host whose BLAS
differs from the build container's.
"""
import json
import os

import numpy as np
import pytest
import scipy.linalg
import scipy.sparse as sps

import cases
import oracle
from ipsolver_b200 import synthetic

RTOL = 1e-13


@pytest.fixture(scope="module")
def kat(golden_dir):
    with open(os.path.join(golden_dir, "kat.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def synth(golden_dir):
    return np.load(os.path.join(golden_dir, "synth_pcg.npz"))


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = max(1.0, float(np.max(np.abs(b))) if b.size else 1.0)
    np.testing.assert_allclose(a, b, rtol=0, atol=rtol * scale)


@pytest.mark.parametrize("name", list(cases.PCG_CASES))
def test_pcg_known_answers(kat, name):
    H, A, c, b, kw, exc = cases.PCG_CASES[name]
    Hs, As = sps.csc_matrix(H), sps.csc_matrix(A)
    Z, _, Y = oracle.build_projectors(As)
    want = kat["pcg"][name]
    if exc is not None:
        with pytest.raises(exc) as ei:
            oracle.eqp_projected_cg(Hs, np.array(c), Z, Y, np.array(b), **kw)
        assert want["raises"] == type(ei.value).__name__
        return
    x, info = oracle.eqp_projected_cg(Hs, np.array(c), Z, Y, np.array(b), **kw)
    close(x, want["x"])
    assert info["niter"] == want["niter"]
    assert info["stop_cond"] == want["stop_cond"]
    assert info["hits_boundary"] == want["hits_boundary"]
    if "allvecs" in want:
        assert len(info["allvecs"]) == len(want["allvecs"])
        for got, ref in zip(info["allvecs"], want["allvecs"]):
            close(got, ref)


def test_pcg_textbook_answers():
    """Nocedal-Wright ex. 16.2 and the direct KKT comparison
    (reference tests/test_qp_subproblem.py:429-458)."""
    H, A, c, b, kw, _ = cases.PCG_CASES["nocedal_16_2"]
    Z, _, Y = oracle.build_projectors(sps.csc_matrix(A))
    x, info = oracle.eqp_projected_cg(sps.csc_matrix(H), np.array(c), Z, Y,
                                      np.array(b))
    assert info["stop_cond"] == 4 and not info["hits_boundary"]
    np.testing.assert_array_almost_equal(x, [2, -1, 1])
    H, A, c, b, kw, _ = cases.PCG_CASES["direct_fact"]
    Hs, As = sps.csc_matrix(H), sps.csc_matrix(A)
    Z, _, Y = oracle.build_projectors(As)
    x, info = oracle.eqp_projected_cg(Hs, np.array(c), Z, Y, np.array(b), tol=0)
    x_kkt, _ = oracle.kkt_direct_solve(Hs, np.array(c), As, np.array(b))
    np.testing.assert_array_almost_equal(x, x_kkt)


@pytest.mark.parametrize("name", list(cases.DOGLEG_CASES))
def test_dogleg(kat, name):
    A, b, radius, lb, ub = cases.DOGLEG_CASES[name]
    A = np.array(A, float)
    _, _, Y = oracle.build_projectors(A)
    x = oracle.dogleg_normal_step(A, Y, np.array(b), radius,
                                  np.array(lb, float), np.array(ub, float))
    close(x, kat["dogleg"][name])


def test_intersections(kat):
    for (z, d, radius, line), want in zip(cases.BALL_CASES, kat["ball"]):
        ta, tb, hit = oracle.segment_ball(np.array(z, float),
                                          np.array(d, float), radius, line)
        assert [float(ta), float(tb), bool(hit)] == want
    k = 0
    for (z, d, lb, ub, line), want in zip(cases.BOX_CASES, kat["box"]):
        args = [np.array(v, float) for v in (z, d, lb, ub)]
        ta, tb, hit = oracle.segment_box(*args, entire_line=line)
        assert [float(ta), float(tb), bool(hit)] == want
        for radius in (2.5, np.inf):
            ta, tb, hit = oracle.segment_box_ball(*args, radius, line)
            assert [float(ta), float(tb), bool(hit)] == kat["box_ball"][k]
            k += 1


@pytest.mark.parametrize("method", ["AugmentedSystem", "NormalEquation"])
def test_projectors_3x8(kat, method):
    """Reference tests/test_projections.py:25-85 properties + golden values.
    NormalEquation (CHOLMOD in the reference, SuperLU on AA' here) is pinned
    against the AugmentedSystem golden values to 1e-12."""
    A_dense = np.array(cases.A38, float)
    A = sps.csc_matrix(A_dense)
    Z, LS, Y = oracle.build_projectors(A, method)
    rtol = RTOL if method == "AugmentedSystem" else 1e-12
    for p, zref, lsref in zip(cases.P38_N, kat["proj38"]["Z"],
                              kat["proj38"]["LS"]):
        p = np.array(p, float)
        z = Z.matvec(p)
        close(z, zref, rtol)
        np.testing.assert_array_almost_equal(A.dot(z), 0)
        np.testing.assert_array_almost_equal(oracle.orthogonality_measure(A, z), 0)
        ls = LS.matvec(p)
        close(ls, lsref, rtol)
        np.testing.assert_array_almost_equal(
            ls, scipy.linalg.lstsq(A_dense.T, p)[0])
    for p, yref in zip(cases.P38_M, kat["proj38"]["Y"]):
        y = Y.matvec(np.array(p, float))
        close(y, yref, rtol)
        np.testing.assert_array_almost_equal(A.dot(y), p)
    for p, oref in zip(cases.P38_N, kat["proj38"]["orth"]):
        assert float(oracle.orthogonality_measure(A, np.array(p, float))) == \
            pytest.approx(oref, rel=1e-14)


def test_dense_projectors_agree_with_sparse():
    """Reference tests/test_projections.py:109-124."""
    D = np.diag(range(1, 101))
    A = np.hstack([D, D, D, D]).astype(float)
    rng = np.random.RandomState(0)
    Zd, LSd, Yd = oracle.build_projectors(A)
    Zs, LSs, Ys = oracle.build_projectors(sps.csc_matrix(A))
    for method in ("SVDFactorization",):
        Zv, LSv, Yv = oracle.build_projectors(A, method)
    for _ in range(5):
        z = rng.normal(size=400)
        np.testing.assert_array_almost_equal(Zd.dot(z), Zs.dot(z))
        np.testing.assert_array_almost_equal(LSd.dot(z), LSs.dot(z))
        np.testing.assert_array_almost_equal(Zv.dot(z), Zs.dot(z))
        y = rng.normal(size=100)
        np.testing.assert_array_almost_equal(Yd.dot(y), Ys.dot(y))
        np.testing.assert_array_almost_equal(Yv.dot(y), Ys.dot(y))


def test_method_validation():
    A = np.array(cases.A38, float)
    with pytest.raises(ValueError):
        oracle.build_projectors(A, "AugmentedSystem")
    with pytest.raises(ValueError):
        oracle.build_projectors(sps.csr_matrix(A), "QRFactorization")


def test_orthogonality_known_vectors():
    """Reference tests/test_projections.py:180-217."""
    A = np.array(cases.A38, float)
    test_vectors = ([-1.98931144, -1.56363389, -0.84115584, 2.2864762,
                     5.599141, 0.09286976, 1.37040802, -0.28145812],
                    [697.92794044, -4091.65114008, -3327.42316335,
                     836.86906951, 99434.98929065, -1285.37653682,
                     -4109.21503806, 2935.29289083])
    expected = (0, 0)
    for z, want in zip(test_vectors, expected):
        np.testing.assert_array_almost_equal(
            oracle.orthogonality_measure(A, np.array(z)), want)
        np.testing.assert_array_almost_equal(
            oracle.orthogonality_measure(sps.csc_matrix(A), np.array(z)), want)


@pytest.mark.parametrize("name", list(cases.SYNTH_RUNS))
def test_synthetic_pcg(synth, name):
    pattern, kappa, kw, tau = cases.SYNTH_RUNS[name]
    n = cases.SYNTH_N
    H, A, c, b = synthetic.make_qp(n, pattern, kappa, seed=0)
    Z, LS, Y = oracle.build_projectors(A)
    kw = dict(kw)
    if tau is not None:
        kw["lb"], kw["ub"] = synthetic.box_bounds(n, tau)
    x, info = oracle.eqp_projected_cg(H, c, Z, Y, b, **kw)
    want = synth[name + "/info"]
    assert [info["niter"], info["stop_cond"], int(info["hits_boundary"])] == \
        list(want)
    close(x, synth[name + "/x"], 1e-12)


@pytest.mark.parametrize("pattern", ["window", "rand"])
def test_synthetic_projectors_and_dogleg(synth, pattern):
    n = cases.SYNTH_N
    H, A, c, b = synthetic.make_qp(n, pattern, 1e4, seed=0)
    Z, LS, Y = oracle.build_projectors(A)
    close(Z.dot(c), synth[pattern + "/Zc"])
    close(LS.dot(c), synth[pattern + "/LSc"])
    close(Y.dot(b), synth[pattern + "/Yb"])
    lb, ub = synthetic.box_bounds(n, 0.3)
    for radius in (4.0, 40.0):
        close(oracle.dogleg_normal_step(A, Y, b, radius, lb, ub),
              synth[pattern + "/dogleg_r%g" % radius])
    close(oracle.dogleg_normal_step(A, Y, b, 6.0, np.full(n, -np.inf),
                                    np.full(n, np.inf)),
          synth[pattern + "/dogleg_free"])
    # NormalEquation variant against the AugmentedSystem golden values
    Zn, LSn, Yn = oracle.build_projectors(A, "NormalEquation")
    close(Zn.dot(c), synth[pattern + "/Zc"], 1e-11)
    close(LSn.dot(c), synth[pattern + "/LSc"], 1e-11)
    close(Yn.dot(b), synth[pattern + "/Yb"], 1e-11)
