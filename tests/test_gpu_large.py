"""BASELINE.json configs at their stated sizes on the GPU, against records the
unmodified reference produced in the build container
(``tests/golden/make_golden_large.py`` -> ``e2e_large.npz``: scalars, norms and
a strided sample of x and of the multipliers).

Config 2 (convex QP, n = 1e6, equality-constrained SQP) is the parity pin
north_star states: iteration counts +-1, objective / optimality / solution to
1e-8 relative.  The reference needs 70 s for it, the device path about 1 s.
Configs 3 and 4 are compared at the largest sizes the reference finishes in
reasonable time (interior point n = 1e5: 24 min; Thomson N = 2000: 5 min);
these runs take > 100 outer iterations on a non-convex or barrier problem and
end by ``xtol``, so the solution is compared, not the path."""
import os

import numpy as np
import pytest

import e2e_problems

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import ipsolver_b200 as ipb
    return dict(minimize_constrained=ipb.minimize_constrained,
                LinearConstraint=ipb.LinearConstraint, BoxConstraint=ipb.BoxConstraint,
                NonlinearConstraint=ipb.NonlinearConstraint)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "e2e_large.npz"))


def _sample_err(res, gold, name):
    s = e2e_problems.LARGE_STRIDE
    xs, vs = gold[name + "/xs"], gold[name + "/vs"]
    nx, _, nv, _, _, _ = gold[name + "/norms"]
    ex = np.linalg.norm(np.asarray(res.x)[::s] - xs) / max(np.linalg.norm(xs), 1e-300)
    ev = np.linalg.norm(np.asarray(res.v)[::s] - vs) / max(np.linalg.norm(vs), 1e-300)
    return ex, ev, abs(np.linalg.norm(res.x) - nx) / nx, abs(np.linalg.norm(res.v) - nv) / nv


def test_config2_sqp_n1e6_matches_the_reference(api, gold, golden_host_threads):
    name = "eqp_window_1000000"
    niter, cg, status, nfev, opt, viol, fval, radius, t_ref = gold[name + "/scalars"]
    res = e2e_problems.LARGE_PROBLEMS[name](api)()
    print("config 2: device niter %d cg %d nfev %d status %d opt %.6e fun %.12e in %.2f s | "
          "reference niter %d cg %d nfev %d status %d opt %.6e fun %.12e in %.1f s"
          % (res.niter, res.cg_niter, res.nfev, res.status, res.optimality, res.fun,
             res.execution_time, niter, cg, nfev, status, opt, fval, t_ref))
    assert res.status == status
    assert abs(res.niter - niter) <= 1 and abs(res.cg_niter - cg) <= 1 and abs(res.nfev - nfev) <= 1
    assert res.fun == pytest.approx(fval, rel=1e-8)
    # optimality is ~1e-8 of the gradient scale at the solution: compared on that scale
    assert abs(res.optimality - opt) <= 1e-8 * max(1.0, np.abs(res.grad).max())
    ex, ev, enx, env = _sample_err(res, gold, name)
    assert ex < 1e-8 and ev < 1e-8 and enx < 1e-8 and env < 1e-8
    assert res.constr_violation < 1e-12


@pytest.mark.parametrize("name", ["ipqp_window_100000"])
def test_config3_interior_point_reaches_the_reference_point(api, gold, name):
    niter, cg, status, nfev, opt, viol, fval, radius, t_ref = gold[name + "/scalars"]
    res = e2e_problems.LARGE_PROBLEMS[name](api)()
    print("config 3 (%s): device niter %d cg %d status %d opt %.3e fun %.10e in %.1f s | reference "
          "niter %d cg %d status %d opt %.3e fun %.10e in %.0f s"
          % (name, res.niter, res.cg_niter, res.status, res.optimality, res.fun,
             res.execution_time, niter, cg, status, opt, fval, t_ref))
    assert res.status in (1, 2)
    assert res.constr_violation < 1e-8
    assert res.fun == pytest.approx(fval, rel=1e-6)
    ex, ev, enx, env = _sample_err(res, gold, name)
    assert ex < 1e-4 and enx < 1e-5
    assert abs(res.niter - niter) <= 0.25 * niter


@pytest.mark.parametrize("name", ["thomson_500_equals", "thomson_2000_equals"])
def test_config4_thomson_reaches_the_reference_energy(api, gold, name):
    """Non-convex, rotation-invariant: local minima differ in the 6th digit of
    the energy; the reference itself stops by xtol at optimality ~1e-5."""
    niter, cg, status, nfev, opt, viol, fval, radius, t_ref = gold[name + "/scalars"]
    res = e2e_problems.LARGE_PROBLEMS[name](api)()
    print("config 4 (%s): device niter %d cg %d status %d opt %.3e fun %.10e in %.1f s | reference "
          "niter %d cg %d status %d opt %.3e fun %.10e in %.0f s"
          % (name, res.niter, res.cg_niter, res.status, res.optimality, res.fun,
             res.execution_time, niter, cg, status, opt, fval, t_ref))
    assert res.status in (1, 2)
    assert res.constr_violation < 1e-8
    assert res.fun == pytest.approx(fval, rel=2e-5)      # neighbouring local minima of the energy
    assert res.optimality < 1e-3
