"""Host logic of the outer loops (``ipsolver_b200.equality_constrained_sqp``,
``tr_interior_point``, ``minimize_constrained``) on the CPU: driven with the
ORACLE's hot-path routines (test-only ``hot_path`` injection) they must
reproduce the reference's end-to-end results -- the golden vectors in
``tests/golden/e2e.npz`` were produced by the unmodified reference
(``tests/golden/make_golden.py``).  Iteration counts must match exactly here:
the oracle is bit-identical to the reference's linear algebra, so any
difference would be in the outer-loop restatement."""
import numpy as np
import pytest

import oracle
import e2e_problems
import ipsolver_b200 as ipb


ORACLE_PATH = oracle.host_hot_path()


def run(name):
    api = dict(minimize_constrained=lambda *a, **k: ipb.minimize_constrained(
                   *a, hot_path=ORACLE_PATH, **k),
               LinearConstraint=ipb.LinearConstraint, BoxConstraint=ipb.BoxConstraint,
               NonlinearConstraint=ipb.NonlinearConstraint)
    return e2e_problems.PROBLEMS[name](api)()


@pytest.mark.parametrize("name", ["eqp_window_2000", "eqp_rand_2000", "eqp_window_2000_gtol6",
                                  "ipqp_window_1000", "thomson_20_equals", "thomson_20_less",
                                  "maratos", "hyperbolic_ineq", "rosenbrock_unconstrained",
                                  "rosenbrock_ineq", "rosenbrock_eq_ineq",
                                  "rosenbrock_box_enforced"])
def test_outer_loop_reproduces_reference(golden_dir, name):
    import os
    gold = np.load(os.path.join(golden_dir, "e2e.npz"))
    niter, cg, status, nfev, opt, viol, fval, radius = gold[name + "/scalars"]
    res = run(name)
    assert (res.niter, res.cg_niter, res.status, res.nfev) == (niter, cg, status, nfev)
    np.testing.assert_allclose(res.x, gold[name + "/x"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(res.v, gold[name + "/v"], rtol=1e-9, atol=1e-12)
    assert res.fun == pytest.approx(fval, rel=1e-13)
    assert res.trust_radius == pytest.approx(radius, rel=1e-12)


def test_kind_vocabulary_and_row_order():
    from ipsolver_b200.minimize_constrained import _Rows
    x0 = np.array([1.0, 2.0, 3.0, 4.0])
    r = _Rows(lambda x: x, lambda x: np.eye(4), None,
              ("interval", [0, -np.inf, 3, -1], [2, 5, 3, np.inf]), False, x0, None)
    assert list(r.eq) == [2] and list(r.val_eq) == [3.0]
    # lower-bound rows first, then upper-bound rows (reference _parse_constraint)
    assert list(r.ineq) == [0, 3, 0, 1]
    assert list(r.sign) == [-1, -1, 1, 1]
    c_in, c_eq = r.values(x0)
    assert list(c_in) == [-1.0, -5.0, -1.0, -3.0] and list(c_eq) == [0.0]
    with pytest.raises(ValueError):
        _Rows(lambda x: x, lambda x: np.eye(4), None, ("between", 0, 1), False, x0, None)
    with pytest.raises(ValueError):
        _Rows(lambda x: x, lambda x: np.eye(4), None, ("interval", 0), False, x0, None)


def test_fd_hessian_is_rejected():
    with pytest.raises(TypeError):
        ipb.minimize_constrained(lambda x: x.dot(x), np.ones(3), lambda x: 2 * x, '2-point',
                                 hot_path=ORACLE_PATH)


def test_operator_hessian_is_accepted():
    """A Hessian given as a LinearOperator (reference _minimize_constrained.py:
    404-406) is an opaque host callback: the outer loops take it as it is."""
    from scipy.sparse.linalg import aslinearoperator
    res = ipb.minimize_constrained(lambda x: x.dot(x), np.ones(3), lambda x: 2 * x,
                                   lambda x: aslinearoperator(2 * np.eye(3)), hot_path=ORACLE_PATH)
    assert res.status in (1, 2) and np.linalg.norm(res.x) < 1e-7


def test_callback_and_return_all():
    seen = []
    res = ipb.minimize_constrained(
        e2e_problems._rosen_fun, np.array([-1.2, 1.0]), e2e_problems._rosen_grad,
        e2e_problems._rosen_hess, (), hot_path=ORACLE_PATH, options={"return_all": True},
        callback=lambda state: seen.append(state.niter) or len(seen) >= 5)
    assert res.status == 3 and len(seen) == 5          # the callback asked to stop
    assert len(res.allvecs) == res.niter and len(res.allmult) == res.niter
