"""The reference itself next to the repo (``baseline/_ref``, installed by
``scripts/install_reference.sh``; see ``tests/ref_harness.py``).

* CPU: the reference's own unit tests run unmodified through the harness
  (numpy.testing shims only) -- the harness is sound and the install is the
  reference.
* GPU: the SAME test files run unchanged with the L1 functions rebound to the
  device path (``IPB_REF_PATCH=1``): ``test_qp_subproblem.py``,
  ``test_projections.py`` and ``test_minimized_constrained.py`` of the
  reference are then tests of ``libipsolver_b200.so``.
* GPU: the INTEGRATION.md patch -- the reference's ``minimize_constrained``,
  constraint classes, SQP and barrier loops on top of the repo's hot path --
  against the golden results the unpatched reference produced.

Skipped (never failed) where ``baseline/_ref`` is absent: the install is a
git-ignored artefact of the build container.
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import e2e_problems
import ref_harness

ROOT = ref_harness.ROOT
needs_ref = pytest.mark.skipif(not ref_harness.available(),
                               reason="baseline/_ref not installed (scripts/install_reference.sh)")

# Reference unit tests that cannot hold for ANY implementation but the reference's own
# factorisations: none so far -- every deselection must be listed here with its reason.
DEVICE_DESELECT = []


def run_reference_tests(patched, extra=()):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests"), ref_harness.REF,
                                         os.path.join(ROOT, "ip-nonlinear-solver_b200"), ROOT,
                                         env.get("PYTHONPATH", "")])
    env["IPB_REF_PATCH"] = "1" if patched else "0"
    lsc = ref_harness.test_dirs()[0]
    files = [os.path.join(lsc, "test_projections.py"), os.path.join(lsc, "test_qp_subproblem.py"),
             os.path.join(ref_harness.test_dirs()[1], "test_minimized_constrained.py")]
    cmd = [sys.executable, "-m", "pytest", "-p", "ref_harness", "-q", "-p", "no:cacheprovider",
           "-W", "ignore", "-rf", "--rootdir", ref_harness.REF] + list(extra) + files
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    text = out.stdout + out.stderr
    m = re.search(r"(\d+) passed", text)
    f = re.search(r"(\d+) failed", text)
    return out.returncode, int(m.group(1)) if m else 0, int(f.group(1)) if f else 0, text


@needs_ref
def test_reference_own_tests_pass_unmodified_on_cpu():
    rc, passed, failed, text = run_reference_tests(patched=False)
    assert rc == 0 and failed == 0 and passed >= 35, text[-3000:]


@needs_ref
@pytest.mark.gpu
def test_reference_own_tests_pass_against_the_device_functions():
    """35 tests of the reference: 11 on projections / orthogonality (sparse AND
    dense ``A``, every ``method`` name), 18 on the intersections, the dogleg and
    ``projected_cg`` (known answers, error cases, ``return_all``), 6 end-to-end
    ``minimize_constrained`` classes (Maratos, Rosenbrock variants, Elec, ...)."""
    extra = []
    for name in DEVICE_DESELECT:
        extra += ["--deselect", name]
    rc, passed, failed, text = run_reference_tests(patched=True, extra=extra)
    print(text[-4000:])
    assert failed == 0 and rc == 0, text[-6000:]
    assert passed >= 35 - len(DEVICE_DESELECT)
    assert "DeviceLibraryError" not in text


# ------------------------------------------------------------------ INTEGRATION.md patch
@pytest.fixture(scope="module")
def patched_api():
    ipsolver = ref_harness.apply_integration_patch()
    return dict(minimize_constrained=ipsolver.minimize_constrained,
                LinearConstraint=ipsolver.LinearConstraint,
                BoxConstraint=ipsolver.BoxConstraint,
                NonlinearConstraint=ipsolver.NonlinearConstraint)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "e2e.npz"))


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - b) / max(np.linalg.norm(b), 1e-300)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["eqp_window_2000_gtol6", "eqp_rand_2000_gtol6",
                                  "eqp_window_20000_gtol6"])
def test_patched_reference_sqp_matches_its_own_golden(patched_api, gold, name):
    """The reference's loop, the repo's L1: same counts (+-1), x / v to 1e-8."""
    from ipsolver_b200.device import Context
    launches0 = Context.default().launches
    niter, cg, status, nfev, opt, viol, fval, radius = gold[name + "/scalars"]
    res = e2e_problems.PROBLEMS[name](patched_api)()
    assert Context.default().launches > launches0          # the device did the work
    assert abs(res.niter - niter) <= 1 and abs(res.cg_niter - cg) <= 1
    assert res.status == status and abs(res.nfev - nfev) <= 1
    assert rel(res.x, gold[name + "/x"]) < 1e-8
    assert rel(res.v, gold[name + "/v"]) < 1e-8
    assert res.optimality == pytest.approx(opt, rel=1e-6)
    assert res.fun == pytest.approx(fval, rel=1e-12)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["maratos", "hyperbolic_ineq", "rosenbrock_ineq",
                                  "rosenbrock_eq_ineq", "rosenbrock_box_enforced"])
def test_patched_reference_small_classics(patched_api, gold, name):
    niter, cg, status, nfev, opt, viol, fval, radius = gold[name + "/scalars"]
    res = e2e_problems.PROBLEMS[name](patched_api)()
    assert res.status == status
    assert abs(res.niter - niter) <= 1 and abs(res.cg_niter - cg) <= 2
    assert np.linalg.norm(res.x - gold[name + "/x"]) <= 1e-7 * max(1.0, np.linalg.norm(gold[name + "/x"]))
    assert res.fun == pytest.approx(fval, rel=1e-8, abs=1e-12)


@needs_ref
@pytest.mark.gpu
def test_patched_reference_interior_point(patched_api, gold):
    """Barrier subproblems of the reference's tr_interior_point with the slack
    block as the device operator's diagonal tail."""
    name = "ipqp_window_1000"
    niter, cg, status, nfev, opt, viol, fval, radius = gold[name + "/scalars"]
    res = e2e_problems.PROBLEMS[name](patched_api)()
    assert res.status == status == 1
    assert res.optimality < 1e-8 and res.constr_violation < 1e-8
    assert abs(res.niter - niter) <= 0.1 * niter
    assert rel(res.x, gold[name + "/x"]) < 1e-6
    assert res.fun == pytest.approx(fval, rel=1e-8)
