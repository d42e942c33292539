"""CPU oracle for the projected-CG / null-space-projection hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The
product path (``ip-nonlinear-solver_b200/``) never imports this package and
fails loudly when its CUDA library is missing.

Parity status: PINNED.  ``oracle/refpath.py`` restates the reference's numpy /
scipy algorithm (reference file:line cited per function) and is checked
bit-for-bit against outputs of the reference itself, imported from
``/root/reference`` in the build container by ``tests/golden/make_golden.py``
(vectors committed under ``tests/golden/``), plus the reference's own
known-answer tests re-stated in ``tests/test_oracle_*.py``.  One caveat, said
here and in DESIGN.md: the reference's ``NormalEquation`` projector calls
CHOLMOD through scikit-sparse, which is not installed in this image (and is
not vendored by the reference), so that variant's bit pattern is unpinned; it
is pinned against the mathematically identical ``AugmentedSystem`` projector
(SuperLU, runnable here) to 1e-12.
"""
from .refpath import (  # noqa: F401
    segment_ball, segment_box, segment_box_ball, in_box, clip_to_box,
    dogleg_normal_step, eqp_projected_cg, kkt_direct_solve,
    orthogonality_measure, build_projectors, HostLagrangianHessian, host_hot_path,
)
