"""B200-native drop-in for the linear-algebra hot path of
antonior92/ip-nonlinear-solver (``ipsolver._large_scale_constrained``):
``projections``, ``projected_cg``, ``modified_dogleg`` and the Hessian /
Jacobian products, as persistent fp64 CUDA kernels behind a C ABI.

Importing the package does not touch CUDA; the first device call creates the
context (and raises if the CUDA library or a GPU is missing -- there is no
CPU fallback)."""

from . import projections, qp_subproblem  # noqa: F401  (submodules, as in the reference)
from .qp_subproblem import (projected_cg, modified_dogleg,  # noqa: F401
                            HessianOperator)
from .equality_constrained_sqp import equality_constrained_sqp  # noqa: F401
from .tr_interior_point import tr_interior_point  # noqa: F401
from .minimize_constrained import (minimize_constrained, NonlinearConstraint,  # noqa: F401
                                   LinearConstraint, BoxConstraint)

__all__ = ["projections", "qp_subproblem", "projected_cg", "modified_dogleg",
           "HessianOperator", "equality_constrained_sqp", "tr_interior_point",
           "minimize_constrained", "NonlinearConstraint", "LinearConstraint",
           "BoxConstraint"]
