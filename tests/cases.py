"""Inputs shared by the golden-vector script and the parity tests.

The small dense problems are the textbook examples the reference's own
tests use (Nocedal & Wright, "Numerical Optimization", ex. 16.2 and
variations; reference ``tests/test_qp_subproblem.py:430-649``,
``tests/test_projections.py:25-139``).  The synthetic ones follow
SURVEY.md section 8d.
"""
import numpy as np

INF = np.inf

H3 = [[6, 2, 1], [2, 5, 2], [1, 2, 4]]
A3 = [[1, 0, 1], [0, 1, 1]]
H4 = [[6, 2, 1, 3], [2, 5, 2, 4], [1, 2, 4, 5], [3, 4, 5, 7]]
A4 = [[1, 0, 1, 0], [0, 1, 1, 1]]
H4_INDEF = [[1, 2, 1, 3], [2, 0, 2, 4], [1, 2, 0, 2], [3, 4, 2, 0]]
A4_INDEF = [[1, 0, 1, 0], [0, 1, 0, 1]]
C4 = [-2, -3, -3, 1]
B2 = [-3, 0]

# name -> (H, A, c, b, kwargs, expected exception or None)
PCG_CASES = {
    "nocedal_16_2": (H3, A3, [-8, -3, -3], B2, {}, None),
    "direct_fact": (H4, A4, C4, B2, dict(tol=0), None),
    "tr_infeasible": (H4, A4, C4, B2, dict(trust_radius=1), ValueError),
    "tr_barely_feasible": (H4, A4, C4, B2,
                           dict(tol=0, trust_radius=2.32379000772445021283),
                           None),
    "hits_boundary": (H4, A4, C4, B2, dict(tol=0, trust_radius=3), None),
    "neg_curv_unconstrained": (H4_INDEF, A4_INDEF, C4, B2, dict(tol=0),
                               ValueError),
    "neg_curv": (H4_INDEF, A4_INDEF, C4, B2,
                 dict(tol=0, trust_radius=1000), None),
    "box_inactive": (H4, A4, C4, B2,
                     dict(tol=0, lb=[0.5, -INF, -INF, -INF], return_all=True),
                     None),
    "box_active_maxiter": (H4, A4, C4, B2,
                           dict(tol=0, lb=[0.8, -INF, -INF, -INF],
                                return_all=True), None),
    "box_active_hits": (H4, A4, C4, B2,
                        dict(tol=0, ub=[INF, INF, 1.6, INF], trust_radius=3,
                             return_all=True), None),
    "box_active_hits_infeasible_iter": (
        H4, A4, C4, B2,
        dict(tol=0, ub=[INF, 0.1, INF, INF], trust_radius=4, return_all=True),
        None),
    "box_active_neg_curv": (H4_INDEF, A4_INDEF, C4, B2,
                            dict(tol=0, ub=[INF, INF, 100, INF],
                                 trust_radius=1000), None),
}

# modified dogleg: name -> (A, b, radius, lb, ub)
DOGLEG_CASES = {
    "2d_newton_inside": ([[1, 8]], [-16], 2, [-INF, -INF], [INF, INF]),
    "2d_sphere_active": ([[1, 8]], [-16], 1, [-INF, -INF], [INF, INF]),
    "2d_box_active": ([[1, 8]], [-16], 2, [-INF, -INF], [0.1, INF]),
    "3d_newton_inside": ([[1, 8, 1], [4, 2, 2]], [-16, 2], 3,
                         [-INF] * 3, [INF] * 3),
    "3d_cauchy_newton_sphere": ([[1, 8, 1], [4, 2, 2]], [-16, 2], 2,
                                [-INF] * 3, [INF] * 3),
    "3d_cauchy_newton_box": ([[1, 8, 1], [4, 2, 2]], [-16, 2], 5,
                             [-1, -INF, -INF], [INF] * 3),
    "3d_origin_cauchy_sphere": ([[1, 8, 1], [4, 2, 2]], [-16, 2], 1,
                                [-INF] * 3, [INF] * 3),
    "3d_origin_newton_box": ([[1, 8, 1], [4, 2, 2]], [-16, 2], 2,
                             [-INF] * 3, [INF, 1, INF]),
}

# projections: 3 x 8 matrix and its probe points
A38 = [[1, 2, 3, 4, 0, 5, 0, 7],
       [0, 8, 7, 0, 1, 5, 9, 0],
       [1, 0, 0, 0, 0, 1, 2, 3]]
P38_N = [[1, 2, 3, 4, 5, 6, 7, 8],
         [1, 10, 3, 0, 1, 6, 7, 8],
         [1.12, 10, 0, 0, 100000, 6, 0.7, 8],
         [1, 0, 0, 0, 0, 1, 2, 3 + 1e-10]]
P38_M = [[1, 2, 3], [1, 10, 3], [1.12, 10, 0]]

# segment / ball and segment / box cases: (z, d, [lb, ub,] radius/None, line)
BALL_CASES = [
    ([0, 0], [1, 0], 0.5, False), ([2, 0], [0, 1], 1, False),
    ([2, 0], [1, 0], 1, False), ([2, 0], [-4, 0], 1, False),
    ([2, 0], [-4, 0], 1, True), ([0, 0], [1, 0], 0.5, True),
    ([2, 0], [1, 0], 1, True), ([1, 1], [0, 0], 3, False),
    ([0.3, -0.2], [1.5, 2.5], INF, False), ([0.3, -0.2], [1.5, 2.5], INF, True),
]
BOX_CASES = [
    ([1, 1], [1, 1], [0, 0], [2, 2], False),
    ([2, 0], [0, 2], [1, 1], [3, 3], False),
    ([0, 0], [4, 4], [1, -3], [3, 3], False),
    ([0, 0], [4, 4], [1, -3], [3, 3], True),
    ([0, 0], [-4, -4], [1, -3], [3, 3], False),
    ([0, 0], [-4, -4], [1, -3], [3, 3], True),
    ([1, 0], [0, 1], [1, 1], [3, 3], False),
    ([0, 0], [0, 1], [1, 1], [3, 3], False),
    ([0, 0, 0], [4, 4, 4], [1, -3, -INF], [3, 3, INF], False),
    ([1, 1, 1], [0, 0, 0], [0, 0, 0], [2, 2, 2], False),
    ([0, 0], [1, 100], [-INF, -INF], [INF, 3], True),
]

# synthetic projected-CG configurations, small sizes
SYNTH_N = 2000
SYNTH_RUNS = {
    # name -> (pattern, kappa, kwargs, box tau or None)
    "window_default": ("window", 1e4, {}, None),
    "window_forced30": ("window", 1e4, dict(tol=1e-30, max_iter=30), None),
    "window_radius": ("window", 1e4, dict(trust_radius=12.0), None),
    "window_box": ("window", 1e4, dict(tol=1e-30, max_iter=40), 1.5),
    "window_box_radius": ("window", 10.0, dict(trust_radius=16.0), 1.4),
    "rand_default": ("rand", 1e4, {}, None),
    "rand_forced30": ("rand", 1e4, dict(tol=1e-30, max_iter=30), None),
    "rand_radius_box": ("rand", 10.0, dict(trust_radius=14.0), 1.4),
}
