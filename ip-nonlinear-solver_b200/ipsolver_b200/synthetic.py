"""Deterministic synthetic sparse QPs (SURVEY.md section 8d / BASELINE.md section 3).

Used by the tests, the golden-vector script and ``bench.py`` so that every arm
(CUDA path, oracle, reference) sees the same inputs.  Pure numpy/scipy, no
device code.

  H : n x n CSR, SPD, half-bandwidth 2 (5 nnz/row), diagonal
      ``logspace(0, log10(kappa), n)`` shuffled, off-diagonals ``0.2*U(-1,1)``
      mirrored so H is symmetric, strictly diagonally dominant.
  A : m x n CSR, m = n/2, 3 nnz/row: pivot ``2.5 + U(0,1)`` at column 2i plus
      two ``U(-1,1)`` entries at ``clip(2i + randint(-8, 8))`` (``window``) or
      ``randint(0, n)`` (``rand``); duplicates are summed.
  c ~ N(0,1)^n, b ~ N(0,1)^m.
"""
import numpy as np
import scipy.sparse as sps


def banded_spd_hessian(n, kappa=1e4, seed=0):
    rng = np.random.RandomState(seed)
    diag = np.logspace(0.0, np.log10(kappa), n)[rng.permutation(n)]
    off1 = 0.2 * rng.uniform(-1.0, 1.0, n - 1)
    off2 = 0.2 * rng.uniform(-1.0, 1.0, n - 2)
    H = sps.diags([off2, off1, diag, off1, off2], [-2, -1, 0, 1, 2],
                  shape=(n, n), format='csr', dtype=np.float64)
    H.sort_indices()
    return H


def sparse_jacobian(n, m=None, pattern='window', seed=1):
    if m is None:
        m = n // 2
    rng = np.random.RandomState(seed)
    rows = np.repeat(np.arange(m, dtype=np.int64), 3)
    pivot = 2 * np.arange(m, dtype=np.int64)
    if pattern == 'window':
        extra = np.clip(pivot[:, None] + rng.randint(-8, 8, size=(m, 2)),
                        0, n - 1)
    elif pattern == 'rand':
        extra = rng.randint(0, n, size=(m, 2))
    else:
        raise ValueError("pattern must be 'window' or 'rand'")
    cols = np.hstack([pivot[:, None], extra]).ravel()
    vals = np.hstack([2.5 + rng.uniform(0.0, 1.0, (m, 1)),
                      rng.uniform(-1.0, 1.0, (m, 2))]).ravel()
    A = sps.coo_matrix((vals, (rows, cols)), shape=(m, n)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def qp_vectors(n, m, seed=2):
    rng = np.random.RandomState(seed)
    return rng.standard_normal(n), rng.standard_normal(m)


def box_bounds(n, tau, frac=0.1, seed=3):
    """``lb=-tau, ub=+tau`` on a random ``frac`` of coordinates, inf elsewhere."""
    rng = np.random.RandomState(seed)
    mask = rng.uniform(size=n) < frac
    lb = np.where(mask, -tau, -np.inf)
    ub = np.where(mask, tau, np.inf)
    return lb, ub


def make_qp(n, pattern='window', kappa=1e4, seed=0):
    """Returns ``(H, A, c, b)`` for the equality-constrained QP
    ``min 1/2 x'Hx + c'x  s.t.  Ax + b = 0``."""
    m = n // 2
    H = banded_spd_hessian(n, kappa, seed)
    A = sparse_jacobian(n, m, pattern, seed + 1)
    c, b = qp_vectors(n, m, seed + 2)
    return H, A, c, b
