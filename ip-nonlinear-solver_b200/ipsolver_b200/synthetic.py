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


# ---------------------------------------------------------------------------
# index-addressable variant ("hashed"): every entry is a pure function of
# (seed, stream, index), so any block of rows can be generated on its own --
# what a row-partitioned run at n = 1e8 (BASELINE config 5) needs: no rank
# ever holds the whole problem.  Same structure as above (5-diagonal SPD H
# with log-uniform diagonal in [1, kappa]; A with pivot 2.5+U(0,1) at column
# 2i and two U(-1,1) entries within +-8 columns), different random stream.
# ---------------------------------------------------------------------------
_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _u01(seed, stream, idx):
    """Uniform(0,1) doubles, a splitmix64 hash of (seed, stream, idx)."""
    with np.errstate(over='ignore'):
        z = (np.asarray(idx, dtype=np.uint64)
             + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
             + np.uint64(stream) * np.uint64(0xD1B54A32D192ED03))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def _normal(seed, stream, idx):
    u1, u2 = _u01(seed, stream, idx), _u01(seed, stream + 1, idx)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def hashed_hessian_rows(n, r0, r1, kappa=1e4, seed=0):
    """Rows [r0, r1) of the n x n Hessian, CSR with GLOBAL column indices."""
    i = np.arange(r0, r1, dtype=np.int64)
    diag = kappa ** _u01(seed, 0, i)
    off = lambda stream, k: 0.2 * (2.0 * _u01(seed, stream, k) - 1.0)   # noqa: E731
    cols = np.stack([i - 2, i - 1, i, i + 1, i + 2], axis=1)
    # entry (i, i+1) is off1[i], (i, i-1) is off1[i-1]; same for the second diagonal
    vals = np.stack([off(2, i - 2), off(1, i - 1), diag, off(1, i), off(2, i)], axis=1)
    keep = (cols >= 0) & (cols < n)
    indptr = np.concatenate(([0], np.cumsum(keep.sum(axis=1))))
    return sps.csr_matrix((vals[keep], cols[keep], indptr), shape=(r1 - r0, n))


def hashed_jacobian_rows(n, r0, r1, seed=0):
    """Rows [r0, r1) of the (n/2) x n windowed Jacobian, CSR with GLOBAL columns."""
    i = np.arange(r0, r1, dtype=np.int64)
    pivot = 2 * i
    e1 = np.clip(pivot + np.floor(16.0 * _u01(seed, 10, 2 * i)).astype(np.int64) - 8, 0, n - 1)
    e2 = np.clip(pivot + np.floor(16.0 * _u01(seed, 10, 2 * i + 1)).astype(np.int64) - 8, 0, n - 1)
    cols = np.stack([pivot, e1, e2], axis=1).ravel()
    vals = np.stack([2.5 + _u01(seed, 11, i), 2.0 * _u01(seed, 12, 2 * i) - 1.0,
                     2.0 * _u01(seed, 12, 2 * i + 1) - 1.0], axis=1).ravel()
    rows = np.repeat(np.arange(r1 - r0, dtype=np.int64), 3)
    A = sps.coo_matrix((vals, (rows, cols)), shape=(r1 - r0, n)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def hashed_vectors(n0, n1, m0, m1, seed=0):
    """c[n0:n1] and b[m0:m1] of the hashed problem."""
    return (_normal(seed, 20, np.arange(n0, n1, dtype=np.int64)),
            _normal(seed, 22, np.arange(m0, m1, dtype=np.int64)))


def make_qp_hashed(n, kappa=1e4, seed=0):
    """The whole hashed problem (for tests and single-GPU comparison)."""
    m = n // 2
    c, b = hashed_vectors(0, n, 0, m, seed)
    return hashed_hessian_rows(n, 0, n, kappa, seed), hashed_jacobian_rows(n, 0, m, seed), c, b
