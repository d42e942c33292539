"""Host logic of the row-partitioned multi-GPU path, on the CPU: the block
partition, local column numbering, ghost lists and send lists reproduce the
global products exactly; a world_size-2 gloo run exercises the collective
plumbing (``all_gather_object`` of ghost lists) the device layer uses."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sps

from ipsolver_b200 import synthetic
from ipsolver_b200 import distributed as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def simulate(H, A, size, x_n, y_m):
    """Run every rank's local products on the host, with ghost values moved by
    the send lists exactly as the kernels move them; compare with global."""
    n, m = H.shape[0], A.shape[0]
    part = D.RowPartition(n, m, size)
    parts = [D.local_parts(part, g, H, A) for g in range(size)]
    all_ghosts = [p["ghosts"] for p in parts]
    glo, ghi = D.ghost_capacity(all_ghosts, size)
    sends = [D.send_lists(part, g, all_ghosts) for g in range(size)]
    vec = {0: x_n, 1: y_m}
    # extended local vectors [ghost_lo capacity | own | ghost_hi capacity], NaN-filled ghosts
    ext = {}
    for g in range(size):
        for sp in (0, 1):
            a, b = part.own(g, sp)
            e = np.full(glo[sp] + (b - a) + ghi[sp], np.nan)
            e[glo[sp]:glo[sp] + (b - a)] = vec[sp][a:b]
            ext[g, sp] = e
    for g in range(size):                 # the push: owner writes into the peer's ghost zone
        for sp in (0, 1):
            for src, peer, dst, slot in sends[g][sp]:
                # the receiver derives the ghost index from the mailbox slot alone
                own_peer = part.own_len(peer, sp)
                assert (slot - glo[sp] if slot < glo[sp] else own_peer + slot - glo[sp]) == dst
                assert 0 <= slot < glo[sp] + ghi[sp]
                ext[peer, sp][glo[sp] + dst] = ext[g, sp][glo[sp] + src]
    An = sps.diags(1.0 / np.sqrt(np.asarray(A.multiply(A).sum(axis=1)).ravel())) @ A
    want = {"H": H.dot(x_n), "A_raw": A.dot(x_n), "A_norm": An.dot(x_n),
            "At_raw": A.T.dot(y_m), "At_norm": An.T.dot(y_m), "M": (An @ An.T).dot(y_m)}
    space = {"H": 0, "A_raw": 0, "A_norm": 0, "At_raw": 1, "At_norm": 1, "M": 1}
    rows = {"H": 0, "A_raw": 1, "A_norm": 1, "At_raw": 0, "At_norm": 0, "M": 1}
    for name in want:
        got = []
        for g in range(size):
            L = parts[g][name]
            sp = space[name]
            a, b = part.own(g, sp)
            e = ext[g, sp]
            idx = L.indices.astype(np.int64) + glo[sp]
            assert idx.min(initial=0) >= 0 and idx.max(initial=0) < e.shape[0]
            E = sps.csr_matrix((L.data, idx, L.indptr), shape=(L.shape[0], e.shape[0]))
            used = np.unique(idx)
            assert not np.isnan(e[used]).any(), "a ghost entry was never pushed"
            got.append(E.dot(np.nan_to_num(e)))
        got = np.concatenate(got)
        ra, rb = 0, (H.shape[0] if rows[name] == 0 else A.shape[0])
        assert got.shape[0] == rb - ra
        np.testing.assert_allclose(got, want[name], rtol=1e-13, atol=1e-13)
    return part, parts, sends


@pytest.mark.parametrize("pattern", ["window", "rand"])
@pytest.mark.parametrize("size", [1, 2, 3, 8])
def test_partitioned_products_match_global(pattern, size):
    n = 4000
    H, A, c, b = synthetic.make_qp(n, pattern, 1e3, seed=0)
    part, parts, sends = simulate(H, A, size, c, b)
    assert part.n_off[0] == 0 and part.n_off[-1] == n and (part.n_off[:-1] % 2 == 0).all()
    if pattern == "window" and size > 1:
        # halo of a windowed matrix is a handful of entries per neighbour
        assert max(len(s[0]) for s in sends) <= 64
        for p in parts:
            assert len(p["ghosts"][0][0]) <= 32 and len(p["ghosts"][0][1]) <= 32


def test_ragged_partition_more_ranks_than_rows():
    H, A, c, b = synthetic.make_qp(12, "window", 10.0, seed=0)
    simulate(H, A, 8, c, b)


WORKER = r'''
import os, sys, numpy as np
sys.path[:0] = [%r, %r]
import torch.distributed as dist
from ipsolver_b200 import synthetic, distributed as D
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% sys.argv[1],
                        rank=int(sys.argv[2]), world_size=2)
rank = dist.get_rank()
H, A, c, b = synthetic.make_qp(2000, "window", 1e3, seed=0)
part = D.RowPartition(2000, 1000, 2)
p = D.local_parts(part, rank, H, A)
allg = [None, None]
dist.all_gather_object(allg, p["ghosts"])
sends = D.send_lists(part, rank, allg)
# every ghost my peer holds that I own is in my send list, once
peer = 1 - rank
for sp in (0, 1):
    lo, hi = part.own(rank, sp)
    want = sum(int(((g >= lo) & (g < hi)).sum()) for g in allg[peer][sp])
    assert sends[sp].shape[0] == want, (sends[sp].shape, want)
    assert (sends[sp][:, 1] == peer).all()
# the local product with ghosts fetched from the peer equals the global one
x_parts = [None, None]
dist.all_gather_object(x_parts, c[part.own(rank, 0)[0]:part.own(rank, 0)[1]])
x = np.concatenate(x_parts)
glo_v, ghi_v = x[p["ghosts"][0][0]], x[p["ghosts"][0][1]]
y = D.apply_local(p["A_raw"], x_parts[rank], glo_v, ghi_v)
m0, m1 = part.own(rank, 1)
assert np.allclose(y, A.dot(c)[m0:m1], rtol=1e-13, atol=1e-13)
dist.barrier()
print("rank", rank, "ok")
'''


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, os.path.join(ROOT, "ip-nonlinear-solver_b200")))
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), port, str(r)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o


@pytest.mark.parametrize("n", [4000, 50, 14])
@pytest.mark.parametrize("size", [1, 2, 3, 8])
def test_generated_parts_equal_sliced_parts(n, size):
    """BASELINE config 5 path: each rank builds its pieces from a window of the
    index-addressable synthetic QP; they must be the pieces sliced out of the
    whole problem."""
    H, A, c, b = synthetic.make_qp_hashed(n, 1e4, 0)
    assert abs(H - H.T).max() == 0 and A.shape == (n // 2, n)
    part = D.RowPartition(n, n // 2, size)
    fro2 = 0.0
    for g in range(size):
        a, w = D.local_parts(part, g, H, A), D.local_parts_generated(part, g, 1e4, 0)
        for k in ("H", "A_raw", "A_norm", "At_raw", "At_norm", "M"):
            assert np.array_equal(a[k].indptr, w[k].indptr), k
            assert np.array_equal(a[k].indices, w[k].indices), k
            np.testing.assert_allclose(a[k].data, w[k].data, rtol=1e-15, atol=0)
        for sp in (0, 1):
            for j in (0, 1):
                assert np.array_equal(a["ghosts"][sp][j], w["ghosts"][sp][j])
        assert np.array_equal(a["dscale"], w["dscale"])
        n0, n1 = part.own(g, 0)
        m0, m1 = part.own(g, 1)
        cl, bl = synthetic.hashed_vectors(n0, n1, m0, m1, 0)
        assert np.array_equal(cl, c[n0:n1]) and np.array_equal(bl, b[m0:m1])
        fro2 += w["fro2_own"]
    assert np.sqrt(fro2) == pytest.approx(a["fro"], rel=1e-13)


@pytest.mark.parametrize("size", [2, 3, 8])
def test_band_window_of_a_rank_solves_like_the_global_factor(size):
    """Banded direct solve on row-partitioned ranks (host model): every rank
    factorises its own rows of M = A~ A~' plus a margin and sweeps with 48 halo
    rows of its neighbours' right-hand side; the own rows of the result equal
    the global solve to round-off, the m-space ghost zones are exactly the
    contiguous halo rows and the window equals the slice of the global M."""
    import scipy.sparse.linalg as spla
    n = 6000
    H, A, c, b = synthetic.make_qp(n, "window", 1e2, seed=0)
    m = A.shape[0]
    part = D.RowPartition(n, m, size)
    d = 1.0 / np.sqrt(np.asarray(A.multiply(A).sum(axis=1)).ravel())
    An = sps.csr_matrix(sps.diags(d) @ A)
    M = sps.csr_matrix(An @ An.T)
    rhs = np.random.RandomState(3).standard_normal(m)
    v_ref = spla.spsolve(M.tocsc(), rhs)
    for g in range(size):
        parts = D.local_parts(part, g, H, A)
        bd = parts["band"]
        assert bd is not None
        m0, m1 = part.own(g, 1)
        e0, e1, gb, ga = D.band_window(part, g)
        lo, hi = parts["ghosts"][1]
        assert np.array_equal(lo, np.arange(m0 - gb, m0)) and np.array_equal(hi, np.arange(m1, m1 + ga))
        assert bd["ext_off"] == m0 - e0
        assert abs(bd["M_ext"] - M[e0:e1, e0:e1]).max() < 1e-15
        # what the device does: exact solve on the window [own - halo, own + halo) with the
        # right-hand side known there (zero beyond), keep the own rows
        w0, w1 = m0 - gb, m1 + ga
        v_loc = spla.spsolve(M[w0:w1, w0:w1].tocsc(), rhs[w0:w1])
        err = np.linalg.norm(v_loc[m0 - w0:m1 - w0] - v_ref[m0:m1]) / np.linalg.norm(v_ref[m0:m1])
        assert err < 1e-13


SQP_WORKER = r'''
import os, sys, numpy as np
sys.path[:0] = [%r, %r]
import torch.distributed as dist
import oracle
import ipsolver_b200 as ipb
from ipsolver_b200 import synthetic, distributed as D
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% sys.argv[1],
                        rank=int(sys.argv[2]), world_size=2)
rank = dist.get_rank()
n = 400
H, A, q, b = synthetic.make_qp(n, "window", 10.0, seed=0)
m = A.shape[0]

class FakeProj:
    """Stands in for the device: every call gathers its block inputs into full
    vectors (as the kernels see them through ghosts and reductions), computes
    with the oracle on the whole problem and hands back the rank's own block."""
    def __init__(self, S):
        self.S = S
        self.Z, self.LS, self.Y = oracle.build_projectors(A)
        self.c = self.b = self.lb = self.ub = None
    def own(self, v, space):
        return self.S.local(v, space)
    def sqp_normal_step(self, c, b, radius, lb, ub, bounds_changed, box_factor):
        if c is not None: self.c = self.S.allgather(c, 0)
        if b is not None: self.b = self.S.allgather(b, 1)
        if lb is not None and bounds_changed:
            self.lb, self.ub = self.S.allgather(lb, 0), self.S.allgather(ub, 0)
        lbn = np.full(n, -np.inf) if self.lb is None else box_factor * self.lb
        ubn = np.full(n, np.inf) if self.ub is None else box_factor * self.ub
        self.dn = oracle.dogleg_normal_step(A, self.Y, self.b, radius, lbn, ubn)
        return self.own(self.dn, 0)
    def sqp_tangential_step(self, Hdev, radius_t):
        assert Hdev == "H of the ranks"
        lb = np.full(n, -np.inf) if self.lb is None else self.lb
        ub = np.full(n, np.inf) if self.ub is None else self.ub
        dt, info = oracle.eqp_projected_cg(H, H.dot(self.dn) + self.c, self.Z, self.Y, np.zeros(m),
                                           radius_t, lb - self.dn, ub - self.dn)
        d = self.dn + dt
        return self.own(d, 0), self.own(H.dot(d), 0), self.own(A.dot(d), 1), info
    def sqp_multipliers(self, c):
        self.c = self.S.allgather(c, 0)
        v = -self.LS.dot(self.c)
        r = np.abs(self.c + A.T.dot(v))
        n0, n1 = self.S.part.own(rank, 0)
        return self.own(v, 1), float(r[n0:n1].max())      # the rank's share of the inf-norm

S = object.__new__(D.DistSolver)
S.rank, S.size, S.group = rank, 2, dist.group.WORLD
S.part = D.RowPartition(n, m, 2)
S.ctx = type("Ctx", (), {"device": 0})()
S.H = "H of the ranks"
S.proj = FakeProj(S)
# the second-order correction's helpers (Y.dot, box_intersections) go to the device too
S.apply_map = lambda which, x: {"Z": S.proj.Z, "LS": S.proj.LS, "Y": S.proj.Y}[which].dot(x)
S.box_intersections = oracle.segment_box
# allgather / allmax plumbing
x = np.arange(n, dtype=float)
assert np.array_equal(S.allgather(S.local(x, 0), 0), x)
assert np.array_equal(S.allgather(S.local(x[:m], 1), 1), x[:m])
assert S.allmax(float(rank)) == 1.0
args = (lambda x: 0.5 * x.dot(H.dot(x)) + q.dot(x), np.zeros(n), lambda x: H.dot(x) + q, lambda x: H)
con = ipb.LinearConstraint(A, ("equals", b))
res = ipb.minimize_constrained(*args, con, method='equality_constrained_sqp', hot_path=S.hot_path(H, A))
ref = ipb.minimize_constrained(*args, con, method='equality_constrained_sqp',
                               hot_path=oracle.host_hot_path())
assert (res.niter, res.cg_niter, res.nfev, res.status) == (ref.niter, ref.cg_niter, ref.nfev, ref.status), \
    ((res.niter, res.cg_niter, res.nfev, res.status), (ref.niter, ref.cg_niter, ref.nfev, ref.status))
assert np.array_equal(res.x, ref.x) and np.array_equal(res.v, ref.v)
assert res.optimality == ref.optimality
# a problem whose matrices are not the solver's is refused, not silently mis-solved
try:
    S.hot_path(H, A).projections(A * 2.0)
    raise SystemExit("changed Jacobian accepted")
except NotImplementedError:
    pass
dist.barrier()
print("rank", rank, "ok", res.niter, res.cg_niter)
'''


def test_row_partitioned_sqp_loop_world_size_2(tmp_path):
    """``DistSolver.hot_path()`` under the SQP loop, two gloo ranks, the device
    replaced by the oracle working on gathered vectors: block slicing,
    all-gathers, the global box / fuse decisions and the replicated host loop
    reproduce the single-process run exactly."""
    script = tmp_path / "sqp_worker.py"
    script.write_text(SQP_WORKER % (ROOT, os.path.join(ROOT, "ip-nonlinear-solver_b200")))
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), port, str(r)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o[-3000:]
        assert "rank %d ok" % r in o
