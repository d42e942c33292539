"""Row-partitioned multi-GPU runs of the hot path (one process per GPU).

The reference is single-process; this is the B200-side
extension ``north_star`` asks for.  Rank ``g`` owns a
contiguous block of variables ``[n_off[g], n_off[g+1])`` -- all n-vectors, the
rows of ``H`` and of ``A'`` -- and of constraints ``[m_off[g], m_off[g+1])`` --
all m-vectors, the rows of ``A`` and of ``M = A~ A~'``.

Pure numpy/scipy part (testable without a GPU):
  :class:`RowPartition`  block offsets
  :func:`local_parts`    this rank's matrices in local column numbering and
                         the ghost entries they address
  :func:`send_lists`     who pushes what to whom before a product

Device part: :class:`DistSolver` sets up the symmetric arena, exchanges the
CUDA-IPC handles through ``torch.distributed`` (plumbing only) and runs the
same persistent kernels; inside them ghost entries are pushed over NVLink peer
memory and reductions go through peer mailboxes (``csrc/device.cuh``,
``halo_sync`` / ``rank_exchange``).
"""
import ctypes as C

import numpy as np
import scipy.sparse as sps

N_SPACE, M_SPACE = 0, 1
# banded direct solve of (A A')^-1 (csrc/device.cuh, BandView): rows of the
# neighbours a rank's sweeps read next to its own block, and the extra margin
# of rows of M it factorises so that the truncated factor recurrence has
# forgotten its start where it is used
BAND_HALO, BAND_MARGIN = 48, 96


class RowPartition:
    """Contiguous block partition of n variables and m constraints."""

    def __init__(self, n, m, size):
        self.n, self.m, self.size = int(n), int(m), int(size)
        # variable blocks start on even indices (16-byte alignment of slices)
        self.n_off = np.array([(g * self.n // size) & ~1 for g in range(size)] + [self.n],
                              dtype=np.int64)
        self.m_off = np.array([g * self.m // size for g in range(size)] + [self.m],
                              dtype=np.int64)

    def offsets(self, space):
        return self.n_off if space == N_SPACE else self.m_off

    def own(self, rank, space):
        off = self.offsets(space)
        return int(off[rank]), int(off[rank + 1])

    def own_len(self, rank, space):
        a, b = self.own(rank, space)
        return b - a

    def own_max(self, space):
        return int(np.diff(self.offsets(space)).max())

    def owner(self, idx, space):
        return np.searchsorted(self.offsets(space), idx, side='right') - 1


def _remote_columns(M, lo, hi):
    """Sorted distinct column indices of M outside [lo, hi)."""
    c = M.indices
    c = c[(c < lo) | (c >= hi)]
    return np.unique(c)


def _localize(M, lo, hi, ghosts_lo, ghosts_hi):
    """Renumber the columns of M: own ``c -> c - lo``; ghosts below / above
    the own block go to ``-len(ghosts_lo) + position`` / ``own + position``."""
    M = sps.csr_matrix(M)
    c = M.indices.astype(np.int64)
    out = c - lo
    below, above = c < lo, c >= hi
    if below.any():
        out[below] = np.searchsorted(ghosts_lo, c[below]) - len(ghosts_lo)
    if above.any():
        out[above] = (hi - lo) + np.searchsorted(ghosts_hi, c[above])
    L = sps.csr_matrix((M.data, out.astype(np.int32), M.indptr), shape=M.shape, copy=False)
    return L


def band_window(part, rank):
    """Rows of M = A~ A~' a rank factorises for the banded direct solve:
    ``(e0, e1, ghost_below, ghost_above)`` -- own rows plus BAND_MARGIN on each
    side (clipped to the matrix), and how many rows of each neighbour the
    sweeps read."""
    m0, m1 = part.own(rank, M_SPACE)
    return (max(0, m0 - BAND_MARGIN), min(part.m, m1 + BAND_MARGIN),
            min(BAND_HALO, m0), min(BAND_HALO, part.m - m1))


def _assemble(part, rank, H_rows, A_rows, An_rows, At_rows, Ant_rows, M_rows, dscale_own,
              fro2_own, M_ext=None):
    """Local column numbering + ghost lists from this rank's row blocks (all
    with GLOBAL column indices).  ``M_ext``: rows and columns
    ``band_window(part, rank)[:2]`` of M in the window's own numbering (or
    None: no banded solve); the m-space ghost zones then also hold the sweeps'
    halo rows."""
    n0, n1 = part.own(rank, N_SPACE)
    m0, m1 = part.own(rank, M_SPACE)
    # ghost entries per space: union over the matrices whose columns live there
    gn = np.union1d(_remote_columns(H_rows, n0, n1), _remote_columns(A_rows, n0, n1))
    gm = np.union1d(_remote_columns(At_rows, m0, m1), _remote_columns(M_rows, m0, m1))
    band = None
    if M_ext is not None and m1 > m0:
        e0, e1, gb, ga = band_window(part, rank)
        halo = np.concatenate((np.arange(m0 - gb, m0), np.arange(m1, m1 + ga)))
        gm_band = np.union1d(gm, halo)
        lo_ok = np.array_equal(gm_band[gm_band < m0], np.arange(m0 - gb, m0))
        hi_ok = np.array_equal(gm_band[gm_band >= m1], np.arange(m1, m1 + ga))
        if lo_ok and hi_ok:        # the sweeps address ghost rows by offset: zones must be contiguous
            gm = gm_band
            band = {"M_ext": sps.csr_matrix(M_ext), "ext_off": m0 - e0, "ghost_below": gb,
                    "ghost_above": ga}
    ghosts = {N_SPACE: (gn[gn < n0], gn[gn >= n1]), M_SPACE: (gm[gm < m0], gm[gm >= m1])}
    loc_n = lambda X: _localize(X, n0, n1, *ghosts[N_SPACE])      # noqa: E731
    loc_m = lambda X: _localize(X, m0, m1, *ghosts[M_SPACE])      # noqa: E731
    return {
        "H": loc_n(H_rows), "A_raw": loc_n(A_rows), "A_norm": loc_n(An_rows),
        "At_raw": loc_m(At_rows), "At_norm": loc_m(Ant_rows), "M": loc_m(M_rows),
        "ghosts": ghosts, "dscale": np.ascontiguousarray(dscale_own), "fro2_own": float(fro2_own),
        "own": {N_SPACE: (n0, n1), M_SPACE: (m0, m1)}, "band": band,
    }


def _row_scaling(A):
    row_norm2 = np.asarray(A.multiply(A).sum(axis=1)).ravel()
    dscale = np.where(row_norm2 > 0, 1.0 / np.sqrt(np.where(row_norm2 > 0, row_norm2, 1.0)), 0.0)
    return row_norm2, dscale


def local_parts(part, rank, H, A):
    """This rank's pieces of the QP matrices, in local column numbering.

    ``H`` (n x n) and ``A`` (m x n) are the global scipy matrices (every rank
    holds them -- fine up to n ~ 1e7; :func:`local_parts_generated` builds the
    same pieces from the rank's own rows only).  Returns a dict with CSR
    matrices ``H``, ``A_raw``, ``A_norm``, ``At_raw``, ``At_norm``, ``M`` whose
    column indices are relative to the rank's own block, the per-space ghost
    index lists (global indices, ascending, split below / above the own
    block), the row scaling of the own rows and ``||A||_F``."""
    H = sps.csr_matrix(H)
    A = sps.csr_matrix(A)
    n0, n1 = part.own(rank, N_SPACE)
    m0, m1 = part.own(rank, M_SPACE)
    row_norm2, dscale = _row_scaling(A)
    An = sps.csr_matrix(sps.diags(dscale) @ A)       # row-normalised A
    An.sort_indices()
    At = sps.csr_matrix(A.T)
    At.sort_indices()
    Ant = sps.csr_matrix(An.T)
    Ant.sort_indices()
    M_rows = sps.csr_matrix(An[m0:m1] @ An.T)        # own rows of A~ A~'
    M_rows.sort_indices()
    e0, e1 = band_window(part, rank)[:2]
    M_ext = sps.csr_matrix(An[e0:e1] @ An[e0:e1].T)  # window of M for the banded factor
    M_ext.sort_indices()
    out = _assemble(part, rank, H[n0:n1], A[m0:m1], An[m0:m1], At[n0:n1], Ant[n0:n1], M_rows,
                    dscale[m0:m1], row_norm2[m0:m1].sum(), M_ext)
    out["fro"] = float(np.sqrt(row_norm2.sum()))
    return out


def local_parts_generated(part, rank, kappa=1e4, seed=0, margin=BAND_MARGIN + 16):
    """The same pieces for the index-addressable synthetic QP
    (``synthetic.hashed_*``), built from a WINDOW of rows around the rank's own
    block: no rank ever holds the whole problem (BASELINE config 5).  The
    window holds every row of A that touches the rank's variables or shares a
    column with one of its rows (the Jacobian's columns stay within +-8 of the
    pivot 2i, so ``margin`` rows on each side are enough); ``fro`` is missing
    -- the caller sums ``fro2_own`` over the ranks."""
    from . import synthetic
    n, m = part.n, part.m
    n0, n1 = part.own(rank, N_SPACE)
    m0, m1 = part.own(rank, M_SPACE)
    w0 = max(0, min(m0, n0 // 2) - margin)
    w1 = min(m, max(m1, n1 // 2 + 1) + margin)
    A_win = synthetic.hashed_jacobian_rows(n, w0, w1, seed)          # global columns
    c_lo = int(A_win.indices.min()) if A_win.nnz else 0
    c_hi = int(A_win.indices.max()) + 1 if A_win.nnz else 1
    assert c_lo <= n0 or n0 == n1 or w0 == 0, "window does not cover the own variables"
    # the window in its own coordinates: (w1 - w0) x (c_hi - c_lo)
    Aw = sps.csr_matrix((A_win.data, A_win.indices - c_lo, A_win.indptr),
                        shape=(w1 - w0, c_hi - c_lo))
    row_norm2, dscale = _row_scaling(Aw)
    Anw = sps.csr_matrix(sps.diags(dscale) @ Aw)
    Anw.sort_indices()
    Atw = sps.csr_matrix(Aw.T)
    Atw.sort_indices()
    Antw = sps.csr_matrix(Anw.T)
    Antw.sort_indices()
    r0, r1 = m0 - w0, m1 - w0                                         # own rows in the window
    Mw = sps.csr_matrix(Anw[r0:r1] @ Anw.T)
    Mw.sort_indices()
    e0, e1 = band_window(part, rank)[:2]
    assert w0 <= e0 and e1 <= w1, "window does not cover the banded factor's margin"
    M_ext = sps.csr_matrix(Anw[e0 - w0:e1 - w0] @ Anw[e0 - w0:e1 - w0].T)
    M_ext.sort_indices()

    def to_global(X, col_off, ncols):
        X = sps.csr_matrix(X)
        return sps.csr_matrix((X.data, X.indices.astype(np.int64) + col_off, X.indptr),
                              shape=(X.shape[0], ncols))
    v0, v1 = max(n0 - c_lo, 0), max(min(n1 - c_lo, c_hi - c_lo), 0)   # own variables in the window
    At_rows, Ant_rows = Atw[v0:v1], Antw[v0:v1]
    if (n1 - n0) != (v1 - v0):          # own variables no row of the window touches: empty rows
        pad_lo, pad_hi = max(c_lo - n0, 0), max(n1 - c_hi, 0)
        z = lambda k: sps.csr_matrix((k, w1 - w0))                    # noqa: E731
        At_rows = sps.vstack([z(pad_lo), At_rows, z(pad_hi)], "csr")
        Ant_rows = sps.vstack([z(pad_lo), Ant_rows, z(pad_hi)], "csr")
    H_rows = synthetic.hashed_hessian_rows(n, n0, n1, kappa, seed)
    return _assemble(part, rank, H_rows, to_global(Aw[r0:r1], c_lo, n),
                     to_global(Anw[r0:r1], c_lo, n), to_global(At_rows, w0, m),
                     to_global(Ant_rows, w0, m), to_global(Mw, w0, m), dscale[r0:r1],
                     row_norm2[r0:r1].sum(), M_ext)


def send_lists(part, rank, all_ghosts, glo_cap=None):
    """``all_ghosts[g][space] = (below, above)`` of every rank ``g`` (global
    indices).  Returns per space an ``(count, 4) int32`` array of
    ``{own index, peer, index on the peer (own-relative), slot in the peer's
    ghost mailbox}``: what this rank sends before a product whose input lives
    in that space.  Mailbox slots: the peer's ghost_lo entries sit right below
    ``glo_cap[space]`` (the common ghost_lo capacity), its ghost_hi entries
    start at it."""
    if glo_cap is None:
        glo_cap = ghost_capacity(all_ghosts, part.size)[0]
    out = {}
    for space in (N_SPACE, M_SPACE):
        lo, hi = part.own(rank, space)
        ents = []
        for g in range(part.size):
            if g == rank:
                continue
            below, above = all_ghosts[g][space]
            own_g = part.own_len(g, space)
            for arr, base in ((below, -len(below)), (above, own_g)):
                arr = np.asarray(arr, dtype=np.int64)
                sel = np.nonzero((arr >= lo) & (arr < hi))[0]
                if sel.size:
                    e = np.zeros((sel.size, 4), dtype=np.int32)
                    e[:, 0] = arr[sel] - lo
                    e[:, 1] = g
                    e[:, 2] = base + sel
                    e[:, 3] = glo_cap[space] + (base + sel if base < 0 else sel)
                    ents.append(e)
        out[space] = (np.vstack(ents) if ents else np.zeros((0, 4), dtype=np.int32))
    return out


def ghost_capacity(all_ghosts, size):
    """Symmetric ghost-zone sizes: maxima over the ranks (ghost_lo even)."""
    glo, ghi = [0, 0], [0, 0]
    for g in range(size):
        for space in (N_SPACE, M_SPACE):
            below, above = all_ghosts[g][space]
            glo[space] = max(glo[space], len(below))
            ghi[space] = max(ghi[space], len(above))
    glo = [(x + 1) & ~1 for x in glo]
    return glo, ghi


def apply_local(L, x_own, ghosts_values_lo, ghosts_values_hi):
    """Host model of a distributed product: ``L`` has local column numbering;
    negative / beyond-own columns index the ghost values (used by the CPU
    tests to check the partitioning against the global product)."""
    own = len(x_own)
    ext = np.concatenate([ghosts_values_lo, x_own, ghosts_values_hi])
    shift = len(ghosts_values_lo)
    E = sps.csr_matrix((L.data, L.indices.astype(np.int64) + shift, L.indptr),
                       shape=(L.shape[0], own + shift + len(ghosts_values_hi)))
    return E.dot(ext)


# ---------------------------------------------------------------------------
# device side
# ---------------------------------------------------------------------------
class DistSolver:
    """One rank of a row-partitioned projected-CG solve.

    ``group`` is a ``torch.distributed`` process group able to move Python
    objects (gloo); it is used ONLY for setup (IPC handles, ghost lists) and
    for max-over-ranks timing -- the data path inside the kernels is NVLink
    peer memory, no NCCL call.
    """

    def __init__(self, H, A, rank, size, group=None, device=None, inner_rtol=1e-13,
                 inner_maxit=1000, orth_tol=1e-12, max_refin=3, explicit_aat=True,
                 generated=None, band=True, band_chunk_rows=0):
        """``H``, ``A``: the global matrices (every rank holds them), or None with
        ``generated=dict(n=..., kappa=..., seed=...)``: the index-addressable
        synthetic QP, each rank generating only a window around its own rows."""
        import torch.distributed as dist
        from . import device as dev
        self.rank, self.size = int(rank), int(size)
        if generated is not None:
            n = int(generated["n"])
            m = n // 2
            self.part = RowPartition(n, m, size)
            parts = local_parts_generated(self.part, rank, generated.get("kappa", 1e4),
                                          generated.get("seed", 0))
            fro2 = [None] * size
            dist.all_gather_object(fro2, parts["fro2_own"], group=group)
            parts["fro"] = float(np.sqrt(sum(fro2)))
        else:
            n, m = H.shape[0], A.shape[0]
            self.part = RowPartition(n, m, size)
            parts = local_parts(self.part, rank, H, A)
        self.parts = parts
        all_ghosts = [None] * size
        dist.all_gather_object(all_ghosts, parts["ghosts"], group=group)
        glo, ghi = ghost_capacity(all_ghosts, size)
        sends = send_lists(self.part, rank, all_ghosts, glo)
        self.ctx = ctx = dev.Context(device)
        lib = ctx.lib
        I64x2 = C.c_int64 * 2
        own = I64x2(self.part.own_len(rank, 0), self.part.own_len(rank, 1))
        own_max = I64x2(self.part.own_max(0), self.part.own_max(1))
        stride = [glo[k] + own_max[k] + ghi[k] + 64 for k in (0, 1)]
        # solver workspaces (7 n + 1 m), projector (1 n + 7 m), the SQP step (13 n + 4 m), helper
        # vectors of the outer loop and the caller's own: room for 48 n- and 32 m-vectors
        arena = 8 * (48 * stride[0] + 32 * stride[1]) + (1 << 20)
        handle = (C.c_char * 64)()
        mine = parts["ghosts"]
        ctx.check(lib.ipb_comm_init(ctx.h, self.rank, self.size, own, own_max, I64x2(*glo),
                                    I64x2(*ghi), I64x2(len(mine[0][0]), len(mine[1][0])),
                                    I64x2(len(mine[0][1]), len(mine[1][1])), n - m, arena, handle))
        handles = [None] * size
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        buf = (C.c_char * (64 * size)).from_buffer_copy(b"".join(handles))
        dist.barrier(group=group)
        ctx.check(lib.ipb_comm_connect(ctx.h, buf))
        for space in (N_SPACE, M_SPACE):
            e = np.ascontiguousarray(sends[space], dtype=np.int32)
            ctx.check(lib.ipb_comm_set_send_list(ctx.h, space, e.shape[0],
                                                 e.ctypes.data_as(C.c_void_p)))
        self.n_own, self.m_own = own[0], own[1]
        mk = lambda M_, space: self._csr(M_, space)        # noqa: E731
        self.A_raw, self.A_norm = mk(parts["A_raw"], 0), mk(parts["A_norm"], 0)
        self.At_raw, self.At_norm = mk(parts["At_raw"], 1), mk(parts["At_norm"], 1)
        self.M = mk(parts["M"], 1) if explicit_aat else None
        self.Hcsr = mk(parts["H"], 0)
        self.H = dev.DeviceHessian(ctx, self.n_own, self.n_own, csr=self.Hcsr)
        ph = C.c_void_p()
        ds = np.ascontiguousarray(parts["dscale"], dtype=np.float64)
        ctx.check(lib.ipb_proj_create_parts(
            ctx.h, self.A_raw.h, self.A_norm.h, self.At_raw.h, self.At_norm.h,
            self.M.h if self.M is not None else None, ds.ctypes.data_as(C.POINTER(C.c_double)),
            parts["fro"], float(orth_tol), int(max_refin), float(inner_rtol), int(inner_maxit),
            C.byref(ph)))
        self.proj = _ProjHandle(ctx, ph, self.m_own, self.n_own)
        # banded direct solve: every rank must take the same decision (the sweeps
        # exchange halos), so it is used only if every rank can
        can = bool(band and explicit_aat and parts["band"] is not None)
        votes = [None] * size
        dist.all_gather_object(votes, can or self.m_own == 0, group=group)
        self.band = False
        if all(votes) and self.m_own > 0:
            bd = parts["band"]
            self.M_ext = dev.DeviceCSR(ctx, bd["M_ext"], with_transpose=False)
            ctx.check(lib.ipb_proj_set_band_parts(
                ctx.h, ph, self.M_ext.h, int(bd["ext_off"]), int(bd["ghost_below"]),
                int(bd["ghost_above"]), int(band_chunk_rows), BAND_HALO))
            info = (C.c_int32 * 6)()
            ctx.check(lib.ipb_proj_band_info(ctx.h, ph, info))
            self.band = bool(info[0])
        used = [None] * size
        dist.all_gather_object(used, self.band, group=group)
        if len(set(used)) != 1:
            raise RuntimeError("ranks disagree on the banded solve: %s" % used)
        dist.barrier(group=group)
        self.group = group

    def _csr(self, M_, space):
        from . import device as dev
        cols = self.n_own if space == N_SPACE else self.m_own
        d = dev.DeviceCSR(self.ctx, M_, with_transpose=False, local_columns=True, cols=cols)
        self.ctx.check(self.ctx.lib.ipb_csr_set_space(self.ctx.h, d.h, space))
        return d

    def local(self, v, space):
        a, b = self.part.own(self.rank, space)
        return np.ascontiguousarray(v[a:b], dtype=np.float64)

    def vector(self, data=None, space=N_SPACE):
        from . import device as dev
        length = self.n_own if space == N_SPACE else self.m_own
        return dev.DeviceVector(self.ctx, data=data, length=None if data is not None else length)

    def projected_cg(self, c, b, x, **kw):
        """``c``, ``b``, ``x``: DeviceVectors holding this rank's blocks."""
        from . import device as dev
        return dev.projected_cg_dev(self.proj, self.H, c, b, x, **kw)

    def gather(self, x_local, space=N_SPACE):
        import torch.distributed as dist
        pieces = [None] * self.size
        dist.all_gather_object(pieces, np.asarray(x_local), group=self.group)
        return np.concatenate(pieces)

    def allgather(self, local, space=N_SPACE):
        """The full vector from the ranks' own blocks, on every rank (blocks padded
        to the largest one; gloo moves host tensors, nccl device tensors)."""
        import torch
        import torch.distributed as dist
        cap = max(self.part.own_max(space), 1)
        buf = np.zeros(cap)
        buf[:len(local)] = local
        t = torch.from_numpy(buf)
        if dist.get_backend(self.group) == "nccl":
            t = t.cuda(self.ctx.device)
        out = [torch.empty_like(t) for _ in range(self.size)]
        dist.all_gather(out, t, group=self.group)
        off = self.part.offsets(space)
        return np.concatenate([o.cpu().numpy()[:off[g + 1] - off[g]] for g, o in enumerate(out)])

    def allmax(self, value):
        import torch
        import torch.distributed as dist
        t = torch.tensor([float(value)], dtype=torch.float64)
        if dist.get_backend(self.group) == "nccl":
            t = t.cuda(self.ctx.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def _scratch(self, space, k):
        """Helper device vectors of the outer loop (allocated once, in the same
        order on every rank: they live in the symmetric arena)."""
        pool = self.__dict__.setdefault("_scratch_pool", {})
        for j in range(k + 1):                       # same allocation order on every rank
            if (space, j) not in pool:
                pool[(space, j)] = self.vector(space=space)
        return pool[(space, k)]

    def apply_map(self, which, x_full):
        """``Z.dot`` / ``LS.dot`` / ``Y.dot`` on a full host vector (every rank
        passes the same one and gets the full result)."""
        from .device import ProjInfo
        lib, ctx = self.ctx.lib, self.ctx
        fn, sin, sout = {"Z": (lib.ipb_proj_null_space_dev, N_SPACE, N_SPACE),
                         "LS": (lib.ipb_proj_least_squares_dev, N_SPACE, M_SPACE),
                         "Y": (lib.ipb_proj_row_space_dev, M_SPACE, N_SPACE)}[which]
        vin = self._scratch(sin, 0)
        vout = self._scratch(sout, 1)
        vin.upload(self.local(x_full, sin))
        info = ProjInfo()
        ctx.check(fn(ctx.h, self.proj.h, vin.h, vout.h, C.byref(info)))
        return self.allgather(vout.download(), sout)

    def box_intersections(self, z, d, lb, ub, entire_line=False):
        """``box_intersections`` (qp_subproblem.py:152-234) on full host vectors,
        evaluated block-wise on the ranks and reduced inside the kernel."""
        lib, ctx = self.ctx.lib, self.ctx
        vs = [self._scratch(N_SPACE, k) for k in range(4)]
        for v, full in zip(vs, (z, d, lb, ub)):
            v.upload(self.local(np.asarray(full, dtype=np.float64), N_SPACE))
        ta, tb, hit = C.c_double(), C.c_double(), C.c_int()
        ctx.check(lib.ipb_box_sphere_intersections_dev(
            ctx.h, vs[0].h, vs[1].h, vs[2].h, vs[3].h, float("inf"), int(bool(entire_line)),
            C.byref(ta), C.byref(tb), C.byref(hit)))
        return ta.value, tb.value, bool(hit.value)

    def hot_path(self, H=None, A=None):
        """The L1 functions of ``equality_constrained_sqp`` for a row-partitioned
        problem with CONSTANT matrices (a QP: BASELINE configs 2 and 5): pass it
        as ``minimize_constrained(..., hot_path=solver.hot_path(H, A))`` on every
        rank.  Every rank runs the same host loop on full host vectors (SPMD, the
        callbacks are the user's); each device step works on the rank's own
        blocks -- ``ipb_sqp_normal_step`` / ``ipb_sqp_tangential_step`` /
        ``ipb_sqp_multipliers`` on the partitioned context -- and the results are
        all-gathered, so all scalars are the numpy expressions of the single-GPU
        loop on identical data: same decisions on every rank, and the same
        iteration path as one GPU.  ``H`` / ``A`` (optional): the global matrices
        the solver was built from, to reject a problem whose matrices change."""
        from types import SimpleNamespace
        from .equality_constrained_sqp import ZETA_NORMAL, BOX_NORMAL
        solver = self

        class _Map:
            def __init__(self, which, shape):
                self.which, self.shape, self.projector = which, shape, solver.proj

            def dot(self, x):
                return solver.apply_map(self.which, np.asarray(x, dtype=np.float64))

        class _Hessian:
            """Marker for the partitioned (constant) Hessian; ``dot`` for callers
            that want a product on the host is not offered on purpose."""
            def __init__(self, n, terms):
                self.shape, self.terms = (n, n), terms

        def same_matrix(X, ref):
            if ref is None or X is ref:
                return True
            X, ref = sps.csr_matrix(X), sps.csr_matrix(ref)
            return (X.shape == ref.shape and X.nnz == ref.nnz
                    and np.array_equal(X.indptr, ref.indptr)
                    and np.array_equal(X.indices, ref.indices) and np.array_equal(X.data, ref.data))

        def projections(A_now, method=None):
            if not same_matrix(A_now, A):
                raise NotImplementedError("row-partitioned runs take a constant Jacobian "
                                          "(linear constraints); this one changed")
            n, m = solver.part.n, solver.part.m
            return _Map("Z", (n, n)), _Map("LS", (m, n)), _Map("Y", (n, m))

        def hessian(n, terms):
            if len(terms) != 1 or not same_matrix(terms[0], H):
                raise NotImplementedError("row-partitioned runs take a constant Hessian (a QP)")
            return _Hessian(n, terms)

        def fused_step(Hop, Z, c, b, radius, lb, ub, bounds_changed):
            loc = solver.local
            dn = solver.allgather(solver.proj.sqp_normal_step(
                None if c is None else loc(c, N_SPACE), None if b is None else loc(b, M_SPACE),
                ZETA_NORMAL * radius, None if lb is None else loc(lb, N_SPACE),
                None if ub is None else loc(ub, N_SPACE), bounds_changed, BOX_NORMAL), N_SPACE)
            radius_t = np.sqrt(radius**2 - np.linalg.norm(dn)**2)
            d, Hd, Ad, cg_info = solver.proj.sqp_tangential_step(solver.H, radius_t)
            return (dn, solver.allgather(d, N_SPACE), solver.allgather(Hd, N_SPACE),
                    solver.allgather(Ad, M_SPACE), cg_info)

        def fused_multipliers(Z, c):
            v, opt = solver.proj.sqp_multipliers(solver.local(c, N_SPACE))
            return solver.allgather(v, M_SPACE), solver.allmax(opt)

        def unavailable(*a, **k):
            raise NotImplementedError("the row-partitioned hot path runs the fused SQP step only")

        return SimpleNamespace(projections=projections, hessian=hessian, fused_step=fused_step,
                               fused_multipliers=fused_multipliers,
                               can_fuse=lambda Hop: isinstance(Hop, _Hessian),
                               box_intersections=solver.box_intersections,
                               projected_cg=unavailable, modified_dogleg=unavailable,
                               jac_dot=unavailable, jac_tdot=unavailable)


from .device import DeviceProjector as _DeviceProjector  # noqa: E402


class _ProjHandle(_DeviceProjector):
    """Projector of one rank (created by ``ipb_proj_create_parts``): the *_dev
    calls and the SQP-step calls of :class:`DeviceProjector` on the rank's own
    blocks (``n``, ``m`` are the own lengths)."""

    keep_given_bounds = True

    def __init__(self, ctx, h, m, n):
        self.ctx, self.h, self.m, self.n = ctx, h, int(m), int(n)
        self.A = None
        self.last_info = None
        self.totals = dict(inner_iters=0, n_solves=0, n_refinements=0, applies=0,
                           n_fallbacks=0)
