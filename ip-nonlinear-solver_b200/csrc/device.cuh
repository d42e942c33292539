// Device side of the B200-native projected-CG / null-space-projection path.
//
// Design:
//   * every operation of the path -- a projector application, a whole
//     projected-CG solve, a dogleg step -- is ONE persistent cooperative
//     kernel: the grid is sized to what is co-resident on the 148 SMs, phases
//     are grid-stride loops separated by a grid barrier, and every scalar the
//     control flow needs (alpha, beta, norms, intersection parameters,
//     convergence flags) is reduced on the device in a fixed order and read
//     by every CTA, so all CTAs take identical branches and nothing but a
//     small result struct returns to the host;
//   * SpMV: a CTA owns a contiguous range of row blocks; the val / col /
//     rowptr slices of a block and the slice of x it addresses land in shared
//     memory by 1-D TMA bulk copies, then one thread per row sums its row in
//     index order (the left-to-right order of scipy's csr_matvec, so with
//     -fmad=false row sums are bit-identical to the reference's);
//   * (A A')^-1: when M = A~ A~' (A~ the row-normalised A) is banded, a
//     banded Cholesky factor applied in overlapping chunks (one chunk per
//     thread, its factor rows streamed through per-warp TMA rings) is a
//     direct solve to round-off, checked by the residual and refined like the
//     reference's own refinement loop; otherwise (or if that check fails) a
//     single-reduction (Chronopoulos-Gear) CG on M: one SpMV (dot products
//     fused in its epilogue) and one fused TMA-pipelined 4-vector update per
//     iteration.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace ipb {

#ifndef IPB_THREADS
#define IPB_THREADS 256
#endif
constexpr int THREADS = IPB_THREADS;  // threads per CTA of every kernel
constexpr int WARPS = THREADS / 32;
#ifndef IPB_CHUNK
#define IPB_CHUNK 2048
#endif
#ifndef IPB_RMAX
#define IPB_RMAX 512
#endif
#ifndef IPB_NSTAGE
#define IPB_NSTAGE 2
#endif
#ifndef IPB_MIN_BLOCKS
#define IPB_MIN_BLOCKS 3
#endif
#ifndef IPB_XMAX
#define IPB_XMAX 1152
#endif
constexpr int XMAX = IPB_XMAX;        // doubles of the x slice staged with a block
constexpr int CHUNK = IPB_CHUNK;      // nnz staged per pipeline stage
constexpr int RMAX = IPB_RMAX;        // max rows of one row block
constexpr int NSTAGE = IPB_NSTAGE;    // pipeline stages per CTA
constexpr int PAD = 8;                // over-read slack of the bulk copies (elements)
constexpr int MIN_BLOCKS = IPB_MIN_BLOCKS;   // CTAs per SM the kernels are compiled for
constexpr int NSLOT = 8;              // reduction slots per barrier
constexpr int MAX_GRID = 2048;        // upper bound on CTAs of a launch
static_assert(CHUNK % 4 == 0 && RMAX % 32 == 0 && XMAX % 2 == 0 && NSTAGE >= 2, "block shape");

enum { OP_SUM = 0, OP_MIN = 1, OP_MAX = 2 };

// ---------------------------------------------------------------- structs
struct CsrView {
    int rows, cols;
    int space;                // vector space of the columns (0: n, 1: m) -- selects the halo list
    int nblk;                 // row blocks (built at upload, <= CHUNK nnz each)
    int xform_ok;             // square, every block stages its x slice, owns its rows' columns and has
                              // room for a second slice: x may be formed on the fly (spmv xform)
    const int4 *desc;         // per block {r0, r1, p0, len}, {c0, xcnt, -, -}
    const int *rowptr;
    const int *col;
    const double *val;
};

struct HessView {
    int n, nx;                // H = blockdiag(S + D (nx x nx), diag(tail))
    int has_csr, has_dense, has_tail;
    CsrView S;
    const double *D;          // row-major nx x nx
    const double *tail;       // n - nx
};

constexpr int MAX_RANKS = 8;

// Multi-GPU state of a row-partitioned solve (one process per GPU; every
// pointer below that belongs to a peer is a CUDA-IPC mapping of the peer's
// allocation, reached over NVLink).  size == 1: single GPU, nothing is used.
//   * all solver vectors of a rank live in one symmetric arena (same offsets
//     on every rank), laid out [ghost_lo | own | ghost_hi]; a matrix's column
//     indices are relative to the own part (negative / beyond the end = ghost);
//   * before an SpMV the owners send the entries their peers hold as ghosts
//     into the peers' ghost mailboxes (halo_sync); no barrier between ranks;
//   * reductions: every rank posts its partial sums into every peer's mailbox
//     and adds the mailbox up in rank order -- bit-identical on all ranks.
struct SendList {
    const int4 *ent;          // {src index (own), peer, dst index on the peer (own-relative), mailbox slot}
    int count;
};
struct CommView {
    int rank, size;
    double *arena;            // this rank's arena
    double *peer_arena[MAX_RANKS];
    double *mbox;             // [4][MAX_RANKS][NSLOT] in this rank's memory
    double *peer_mbox[MAX_RANKS];
    unsigned *seq;            // sequence counter of this rank (persists across launches)
    SendList send[2];         // 0: n-space vectors, 1: m-space vectors
    // ghost mailboxes ("LL": 2 words {sequence, 32 bits} per entry, 4 generations)
    unsigned long long *gmbox[2];                  // this rank's, per space
    unsigned long long *peer_gmbox[2][MAX_RANKS];
    int gcap[2], gcap_lo[2];  // slots per generation; slots below gcap_lo are ghost_lo entries
    int n_lo[2], n_hi[2];     // ghost entries this rank really holds
    int own_len[2];
    unsigned *ghost_flag;     // [2] local: sequence number of the last ghost zone filled
};

struct SyncView {
    unsigned *barrier;        // sense-reversing arrival counter
    double *partials;         // [2][NSLOT][MAX_GRID], double-buffered by barrier parity
    int *abort_flag;          // set when a spin times out
    unsigned long long *trace;  // optional: [0] = count, then the %globaltimer of every
    int trace_cap;              // grid barrier as CTA 0 leaves it (phase timeline, ns)
    CommView C;
};

// Banded Cholesky factor of M = A~ A~' (half-bandwidth < HB), applied in
// overlapping chunks: chunk c owns rows [c*B, (c+1)*B) and solves rows
// [c*B - ov, (c+1)*B + ov) from a zero start.  M^-1 (and L^-1) of these
// matrices decay so fast that ov = 48 rows reproduce the global solve to
// round-off (measured: residual 1.7e-16), which makes the sweep a direct
// solver; the callers check the residual and refine / fall back to CG, so
// correctness never depends on the decay.  One thread per chunk, 32 chunks
// (a "group") per warp; the factor rows of a group are stored interleaved,
// [step][k][lane], so that a step of a warp is one contiguous 2 KB block.
constexpr int HB = 8;             // doubles per factor row: 1/L_ii, L[i][i-1..i-7]
static_assert(HB == 8, "k_band_interleave assumes 256 doubles per step");
struct BandView {
    int enabled;
    int B, ov, steps;             // owned rows per chunk, overlap, steps = B + 2 ov (multiples of 4)
    int ngroup;                   // groups of 32 chunks
    int in_lo, in_hi;             // rows of a sweep's input that exist, own-relative: [0, m) on one
                                  // GPU; a row-partitioned rank also reads up to ov ghost rows of its
                                  // neighbours on each side (rows outside read as 0)
    const double *L;              // [ngroup][steps][HB][32]
    double *y;                    // [ngroup][B + ov][32]: forward result kept for the backward sweep
    int *ok;                      // 1 while the factor is trusted (0: breakdown / no contraction seen)
};

struct ProjView {
    int m, n;
    BandView band;
    CsrView A;                // row-normalised A  (At = its transpose)
    CsrView At;
    CsrView M;                // explicit A~ A~' (m x m), valid when has_M
    int has_M;
    const double *Araw;       // un-normalised values of A (same pattern)
    const double *dscale;     // 1/||a_i|| (0 for empty rows)
    const double *dnorm2;     // ||a_i||^2
    double fro;               // ||A||_F of the un-normalised A
    double orth_tol;
    int max_refin;
    double rtol;
    int maxit;
    // work vectors
    double *v, *r, *w, *p, *s, *rhs;   // m
    double *q;                          // m: intermediate of the squared operator (singular A A')
    double *t;                          // n
};

struct ProjStats {
    long long inner_iters, n_solves, n_refinements;
    int status;
    double orth, max_orth;
    int band_ok;                  // banded direct solve usable (same value in every CTA)
    long long n_sweeps, n_fallbacks;
    long long n_singular;         // solves redone as least-squares problems (singular A A')
};

__device__ __forceinline__ ProjStats proj_stats_init(const ProjView &P) {
    ProjStats st;
    st.inner_iters = st.n_solves = st.n_refinements = 0;
    st.status = 0; st.orth = st.max_orth = 0.0;
    st.n_sweeps = st.n_fallbacks = st.n_singular = 0;
    // read before any CTA of this launch can have cleared it (the first write
    // is behind at least one grid barrier)
    st.band_ok = (P.band.enabled && P.m > 0) ? *reinterpret_cast<volatile int *>(P.band.ok) : 0;
    return st;
}

// ---------------------------------------------------------------- loads
__device__ __forceinline__ double ldg_stream(const double *p) {
    return __ldg(p);
}
__device__ __forceinline__ int ldg_stream(const int *p) { return __ldg(p); }

// 128-bit load / store that do not allocate in L1: the streaming passes run
// next to ~216 KB of shared memory per SM, which leaves almost no L1, and an
// allocating load then waits for a free line (measured: 80 % instead of 84 %
// of the copy bandwidth for the fused CG update at this carve-out).
__device__ __forceinline__ double2 ld_stream2(const double2 *p) {
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
                 : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_stream2(double2 *p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};"
                 ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// ---------------------------------------------------------------- barrier
// Same protocol as cooperative_groups::grid_group::sync (release add by the
// CTA master, acquire poll), written out so the reduction and -- later -- the
// cross-GPU exchange can ride on it.  Requires all CTAs co-resident
// (cudaLaunchCooperativeKernel).
struct Smem;
__device__ __forceinline__ void bump_phase(Smem &sm);
__device__ __forceinline__ void grid_sync(const SyncView &S, Smem &sm) {
    // vectors written in this phase are read by TMA bulk copies (async proxy)
    // in the next one: order this thread's generic-proxy writes before them
    asm volatile("fence.proxy.async.global;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned nb = 1;
        if (blockIdx.x == 0) nb = 0x80000000u - (gridDim.x - 1);
        unsigned old, cur;
        asm volatile("atom.add.release.gpu.u32 %0,[%1],%2;"
                     : "=r"(old) : "l"(S.barrier), "r"(nb) : "memory");
        long long spins = 0;
        do {
            asm volatile("ld.acquire.gpu.u32 %0,[%1];"
                         : "=r"(cur) : "l"(S.barrier) : "memory");
            if (++spins > (1ll << 31)) { *S.abort_flag = 1; __trap(); }
        } while (((old ^ cur) & 0x80000000u) == 0);
        bump_phase(sm);
        if (S.trace && blockIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            const unsigned long long k = S.trace[0];
            if ((long long)k + 1 < S.trace_cap) S.trace[k + 1] = t;
            S.trace[0] = k + 1;
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ double comb(double a, double b, int op) {
    if (op == OP_SUM) return a + b;
    if (op == OP_MIN) return fmin(a, b);
    return fmax(a, b);
}
__device__ __forceinline__ double ident(int op) {
    if (op == OP_SUM) return 0.0;
    if (op == OP_MIN) return INFINITY;
    return -INFINITY;
}

__device__ __forceinline__ double warp_reduce(double v, int op) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = comb(v, __shfl_xor_sync(0xffffffffu, v, o), op);
    return v;
}

// The per-CTA partials are double-buffered by the parity of the grid-barrier
// count: publish() writes the buffer of the current phase, collect() -- always
// called right after the barrier that ends the phase the values were published
// in -- reads that buffer.  A CTA that leaves the barrier early and publishes
// its next partial therefore never overwrites a value a slower CTA is still
// collecting (it can be at most one barrier ahead).  RULE: a published value is
// collected after the NEXT grid_sync, never a later one; a partial that has to
// survive another phase stays in a register and is published in that phase.
__device__ __forceinline__ double *partial_buf(const SyncView &S, unsigned phase) {
    return S.partials + (size_t)(phase & 1u) * (NSLOT * MAX_GRID);
}
// ---------------------------------------------------------------- shared memory
// One pipeline stage = the val / col / rowptr slices of one row block, landed
// by three 1-D TMA bulk copies (cp.async.bulk) that complete on the stage's
// ``full`` mbarrier.  NSTAGE stages per CTA; a warp that has finished a block
// bumps the stage's ``done`` counter and the last one to do so refills the
// stage with the block NSTAGE further on -- the warps of a CTA run decoupled
// (no __syncthreads in the block loop) and NSTAGE - 1 copies stay in flight.
struct __align__(128) Stage {
    double val[CHUNK + PAD];
    double xs[XMAX];              // x[c0 .. c0 + xcnt) of the block, when it fits
    int col[CHUNK + PAD];
    int rowptr[RMAX + PAD];
};

struct __align__(128) Smem {
    Stage st[NSTAGE];
    unsigned long long full[NSTAGE];
    unsigned done[NSTAGE];        // warps that have finished with the stage's block
    unsigned seq;                 // stage uses of this CTA so far (slot = seq % NSTAGE)
    unsigned gseq;                // cross-GPU sequence number (multi-GPU only)
    unsigned ghost_seq;           // sequence number the ghost zone of the current product carries
    unsigned phase;               // grid barriers passed so far (parity selects the partials buffer)
    unsigned long long bfull[8];          // banded sweep: "full" mbarriers of the 8 block buffers
    unsigned long long bempty[8];         // ... and their "released" mbarriers
    unsigned bseq[WARPS];         // ring uses of each warp so far
    double red[WARPS];
    double out[NSLOT];
    int ops[NSLOT];
};

__device__ __forceinline__ void bump_phase(Smem &sm) { sm.phase += 1u; }

// Block-reduce ``v`` and publish it as this CTA's partial of ``slot``.
// Fixed tree => deterministic.
__device__ __forceinline__ void publish(const SyncView &S, Smem &sm, int slot, double v,
                                        int op) {
    v = warp_reduce(v, op);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm.red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = sm.red[0];
#pragma unroll
        for (int w = 1; w < WARPS; ++w) a = comb(a, sm.red[w], op);
        partial_buf(S, sm.phase)[slot * MAX_GRID + blockIdx.x] = a;
    }
}

// After a grid_sync: every CTA combines all CTAs' partials of slots
// [0, nslots) in the same fixed order, so every thread of the grid sees
// bit-identical values.
__device__ __forceinline__ void collect_local(const SyncView &S, Smem &sm, int nslots) {
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double *part = partial_buf(S, sm.phase - 1u);
    for (int slot = warp; slot < nslots; slot += WARPS) {
        int op = sm.ops[slot];
        double a = ident(op);
        for (int i = lane; i < (int)gridDim.x; i += 32)
            a = comb(a, __ldcg(&part[slot * MAX_GRID + i]), op);
        a = warp_reduce(a, op);
        if (lane == 0) sm.out[slot] = a;
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned smem_addr(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void pipeline_init(Smem &sm, const SyncView &S) {
    if (threadIdx.x == 0) {
        sm.gseq = S.C.size > 1 ? *reinterpret_cast<volatile unsigned *>(S.C.seq) : 0u;
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&sm.full[s])));
            sm.done[s] = 0;
        }
        sm.seq = 0;
        sm.phase = 0;
        for (int w = 0; w < WARPS; ++w) sm.bseq[w] = 0;
        for (int q = 0; q < 8; ++q) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&sm.bfull[q])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&sm.bempty[q])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes,
                                         unsigned long long *mbar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(mbar)) : "memory");
}

__device__ __forceinline__ bool mbar_try(unsigned long long *mbar, unsigned parity) {
    unsigned done;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done) : "r"(smem_addr(mbar)), "r"(parity) : "memory");
    return done != 0;
}

__device__ __forceinline__ void mbar_wait(unsigned long long *mbar, unsigned parity,
                                          int *abort_flag) {
    long long spins = 0;
    while (!mbar_try(mbar, parity))
        if (++spins > (1ll << 26)) { *abort_flag = 2; __trap(); }
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(mbar)) : "memory");
}

// ---------------------------------------------------------------- cross-GPU
// Cross-GPU all-reduce of sm.out[0..nslots) in rank order, "LL" style: every
// value travels as two 8-byte words {sequence number, 32 bits of the double},
// so a word is its own arrival flag (8-byte stores are atomic over NVLink)
// and no fence or separate flag is needed: one NVLink hop.  CTA 0 posts this
// rank's values into every peer's mailbox; EVERY CTA polls its own rank's
// mailbox and adds the ranks up in the same order -- bit-identical everywhere.
// Mailbox generation = seq & 3 (a rank is never two exchanges ahead of a peer).
__device__ __forceinline__ void rank_allreduce(const SyncView &S, Smem &sm, int nslots) {
    const CommView &C = S.C;
    const unsigned seq = sm.gseq + 1u;
    const int gen = (int)(seq & 3u);
    const int t = threadIdx.x;
    if (blockIdx.x == 0 && t < nslots * C.size) {
        const int slot = t % nslots, q = t / nslots;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(sm.out[slot]);
        const unsigned long long lo = ((unsigned long long)seq << 32) | (bits & 0xffffffffull);
        const unsigned long long hi = ((unsigned long long)seq << 32) | (bits >> 32);
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(C.peer_mbox[q]) +
                                  2 * (size_t)((gen * MAX_RANKS + C.rank) * NSLOT + slot);
        asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" ::"l"(dst), "l"(lo), "l"(hi) : "memory");
    }
    __syncthreads();                      // everybody has read sm.out of the local collect
    if (t < nslots) {
        const int op = sm.ops[t];
        double a = 0.0;
        for (int q = 0; q < C.size; ++q) {
            const unsigned long long *src = reinterpret_cast<const unsigned long long *>(C.mbox) +
                                            2 * (size_t)((gen * MAX_RANKS + q) * NSLOT + t);
            unsigned long long lo, hi;
            long long spins = 0;
            do {
                asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(src) : "memory");
                if (++spins > (1ll << 28)) { *S.abort_flag = 3; __trap(); }
            } while ((unsigned)(lo >> 32) != seq || (unsigned)(hi >> 32) != seq);
            const double v = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
            a = q == 0 ? v : comb(a, v, op);
        }
        sm.out[t] = a;
    }
    if (t == 0) {
        sm.gseq = seq;
        if (blockIdx.x == 0) *reinterpret_cast<volatile unsigned *>(C.seq) = seq;
    }
    __syncthreads();
}

// After a grid_sync: sm.out[slot] = combination of every CTA's (and every
// rank's) partial of slots [0, nslots), in a fixed order -- bit-identical in
// every thread of every rank.
__device__ __forceinline__ void collect(const SyncView &S, Smem &sm, int nslots) {
    collect_local(S, sm, nslots);
    if (S.C.size > 1) rank_allreduce(S, sm, nslots);
}

// Ghost exchange before a product, called by ALL CTAs with x final on this
// rank (the condition any SpMV needs anyway).  No barrier between the ranks:
//   * the COMMUNICATOR CTA (the last one; it takes no row blocks) sends the
//     entries its peers hold as ghosts into their ghost mailboxes, LL style
//     -- {sequence number, 32 bits} words that are their own arrival flag,
//     one NVLink hop, no fence -- then polls its own mailbox for the entries
//     this rank needs, writes them into x's ghost zone and releases the local
//     flag ghost_flag[space] = sequence number;
//   * every other CTA starts its blocks at once; only a block whose columns
//     reach into a ghost zone (descriptor flag) waits for that flag, and the
//     CTA that owns block 0 walks its range backwards so that this block
//     comes last -- the exchange hides behind the interior blocks.
// Long lists (random sparsity) are sent and received by all CTAs, followed by
// a grid barrier.  Mailbox generation = seq & 3; the ghost zone itself is only
// ever written by its own rank, so a fast peer cannot overwrite values in use.
__device__ __forceinline__ void ll_store(unsigned long long *dst, unsigned seq, double v) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long lo = ((unsigned long long)seq << 32) | (bits & 0xffffffffull);
    const unsigned long long hi = ((unsigned long long)seq << 32) | (bits >> 32);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" ::"l"(dst), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ double ll_wait(const unsigned long long *src, unsigned seq, int *abort_flag) {
    unsigned long long lo, hi;
    long long spins = 0;
    do {
        asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(src) : "memory");
        if (++spins > (1ll << 28)) { *abort_flag = 4; __trap(); }
    } while ((unsigned)(lo >> 32) != seq || (unsigned)(hi >> 32) != seq);
    return __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
}

__device__ __noinline__ void halo_sync(const SyncView &S, Smem &sm, const double *x, int space) {
    const CommView &C = S.C;
    const SendList L = C.send[space];
    const unsigned seq = sm.gseq + 1u;
    const int gen = (int)(seq & 3u);
    const int cap = C.gcap[space], cap_lo = C.gcap_lo[space];
    const int nl = C.n_lo[space], ng = nl + C.n_hi[space];
    double *xw = const_cast<double *>(x);
    const bool small = L.count <= 4 * THREADS && ng <= 4 * THREADS;
    const bool comm = blockIdx.x == gridDim.x - 1;
    int i0, istep;
    if (small) { i0 = comm ? (int)threadIdx.x : 0x7fffffff; istep = THREADS; }
    else { i0 = blockIdx.x * THREADS + threadIdx.x; istep = gridDim.x * THREADS; }
    for (int i = i0; i < L.count; i += istep) {
        const int4 e = __ldg(L.ent + i);
        ll_store(C.peer_gmbox[space][e.y] + 2 * ((size_t)gen * cap + e.w), seq, x[e.x]);
    }
    for (int g = i0; g < ng; g += istep) {
        const int slot = g < nl ? cap_lo - nl + g : cap_lo + (g - nl);
        const double v = ll_wait(C.gmbox[space] + 2 * ((size_t)gen * cap + slot), seq, S.abort_flag);
        xw[slot < cap_lo ? slot - cap_lo : C.own_len[space] + (slot - cap_lo)] = v;
    }
    if (small) {
        if (comm) {
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) {
                asm volatile("st.release.gpu.global.u32 [%0],%1;" ::"l"(C.ghost_flag + space), "r"(seq) : "memory");
                *reinterpret_cast<volatile unsigned *>(C.seq) = seq;
            }
        }
    } else {
        __threadfence();
        grid_sync(S, sm);
        if (comm && threadIdx.x == 0) {
            asm volatile("st.release.gpu.global.u32 [%0],%1;" ::"l"(C.ghost_flag + space), "r"(seq) : "memory");
            *reinterpret_cast<volatile unsigned *>(C.seq) = seq;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { sm.gseq = seq; sm.ghost_seq = seq; }
    __syncthreads();
}

// A block that reads ghost entries may only be staged once they are in place.
__device__ __forceinline__ void ghost_wait(const SyncView &S, const Smem &sm, int space) {
    const unsigned want = sm.ghost_seq;
    unsigned have;
    long long spins = 0;
    do {
        asm volatile("ld.acquire.gpu.global.u32 %0,[%1];" : "=r"(have) : "l"(S.C.ghost_flag + space) : "memory");
        if (++spins > (1ll << 28)) { *S.abort_flag = 5; __trap(); }
    } while ((int)(have - want) < 0);
}

// ---------------------------------------------------------------- SpMV
// Row epilogue: y_r = sa*s + sb*addv[r] (+ y_r when accumulate); optional
// running sums  acc0 += y_r*dotv[r],  acc1 += y_r^2 * (wv ? wv[r] : 1).
struct Epi {
    double *y;
    const double *addv;
    double sa, sb;
    const double *dotv;
    const double *wv;
    const double *post;       // stored value is y_r * post[r] (sums use y_r)
    int want_sq;
    int accumulate;
    int want_psq;             // acc0 += (y_r * post[r])^2 (dotv unused then)
    // x formed on the fly (EK_FAST, matrices with xform_ok): the product is taken
    // with  x' = -xg + xbeta * x  and x' is written to xout for the rows of each
    // block -- the update p = -g + beta p of projected CG fused into H p
    const double *xg;
    double xbeta;
    double *xout;
};

__device__ __forceinline__ Epi epi_plain(double *y) {
    Epi e; e.y = y; e.addv = nullptr; e.sa = 1.0; e.sb = 0.0; e.dotv = nullptr;
    e.wv = nullptr; e.post = nullptr; e.want_sq = 0; e.accumulate = 0; e.want_psq = 0;
    e.xg = nullptr; e.xbeta = 0.0; e.xout = nullptr; return e;
}

// addv[r] and dotv[r] are passed in (fetched together with the gathers, so
// their latency overlaps); the rarely used operands are read here.
__device__ __forceinline__ void row_out(const Epi &e, int r, double s, double addv_r,
                                        double dotv_r, double post_r, double &acc0, double &acc1) {
    double y = s;
    if (e.sa != 1.0) y = e.sa * y;
    if (e.addv) y = y + e.sb * addv_r;
    if (e.accumulate) y = e.y[r] + y;
    if (e.dotv) acc0 += y * dotv_r;
    if (e.want_sq) acc1 += (e.wv ? e.wv[r] * (y * y) : y * y);
    if (e.post) {
        const double yp = y * post_r;
        if (e.want_psq) acc0 += yp * yp;
        e.y[r] = yp;
    } else {
        e.y[r] = y;
    }
}

// Row block descriptor: {r0, r1, p0, len} rows [r0, r1), nnz [p0, p0 + len);
// {c0, xcnt}: the block's columns lie in [c0, c0 + xcnt) (c0, xcnt even) and
// xcnt <= XMAX, or xcnt = 0 when the column span is too wide to stage.
// Blocks are cut at upload so that len <= CHUNK and r1 - r0 <= RMAX; a row
// longer than CHUNK is a block of its own (len > CHUNK).
struct BlockDesc { int r0, r1, p0, len, c0, xcnt, ghost; };

__device__ __forceinline__ BlockDesc load_desc(const int4 *desc, size_t blk) {
    const int4 a = __ldg(desc + 2 * blk), b = __ldg(desc + 2 * blk + 1);
    BlockDesc d; d.r0 = a.x; d.r1 = a.y; d.p0 = a.z; d.len = a.w; d.c0 = b.x; d.xcnt = b.y;
    d.ghost = b.z;
    return d;
}

// The rows are so short that a global load of an epilogue operand (addv[r] or
// post[r]) issued inside the row loop is fully exposed.  Its slice for the
// block's rows is therefore staged too, in whichever array of the stage has
// room behind the block's own data: the tail of xs (blocks with a narrow
// column span: A', M) or the tail of val (blocks with few non-zeros: A).
constexpr int OPND_MAX = RMAX + PAD;            // doubles of an operand slice
__device__ __forceinline__ int opnd_slot(const BlockDesc &d, bool xstage) {
    // 1: tail of xs, 2: tail of val, 0: not staged
    if (d.len > CHUNK) return 0;
    if (xstage && d.xcnt <= XMAX - OPND_MAX) return 1;
    if (((d.p0 + d.len - (d.p0 & ~3) + 3) & ~3) <= CHUNK + PAD - OPND_MAX) return 2;
    return 0;
}
__device__ __forceinline__ const double *opnd_ptr(const Stage &st, int slot) {
    return slot == 1 ? st.xs + (XMAX - OPND_MAX) : st.val + (CHUNK + PAD - OPND_MAX);
}
static_assert(((XMAX - OPND_MAX) * 8) % 16 == 0 && ((CHUNK + PAD - OPND_MAX) * 8) % 16 == 0 &&
              XMAX > OPND_MAX, "operand slots must be 16-byte aligned");

// The calling thread starts the bulk copies of a row block into stage ``s``: the val /
// col / rowptr slices and, when xstage, the slice of x the block's columns
// address -- so the block is then processed entirely out of shared memory.
// Sources are aligned down to 4 elements (16 B for ints, 32 B for doubles)
// and sizes rounded up to 4; the arrays are allocated with PAD slack.  A
// long-row block does not use its stage: a plain arrive completes the phase
// so that the parity of every slot keeps advancing once per use.
__device__ __forceinline__ void stage_issue(const CsrView &M, const BlockDesc d,
                                            const double *x, bool xstage,
                                            Smem &sm, int s, const double *opnd = nullptr,
                                            const double *xg = nullptr) {
    if (d.len <= CHUNK) {
        const int p0a = d.p0 & ~3, r0a = d.r0 & ~3;
        const unsigned cnt = (unsigned)((d.p0 + d.len - p0a + 3) & ~3);
        const unsigned rcnt = (unsigned)((d.r1 - r0a + 1 + 3) & ~3);
        const unsigned xcnt = xstage ? (unsigned)d.xcnt : 0u;
        Stage &st = sm.st[s];
        // the stage was last read through the generic proxy: order those reads
        // before the async-proxy copies.  (x, written through the generic proxy
        // in an earlier phase, was fenced by its writers at the grid barrier;
        // an unqualified fence here would wait for this thread's global stores.)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const int oslot = opnd ? opnd_slot(d, xstage) : 0;
        const unsigned ocnt = oslot ? (unsigned)((d.r1 - r0a + 3) & ~3) : 0u;
        unsigned tx = cnt * 12u + rcnt * 4u + xcnt * 8u + ocnt * 8u;
        if (xg) tx += xcnt * 8u;        // second slice (xform): same columns, in the xs-tail slot
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                     ::"r"(smem_addr(&sm.full[s])), "r"(tx) : "memory");
        if (ocnt) bulk_g2s(const_cast<double *>(opnd_ptr(st, oslot)), opnd + r0a, ocnt * 8u, &sm.full[s]);
        if (xg && xcnt) bulk_g2s(const_cast<double *>(opnd_ptr(st, 1)), xg + d.c0, xcnt * 8u, &sm.full[s]);
        bulk_g2s(st.val, M.val + p0a, cnt * 8u, &sm.full[s]);
        bulk_g2s(st.col, M.col + p0a, cnt * 4u, &sm.full[s]);
        bulk_g2s(st.rowptr, M.rowptr + r0a, rcnt * 4u, &sm.full[s]);
        if (xcnt) bulk_g2s(st.xs, x + d.c0, xcnt * 8u, &sm.full[s]);
    } else {
        mbar_arrive(&sm.full[s]);
    }
}

// Rows of one staged block, one thread per row, single pass.  XS: x comes
// from the staged slice (shared memory) instead of global gathers.  FAST:
// the epilogue is y = sum (plus the optional dot with dotv), the form of the
// two products inside the CG loops.  Entries are added in index order -- the
// left-to-right order of scipy's csr_matvec, so with -fmad=false row sums are
// bit-identical to it.
// Epilogue kinds the products of the solve loops use, compiled without the
// generic epilogue's run-time branches:
//   EK_FAST  y = s (+ acc0 += s * dotv)            H p, M r, A' r
//   EK_AXPY  y = addv - s, acc1 += y^2             g = r - A~' v
//   EK_POST  y = s * post, acc1 += s^2, acc0 += y^2   D^-1/2 (A z)
//   EK_SQ    y = s, acc1 += s^2                    rhs = A~ x
//   EK_ANY   everything else (row_out)
enum { EK_FAST = 0, EK_AXPY = 1, EK_POST = 2, EK_SQ = 3, EK_ANY = 4 };

template <bool XS, int KIND>
__device__ __forceinline__ void block_rows(const Stage &st, const BlockDesc &d,
                                           const double *x, const Epi &e, const double *opnd,
                                           bool xstage, double &acc0, double &acc1) {
    const int p0a = d.p0 & ~3, r0a = d.r0 & ~3;
    constexpr bool FAST = KIND == EK_FAST;
    // staged slice of the epilogue operand (addv, else post), indexed by global row
    const int oslot = (!FAST && opnd) ? opnd_slot(d, xstage) : 0;
    const double *so = oslot ? opnd_ptr(st, oslot) - r0a : nullptr;
    const double *sv = st.val - p0a;          // sv[q], q a global nnz index
    const int *sc = st.col - p0a;
    const int *rp = st.rowptr - r0a;          // rp[r], r a global row index
    const double *xv = XS ? (st.xs - d.c0) : x;   // xv[c], c a global column index
    const bool dot_is_x = XS && (e.dotv == x);
#pragma unroll 1
    for (int r = d.r0 + (int)threadIdx.x; r < d.r1; r += THREADS) {
        int q = rp[r];
        const int qe = rp[r + 1];
        double dvv = 0.0, avv = 0.0;
        if (e.dotv) {
            const unsigned off = (unsigned)(r - d.c0);
            dvv = (dot_is_x && off < (unsigned)d.xcnt) ? st.xs[off] : e.dotv[r];
        }
        if ((KIND == EK_AXPY || KIND == EK_ANY) && e.addv) avv = so ? so[r] : e.addv[r];
        double sum = 0.0;
        // rows are short (1.4 - 5 entries): the compiler's default 16-fold unrolling of
        // this loop costs a cascade of remainder branches per row (~160 warp
        // instructions for a 1-entry row); keep it tight
#pragma unroll 1
        for (; q + 1 < qe; q += 2) {
            const int c0 = sc[q], c1 = sc[q + 1];
            const double v0 = sv[q], v1 = sv[q + 1];
            const double x0 = xv[c0], x1 = xv[c1];
            sum += v0 * x0;
            sum += v1 * x1;
        }
        if (q < qe) sum += sv[q] * xv[sc[q]];
        if (FAST) {
            if (e.dotv) acc0 += sum * dvv;
            e.y[r] = sum;
        } else if (KIND == EK_AXPY) {
            const double y = -sum + avv;          // the generic path's (-1 * s) + 1 * addv, bit for bit
            acc1 += y * y;
            e.y[r] = y;
        } else if (KIND == EK_POST) {
            const double yp = sum * (so ? so[r] : e.post[r]);
            acc1 += sum * sum;
            acc0 += yp * yp;
            e.y[r] = yp;
        } else if (KIND == EK_SQ) {
            acc1 += sum * sum;
            e.y[r] = sum;
        } else {
            row_out(e, r, sum, avv, dvv, (so && !e.addv) ? so[r] : (e.post ? e.post[r] : 1.0),
                    acc0, acc1);
        }
    }
}

// y = M x over the row blocks of M; CTA b takes the contiguous block range
// [b*per, (b+1)*per).  Thread 0 starts the first NSTAGE copies; after that a
// stage is refilled by the last warp that finishes with it.
template <int KIND>
__device__ __noinline__ void spmv_impl(const CsrView M, const double *x, const Epi e,
                                       Smem &sm, const SyncView &S, double &acc0,
                                       double &acc1) {
    const int tid = threadIdx.x;
    __syncthreads();                       // sm.seq of the previous call is visible
    // multi-GPU: the last CTA is the communicator (halo_sync) and takes no blocks
    const bool dist = S.C.size > 1;
    const int workers = (dist && gridDim.x > 1) ? (int)gridDim.x - 1 : (int)gridDim.x;
    const int per = (M.nblk + workers - 1) / workers;
    const int first = (int)blockIdx.x * per;
    const int nmine = (int)blockIdx.x < workers ? min(per, M.nblk - first) : 0;
    if (nmine <= 0) return;
    const unsigned base = sm.seq;
    // the CTA that owns block 0 (the one that reads ghost_lo) walks backwards
    const bool rev = dist && first == 0;
    const int4 *dp = M.desc + 2 * (size_t)first;
#define IPB_BLK(j) (rev ? nmine - 1 - (j) : (j))
    // x slices are staged only when the bulk-copy source is 16-byte aligned
    const bool xstage = (reinterpret_cast<size_t>(x) & 15) == 0;
    // epilogue operand staged with the blocks (never aliases y: a block's slice is
    // read before its rows are written, by the copy this CTA waits for)
    constexpr bool FAST = KIND == EK_FAST;
    const double *opnd = FAST ? nullptr : (e.addv ? e.addv : e.post);
    if (opnd && (reinterpret_cast<size_t>(opnd) & 15) != 0) opnd = nullptr;
    const double *xg = (FAST && M.xform_ok && xstage) ? e.xg : nullptr;
    if (M.space == 2) {
        // measurement only: the row arithmetic without the copy pipeline -- the
        // CTA's first block is staged once and processed nmine times
        const BlockDesc d0 = load_desc(dp, 0);
        if (d0.len > CHUNK) return;
        if (tid == 0) stage_issue(M, d0, x, xstage, sm, base % NSTAGE, opnd, nullptr);
        mbar_wait(&sm.full[base % NSTAGE], (base / NSTAGE) & 1u, S.abort_flag);
        for (int j = 0; j < nmine; ++j) {
            if (xstage && d0.xcnt > 0) block_rows<true, KIND>(sm.st[base % NSTAGE], d0, x, e, opnd, xstage, acc0, acc1);
            else block_rows<false, KIND>(sm.st[base % NSTAGE], d0, x, e, opnd, xstage, acc0, acc1);
        }
        __syncthreads();
        if (tid == 0) sm.seq = base + 1u;
        return;
    }
    if (tid == 0)
        for (int k = 0; k < NSTAGE && k < nmine; ++k) {
            const BlockDesc dk = load_desc(dp, IPB_BLK(k));
            if (dist && dk.ghost) ghost_wait(S, sm, M.space);
            stage_issue(M, dk, x, xstage, sm, (base + k) % NSTAGE, opnd, xg);
        }
    BlockDesc d = load_desc(dp, IPB_BLK(0));
    for (int j = 0; j < nmine; ++j) {
        const BlockDesc dnext = load_desc(dp, IPB_BLK(min(j + 1, nmine - 1)));
        const BlockDesc dfar = load_desc(dp, IPB_BLK(min(j + NSTAGE, nmine - 1)));
        const unsigned u = base + j;
        const int s = u % NSTAGE;
        if (d.len <= CHUNK) {
            mbar_wait(&sm.full[s], (u / NSTAGE) & 1u, S.abort_flag);
            if (FAST && xg) {
                // form the block's x slice in place: x' = -g + beta x; the block's own
                // rows of x' go to xout (blocks own the columns of their rows)
                Stage &stw = sm.st[s];
                const double *gs = opnd_ptr(stw, 1);
                for (int i = tid; i < d.xcnt; i += THREADS) {
                    const double v = -gs[i] + e.xbeta * stw.xs[i];
                    stw.xs[i] = v;
                    const int c = d.c0 + i;
                    if (c >= d.r0 && c < d.r1) e.xout[c] = v;
                }
                __syncthreads();
            }
            if (!e.y) { /* copy pipeline only (bench) */ }
            else if (xstage && d.xcnt > 0) block_rows<true, KIND>(sm.st[s], d, x, e, opnd, xstage, acc0, acc1);
            else block_rows<false, KIND>(sm.st[s], d, x, e, opnd, xstage, acc0, acc1);
        } else {
            // one long row: the whole CTA strides over it, fixed-order tree
            double part = 0.0;
            for (int q = d.p0 + tid; q < d.p0 + d.len; q += THREADS)
                part += ldg_stream(M.val + q) * x[ldg_stream(M.col + q)];
            part = warp_reduce(part, OP_SUM);
            __syncthreads();
            if ((tid & 31) == 0) sm.red[tid >> 5] = part;
            __syncthreads();
            if (tid == 0) {
                double sum = sm.red[0];
#pragma unroll
                for (int w = 1; w < WARPS; ++w) sum += sm.red[w];
                row_out(e, d.r0, sum, e.addv ? e.addv[d.r0] : 0.0,
                        e.dotv ? e.dotv[d.r0] : 0.0, e.post ? e.post[d.r0] : 1.0, acc0, acc1);
            }
        }
        // hand the stage back; the last warp to do so refills it
        __syncwarp();
        if ((tid & 31) == 0) {
            __threadfence_block();
            const unsigned old = atomicAdd(&sm.done[s], 1u);
            if (old == WARPS - 1) {
                atomicExch(&sm.done[s], 0u);
                __threadfence_block();
                if (j + NSTAGE < nmine) {
                    if (dist && dfar.ghost) ghost_wait(S, sm, M.space);
                    stage_issue(M, dfar, x, xstage, sm, s, opnd, xg);
                }
            }
        }
        d = dnext;
    }
#undef IPB_BLK
    __syncthreads();
    if (tid == 0) sm.seq = base + (unsigned)nmine;
}

__device__ __forceinline__ void spmv(const CsrView M, const double *x, const Epi e,
                                     Smem &sm, const SyncView &S, double &acc0,
                                     double &acc1) {
    if (S.C.size > 1) halo_sync(S, sm, x, M.space);
    const bool base = !e.accumulate && !e.wv && !e.dotv;
    if (!e.addv && e.sa == 1.0 && !e.accumulate && !e.want_sq && !e.post && !e.wv)
        spmv_impl<EK_FAST>(M, x, e, sm, S, acc0, acc1);
    else if (base && e.addv && e.sa == -1.0 && e.sb == 1.0 && e.want_sq && !e.post)
        spmv_impl<EK_AXPY>(M, x, e, sm, S, acc0, acc1);
    else if (base && !e.addv && e.sa == 1.0 && e.want_sq && e.post && e.want_psq)
        spmv_impl<EK_POST>(M, x, e, sm, S, acc0, acc1);
    else if (base && !e.addv && e.sa == 1.0 && e.want_sq && !e.post)
        spmv_impl<EK_SQ>(M, x, e, sm, S, acc0, acc1);
    else
        spmv_impl<EK_ANY>(M, x, e, sm, S, acc0, acc1);
}

// Dense row-major GEMV, one warp per row, same epilogue.
__device__ __noinline__ void gemv_dense(const double *D, int nx, const double *x,
                                        const Epi e, double &acc0, double &acc1) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * WARPS + (threadIdx.x >> 5);
    const int nw = gridDim.x * WARPS;
    for (int r = gw; r < nx; r += nw) {
        const double *row = D + (size_t)r * nx;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int c = lane;
        for (; c + 96 < nx; c += 128) {
            s0 += ldg_stream(row + c) * x[c];
            s1 += ldg_stream(row + c + 32) * x[c + 32];
            s2 += ldg_stream(row + c + 64) * x[c + 64];
            s3 += ldg_stream(row + c + 96) * x[c + 96];
        }
        for (; c < nx; c += 32) s0 += ldg_stream(row + c) * x[c];
        double s = warp_reduce((s0 + s1) + (s2 + s3), OP_SUM);
        if (lane == 0)
            row_out(e, r, s, e.addv ? e.addv[r] : 0.0, e.dotv ? e.dotv[r] : 0.0,
                    e.post ? e.post[r] : 1.0, acc0, acc1);
    }
}

// y = H x  with optional  acc0 += y.dotv.  Phases that accumulate into y are
// separated by a grid barrier because row->CTA maps differ between terms.
// H can form its argument on the fly (see Epi::xg): CSR term only (+ diagonal tail).
__device__ __forceinline__ bool hess_can_xform(const HessView &H, const SyncView &S) {
    return H.has_csr && !H.has_dense && H.S.xform_ok && H.nx > 0;
}

// y = H p' with p' = -g + beta p formed on the fly and written to pout (the
// update of projected CG, qp_subproblem.py:628, fused into the product of
// :634);  acc0 += p'.y.  Requires hess_can_xform.
__device__ __noinline__ void hess_apply_xform(const HessView &H, const SyncView &S,
                                              const double *g, const double *p, double beta,
                                              double *pout, double *y, Smem &sm, double &acc0) {
    double dummy = 0.0;
    Epi e = epi_plain(y);
    e.dotv = p;                  // = the staged x: the dot is taken with the transformed slice
    e.xg = g; e.xbeta = beta; e.xout = pout;
    // row-partitioned: the slices reach into the neighbours' rows of BOTH vectors
    // (spmv exchanges the ghosts of p; the ghost entries of p' are formed from
    // them locally, bit-identical to the owner's)
    if (S.C.size > 1) halo_sync(S, sm, g, H.S.space);
    spmv(H.S, p, e, sm, S, acc0, dummy);
    const int gt = blockIdx.x * THREADS + threadIdx.x, nt = gridDim.x * THREADS;
    for (int i = H.nx + gt; i < H.n; i += nt) {
        const double pn = -g[i] + beta * p[i];
        pout[i] = pn;
        const double v = H.has_tail ? H.tail[i - H.nx] * pn : 0.0;
        y[i] = v;
        if (H.has_tail) acc0 += v * pn;
    }
}

__device__ __noinline__ void hess_apply(const HessView &H, const SyncView &S,
                                        const double *x, double *y,
                                        const double *dotv, Smem &sm,
                                        double &acc0) {
    double dummy = 0.0;
    Epi e = epi_plain(y);
    if (H.has_dense) {
        Epi ed = e;
        if (!H.has_csr) ed.dotv = dotv;
        gemv_dense(H.D, H.nx, x, ed, acc0, dummy);
        if (H.has_csr) grid_sync(S, sm);
    }
    if (H.has_csr) {
        Epi es = e;
        es.accumulate = H.has_dense;
        es.dotv = dotv;
        spmv(H.S, x, es, sm, S, acc0, dummy);
    }
    const int gt = blockIdx.x * THREADS + threadIdx.x, nt = gridDim.x * THREADS;
    if (!H.has_dense && !H.has_csr) {
        for (int i = gt; i < H.nx; i += nt) y[i] = 0.0;
    }
    if (H.has_tail) {
        for (int i = H.nx + gt; i < H.n; i += nt) {
            double v = H.tail[i - H.nx] * x[i];
            y[i] = v;
            if (dotv) acc0 += v * dotv[i];
        }
    } else {
        for (int i = H.nx + gt; i < H.n; i += nt) y[i] = 0.0;
    }
}

// ---------------------------------------------------------------- reduction ops
__device__ __forceinline__ void set_ops(Smem &sm, int o0, int o1 = 0, int o2 = 0,
                                        int o3 = 0, int o4 = 0, int o5 = 0,
                                        int o6 = 0, int o7 = 0) {
    if (threadIdx.x == 0) {
        sm.ops[0] = o0; sm.ops[1] = o1; sm.ops[2] = o2; sm.ops[3] = o3;
        sm.ops[4] = o4; sm.ops[5] = o5; sm.ops[6] = o6; sm.ops[7] = o7;
    }
    __syncthreads();
}

// ---------------------------------------------------------------- streaming passes
// Elementwise pass over NV vectors of length ``len`` as a TMA pipeline through
// the same shared-memory stages as the SpMV: tiles of TS entries of the NV
// input vectors come in by bulk copies, ``f(t, e)`` updates entry e of the
// tiles in place (t[k * TS + e] is vector k) and the tiles k with out[k] != 0
// leave by bulk stores (cp.async.bulk.global.shared::cta).  The pass does not
// depend on L1 (almost none is left next to the stages: a register-path
// version of the CG update reached 80 % of the copy bandwidth, this one 90 %)
// and keeps deep queues of reads and writes in flight.  in[k] / out[k] must be
// 16-byte aligned with one entry of slack behind ``len``.
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(smem_addr(src)), "r"(bytes) : "memory");
}

template <int NV>
struct TileIO {
    const double *in[NV];
    double *out[NV];
};

template <int NV>
__device__ __forceinline__ void tile_issue(const TileIO<NV> &io, Smem &sm, int s, int t0, int cnt) {
    constexpr int TS = (int)((sizeof(Stage) / (NV * sizeof(double))) & ~(size_t)31);
    double *buf = reinterpret_cast<double *>(&sm.st[s]);
    const unsigned bytes = (unsigned)(((cnt + 1) & ~1) * 8);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 ::"r"(smem_addr(&sm.full[s])), "r"((unsigned)NV * bytes) : "memory");
#pragma unroll
    for (int k = 0; k < NV; ++k) bulk_g2s(buf + k * TS, io.in[k] + t0, bytes, &sm.full[s]);
}

template <int NV, class F>
__device__ __forceinline__ void tile_pass(Smem &sm, const SyncView &S, int len,
                                          const TileIO<NV> &io, F &&f) {
    constexpr int TS = (int)((sizeof(Stage) / (NV * sizeof(double))) & ~(size_t)31);
    const int tid = threadIdx.x;
    __syncthreads();
    const int ntile = (len + TS - 1) / TS;
    const int per = (ntile + (int)gridDim.x - 1) / (int)gridDim.x;
    const int first = (int)blockIdx.x * per;
    const int nmine = min(per, ntile - first);
    if (nmine <= 0) return;
    const unsigned base = sm.seq;
    if (tid == 0)
        for (int k = 0; k < NSTAGE && k < nmine; ++k) {
            const int t0 = (first + k) * TS;
            tile_issue<NV>(io, sm, (base + k) % NSTAGE, t0, min(TS, len - t0));
        }
    for (int j = 0; j < nmine; ++j) {
        const unsigned u = base + j;
        const int s = u % NSTAGE;
        const int t0 = (first + j) * TS;
        const int cnt = min(TS, len - t0);
        double *buf = reinterpret_cast<double *>(&sm.st[s]);
        mbar_wait(&sm.full[s], (u / NSTAGE) & 1u, S.abort_flag);
        for (int e = tid; e < cnt; e += THREADS) f(buf, e);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            // never write past the own range (the next entry may be a ghost a
            // peer has just pushed): an odd last entry goes out by plain stores
            const unsigned bytes = (unsigned)((cnt & ~1) * 8);
            if (bytes) {
#pragma unroll
                for (int k = 0; k < NV; ++k)
                    if (io.out[k]) bulk_s2g(io.out[k] + t0, buf + k * TS, bytes);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (cnt & 1) {
                const int e = cnt - 1;
#pragma unroll
                for (int k = 0; k < NV; ++k)
                    if (io.out[k]) io.out[k][t0 + e] = buf[k * TS + e];
            }
            if (j + NSTAGE < nmine) {
                // the stores must have read the stage before it is refilled
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                const int n0 = (first + j + NSTAGE) * TS;
                tile_issue<NV>(io, sm, s, n0, min(TS, len - n0));
            }
        }
    }
    if (tid == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // writes complete
        asm volatile("fence.proxy.async.global;" ::: "memory");
        sm.seq = base + (unsigned)nmine;
    }
    __syncthreads();
}

// ---------------------------------------------------------------- fused CG update
// p = r + beta p;  s = w + beta s;  v += alpha p;  r -= alpha s;  returns this
// thread's share of r.r -- the one vector pass of the inner CG iteration
// (5 reads + 4 writes of m-vectors).
__device__ __noinline__ double cg_update(const ProjView &P, Smem &sm, const SyncView &S,
                                         double alpha, double beta) {
    constexpr int TS = (int)((sizeof(Stage) / (5 * sizeof(double))) & ~(size_t)31);
    TileIO<5> io = {{P.r, P.w, P.p, P.s, P.v}, {P.r, nullptr, P.p, P.s, P.v}};
    double gpart = 0.0;
    tile_pass<5>(sm, S, P.m, io, [&](double *t, int e) {
        const double ri = t[e], wi = t[TS + e];
        const double pi = ri + beta * t[2 * TS + e];
        const double si = wi + beta * t[3 * TS + e];
        const double vi = t[4 * TS + e] + alpha * pi;
        const double rn = ri - alpha * si;
        t[2 * TS + e] = pi; t[3 * TS + e] = si; t[4 * TS + e] = vi; t[e] = rn;
        gpart += rn * rn;
    });
    return gpart;
}

// ---------------------------------------------------------------- banded sweep
// 4 consecutive entries of a vector (row % 4 == 0, vectors are 256-byte
// aligned); rows outside [0, m) read as 0.
__device__ __forceinline__ void ld4_guard(const double *v, int row, int lo, int hi, double out[4]) {
    if (row >= lo && row + 3 < hi) {
        const double2 a = *reinterpret_cast<const double2 *>(v + row);
        const double2 b = *reinterpret_cast<const double2 *>(v + row + 2);
        out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = row + q;
            out[q] = (r >= lo && r < hi) ? v[r] : 0.0;
        }
    }
}

// ---- rings of sweep blocks.  One step = the factor rows of a warp's 32
// chunks, [HB][32] doubles = 2 KB; a BLOCK = 4 consecutive steps = 8 KB
// contiguous in HBM, landed by ONE TMA bulk copy on the block's "full"
// mbarrier (a thread issues only ~3 bulk copies per microsecond, so a copy
// per step would cap a warp at 3 steps/us).  Warps 0..3 of a CTA consume
// (thread = chunk), warp w + 4 produces for warp w: it waits for a buffer's
// "released" mbarrier and issues the next copy, so neither the wait nor the
// issue is on the consumer's step-to-step path.  The 8 buffers (64 KB of the
// stages) are shared by the consumer warps that have a group at all: 8, 4 or
// 2 blocks deep.  All addresses are 32-bit shared-window addresses.
constexpr int BAND_STEP = HB * 32;                 // doubles per step
constexpr int BAND_BLK = 4 * BAND_STEP;            // doubles per block
constexpr int BAND_CW = WARPS / 2;                 // consumer warps per CTA

struct BandRing {
    unsigned data;      // shared address of buffer 0 (+ lane * 8 for a consumer)
    unsigned fulls;     // shared address of the "full" mbarrier of buffer 0
    unsigned empties;   // shared address of the "released" mbarrier of buffer 0
    unsigned mask;      // depth - 1 (depth a power of two)
};

__device__ __forceinline__ void mbar_wait_s(unsigned bar, unsigned parity, int *abort_flag) {
    unsigned spins = 0, done;
    do {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 24)) { *abort_flag = 6; __trap(); }
    } while (!done);
}

template <int OFF>
__device__ __forceinline__ double lds_off(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
    return v;
}

struct BandRow { double l0, l1, l2, l3, l4, l5, l6, l7; };

template <int Q>
__device__ __forceinline__ BandRow band_row(unsigned a) {       // step Q of the block at ``a``
    constexpr int O = Q * BAND_STEP * 8;
    BandRow r;
    r.l0 = lds_off<O>(a);        r.l1 = lds_off<O + 256>(a);  r.l2 = lds_off<O + 512>(a);
    r.l3 = lds_off<O + 768>(a);  r.l4 = lds_off<O + 1024>(a); r.l5 = lds_off<O + 1280>(a);
    r.l6 = lds_off<O + 1536>(a); r.l7 = lds_off<O + 1792>(a);
    return r;
}

// y_j = (r_j - sum_k L[j][j-k] y_{j-k}) / L_jj; y_{j-1} enters last, so only one
// multiply-add and one multiply lie between consecutive steps
#define IPB_FWD_STEP(Q)                                                              \
    {                                                                                \
        const BandRow L = band_row<Q>(a);                                            \
        const double pa = __fma_rn(L.l7, y7, __fma_rn(L.l5, y5, L.l3 * y3));         \
        const double pb = __fma_rn(L.l6, y6, __fma_rn(L.l4, y4, L.l2 * y2));         \
        const double t = __fma_rn(-L.l1, y1, rc[Q] - (pa + pb));                     \
        const double y = t * L.l0;                                                   \
        y7 = y6; y6 = y5; y5 = y4; y4 = y3; y3 = y2; y2 = y1; y1 = y;                \
        if (keep) yp[Q * 32] = y;                                                    \
        if (own) ss += rc[Q] * rc[Q];                                                \
    }

// z_j = (y_j - c0) / L_jj, then row j's entries scatter into the rows above it;
// c_k: what the rows already done contribute to row j-k
#define IPB_BWD_STEP(Q)                                                              \
    {                                                                                \
        const BandRow L = band_row<Q>(a);                                            \
        const double z = (yc[Q] - c0) * L.l0;                                        \
        c0 = __fma_rn(L.l1, z, c1);                                                  \
        c1 = __fma_rn(L.l2, z, c2);                                                  \
        c2 = __fma_rn(L.l3, z, c3);                                                  \
        c3 = __fma_rn(L.l4, z, c4);                                                  \
        c4 = __fma_rn(L.l5, z, c5);                                                  \
        c5 = __fma_rn(L.l6, z, c6);                                                  \
        c6 = L.l7 * z;                                                               \
        zo[Q] = z;                                                                   \
    }

// out (+)= L^-T L^-1 in, chunk by chunk (see BandView).  Thread = chunk: a
// forward substitution over the chunk's B + 2 ov rows from a zero start, the
// result of rows >= ov kept in Bd.y, then the backward substitution down to
// the first owned row; only the B owned rows are written.  Returns this
// thread's share of ||in||^2 (owned rows).  ``in`` and ``out`` must not alias.
__device__ __noinline__ double band_sweep(const ProjView &P, const SyncView &S, Smem &sm,
                                          const double *in, double *out, int accumulate) {
    const BandView &Bd = P.band;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = P.m, B = Bd.B, ov = Bd.ov, steps = Bd.steps;
    const int nfb = steps / 4, nbb = (steps - ov) / 4;        // forward / backward blocks of a group
    // consumer warps of this CTA that have a group at all share the 8 buffers
    const int wact = max(1, min(BAND_CW, (Bd.ngroup + (int)gridDim.x - 1) / (int)gridDim.x));
    const int depth = wact == 1 ? 8 : (wact == 2 ? 4 : 2);
    const bool producer = warp >= BAND_CW;
    const int cw = producer ? warp - BAND_CW : warp;             // the ring this warp works on
    const int in_lo = Bd.in_lo, in_hi = Bd.in_hi;
    double ss = 0.0;
    // row-partitioned: the neighbours' rows next to the own block are ghosts of
    // ``in``; only the first and the last group read them and wait for them
    const bool dist = S.C.size > 1;
    if (dist) halo_sync(S, sm, in, 1);
    __syncthreads();                       // the stages are free (previous phase of this CTA is over)
    BandRing R;
    R.data = smem_addr(reinterpret_cast<double *>(&sm.st[0]) + (size_t)cw * (depth * BAND_BLK));
    R.fulls = smem_addr(sm.bfull + cw * depth);
    R.empties = smem_addr(sm.bempty + cw * depth);
    R.mask = (unsigned)depth - 1u;
    unsigned u = sm.bseq[cw];              // blocks this ring has carried so far
    const int gstride = (int)gridDim.x * BAND_CW;
    if (producer) {
        // ---- feed the ring of warp cw: forward blocks in order, then the backward ones
        if (lane == 0 && cw < wact) {
            constexpr unsigned BYTES = BAND_BLK * sizeof(double);
            for (int g = (int)blockIdx.x + (int)gridDim.x * cw; g < Bd.ngroup; g += gstride) {
                const double *Lg = Bd.L + (size_t)g * steps * BAND_STEP;
                for (int k = 0; k < nfb + nbb; ++k, ++u) {
                    const double *src = Lg + (size_t)(k < nfb ? k : (nfb - 1 - (k - nfb))) * BAND_BLK;
                    const unsigned buf = u & R.mask;
                    // round r of a buffer waits for the release of round r-1 (round 0 passes)
                    mbar_wait_s(R.empties + buf * 8u, ((u / (unsigned)depth) & 1u) ^ 1u, S.abort_flag);
                    const unsigned bar = R.fulls + buf * 8u;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                                 ::"r"(bar), "r"(BYTES) : "memory");
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                        ::"r"(R.data + buf * BYTES), "l"(src), "r"(BYTES), "r"(bar) : "memory");
                }
            }
        }
        __syncthreads();
        return 0.0;
    }
    R.data += lane * 8u;
    const unsigned u0 = u;
    for (int g = (int)blockIdx.x + (int)gridDim.x * warp; g < Bd.ngroup; g += gstride) {
        const int e0 = (g * 32 + lane) * B - ov;                       // row of step 0
        double *yp = Bd.y + (size_t)g * (B + ov) * 32 + lane;          // forward result of row e0 + ov
        if (dist && (g == 0 || g == Bd.ngroup - 1)) ghost_wait(S, sm, 1);
        // ================= forward: steps 0 .. steps-1
        {
            double y1 = 0, y2 = 0, y3 = 0, y4 = 0, y5 = 0, y6 = 0, y7 = 0;
            double rn[4];
            ld4_guard(in, e0, in_lo, in_hi, rn);
            for (int jb = 0; jb < steps; jb += 4, ++u) {
                const double rc[4] = {rn[0], rn[1], rn[2], rn[3]};
                if (jb + 4 < steps) ld4_guard(in, e0 + jb + 4, in_lo, in_hi, rn);
                const bool keep = jb >= ov, own = keep && jb < ov + B;
                const unsigned buf = u & R.mask;
                mbar_wait_s(R.fulls + buf * 8u, (u / (unsigned)depth) & 1u, S.abort_flag);
                const unsigned a = R.data + buf * (unsigned)(BAND_BLK * sizeof(double));
                IPB_FWD_STEP(0) IPB_FWD_STEP(1) IPB_FWD_STEP(2) IPB_FWD_STEP(3)
                if (keep) yp += 4 * 32;
                __syncwarp();               // every lane has read the block: hand the buffer back
                if (lane == 0)
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(R.empties + buf * 8u) : "memory");
            }
        }
        // ================= backward: steps steps-1 .. ov
        {
            double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0;
            double yn[4];
            yp -= 4 * 32;                   // forward result of the 4 rows of the first block
#pragma unroll
            for (int q = 0; q < 4; ++q) yn[q] = yp[q * 32];
            for (int jb = steps - 4; jb >= ov; jb -= 4, ++u) {
                const double yc[4] = {yn[0], yn[1], yn[2], yn[3]};
                if (jb - 4 >= ov) {
                    yp -= 4 * 32;
#pragma unroll
                    for (int q = 0; q < 4; ++q) yn[q] = yp[q * 32];
                }
                double zo[4];
                const unsigned buf = u & R.mask;
                mbar_wait_s(R.fulls + buf * 8u, (u / (unsigned)depth) & 1u, S.abort_flag);
                const unsigned a = R.data + buf * (unsigned)(BAND_BLK * sizeof(double));
                IPB_BWD_STEP(3) IPB_BWD_STEP(2) IPB_BWD_STEP(1) IPB_BWD_STEP(0)
                __syncwarp();
                if (lane == 0)
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(R.empties + buf * 8u) : "memory");
                if (jb < ov + B) {
                    const int row = e0 + jb;             // >= 0: an owned row
                    if (row + 3 < m) {
                        double2 a2 = make_double2(zo[0], zo[1]), b2 = make_double2(zo[2], zo[3]);
                        if (accumulate) {
                            const double2 pa = *reinterpret_cast<const double2 *>(out + row);
                            const double2 pb = *reinterpret_cast<const double2 *>(out + row + 2);
                            a2.x += pa.x; a2.y += pa.y; b2.x += pb.x; b2.y += pb.y;
                        }
                        *reinterpret_cast<double2 *>(out + row) = a2;
                        *reinterpret_cast<double2 *>(out + row + 2) = b2;
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (row + q < m) out[row + q] = accumulate ? out[row + q] + zo[q] : zo[q];
                    }
                }
            }
        }
    }
    __syncwarp();
    if (lane == 0) sm.bseq[warp] += u - u0;
    __syncthreads();
    return ss;
}
#undef IPB_FWD_STEP
#undef IPB_BWD_STEP

// ---------------------------------------------------------------- (A A')^-1
// out = (A~ A~') in with  acc0 += out.dotv : one SpMV with the explicit
// product M when the projector holds it (half the matrix bytes and no
// n-vector round trip), otherwise A~ (A~' in) through P.t with a barrier in
// between.
__device__ __forceinline__ void normal_apply_to(const ProjView &P, const SyncView &S, Smem &sm,
                                                const double *in, double *out,
                                                const double *dotv, double &acc0) {
    double dummy = 0.0;
    Epi e = epi_plain(out);
    e.dotv = dotv;
    if (P.has_M) {
        spmv(P.M, in, e, sm, S, acc0, dummy);
    } else {
        spmv(P.At, in, epi_plain(P.t), sm, S, acc0, dummy);
        grid_sync(S, sm);
        acc0 = 0.0;
        spmv(P.A, P.t, e, sm, S, acc0, dummy);
    }
}

// w = K r with acc0 += w.r, K = A~ A~' -- or K = (A~ A~')^2 when ``squared``
// (the normal equations of a singular system, see inner_solve).
__device__ __forceinline__ void normal_apply(const ProjView &P, const SyncView &S,
                                             Smem &sm, double &acc0, bool squared = false) {
    if (!squared) {
        normal_apply_to(P, S, sm, P.r, P.w, P.r, acc0);
    } else {
        double dummy = 0.0;
        normal_apply_to(P, S, sm, P.r, P.q, nullptr, dummy);
        grid_sync(S, sm);
        acc0 = 0.0;
        normal_apply_to(P, S, sm, P.q, P.w, P.r, acc0);
    }
}

__device__ __forceinline__ void band_distrust(const ProjView &P, ProjStats &st) {
    st.band_ok = 0;
    st.n_fallbacks += 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) *P.band.ok = 0;
}

// Direct solve of M v = rhs (rhs in P.rhs, v in P.v) with the banded factor,
// verified by the residual and refined:  v = sweep(rhs);  r = rhs - M v;
// while ||r|| > rtol ||rhs||:  v += sweep(r).  Normally the first residual
// passes (the sweep is exact to round-off).  Returns false when the residual
// does not contract (slowly decaying M^-1, breakdown): the caller then runs
// the CG below from scratch and the factor is not used again.
__device__ __noinline__ bool inner_solve_band(const ProjView &P, const SyncView &S,
                                              Smem &sm, ProjStats &st) {
    double ss = band_sweep(P, S, sm, P.rhs, P.v, 0);
    st.n_sweeps += 1;
    grid_sync(S, sm);
    double rr0 = 0.0, rr_prev = INFINITY;
    int k = 0;
    while (true) {
        Epi e = epi_plain(P.r);
        e.sa = -1.0; e.addv = P.rhs; e.sb = 1.0; e.want_sq = 1;
        double a0 = 0.0, a1 = 0.0;
        spmv(P.M, P.v, e, sm, S, a0, a1);
        publish(S, sm, 0, a1, OP_SUM);
        publish(S, sm, 1, ss, OP_SUM);
        set_ops(sm, OP_SUM, OP_SUM);
        grid_sync(S, sm);
        collect(S, sm, 2);
        const double rr = sm.out[0];
        if (k == 0) rr0 = sm.out[1];
        ss = 0.0;
        ++k;
        if (rr <= P.rtol * P.rtol * rr0) break;                  // also rr0 == 0: v = 0
        if (!(rr == rr) || k >= 6 || (k >= 2 && rr > 0.0625 * rr_prev)) {
            band_distrust(P, st);
            st.inner_iters += k;
            return false;
        }
        rr_prev = rr;
        band_sweep(P, S, sm, P.r, P.v, 1);
        st.n_sweeps += 1;
        grid_sync(S, sm);
    }
    st.inner_iters += k;
    return true;
}

// Solve (A~ A~') v = rhs by CG, A~ the row-normalised A (Jacobi
// preconditioning folded into the matrix), rhs in P.rhs, v in P.v.
// Chronopoulos-Gear single-reduction CG:
//   r = rhs; w = M r; gamma = r.r; delta = r.w; alpha = gamma/delta; beta = 0
//   loop:  p = r + beta p;  s = w + beta s;  v += alpha p;  r -= alpha s
//          w = M r; gamma' = r.r; delta' = r.w           (one reduction)
//          beta = gamma'/gamma; alpha = gamma'/(delta' - beta*gamma'/alpha)
// Every CTA returns the same iteration count.  fixed_iters > 0 runs exactly
// that many iterations without a convergence test (bench only).  The target
// is ||r|| <= rtol ||rhs||, whatever the size of rhs: a right-hand side that
// is round-off noise (the second projection g = Z r of qp_subproblem.py:504)
// is solved like any other, as the reference's factorisation does.
__device__ __noinline__ int inner_solve_cg(const ProjView &P, const SyncView &S,
                                           Smem &sm, ProjStats &st, int fixed_iters = 0,
                                           bool squared = false) {
    const int gt = blockIdx.x * THREADS + threadIdx.x, nt = gridDim.x * THREADS;
    const int m = P.m;
    double a0;
    const double *rhs = P.rhs;
    if (squared) {
        // K v = M rhs with K = M^2: consistent whatever rhs is; from v = 0 CG
        // converges to the minimum-norm least-squares solution M^+ rhs
        double dummy = 0.0;
        normal_apply_to(P, S, sm, P.rhs, P.q, nullptr, dummy);
        grid_sync(S, sm);
        for (int i = gt; i < m; i += nt) P.w[i] = P.q[i];
        grid_sync(S, sm);
        rhs = P.w;                      // copied to r below before w is overwritten
    }
    // r = rhs, v = p = s = 0
    double gpart = 0.0;
    for (int i = gt; i < m; i += nt) {
        double ri = rhs[i];
        P.r[i] = ri; P.v[i] = 0.0; P.p[i] = 0.0; P.s[i] = 0.0;
        gpart += ri * ri;
    }
    grid_sync(S, sm);
    a0 = 0.0;
    normal_apply(P, S, sm, a0, squared);
    publish(S, sm, 0, a0, OP_SUM);
    publish(S, sm, 1, gpart, OP_SUM);
    set_ops(sm, OP_SUM, OP_SUM);
    grid_sync(S, sm);
    collect(S, sm, 2);
    double delta = sm.out[0], gamma = sm.out[1];
    const double gamma0 = gamma;
    if (!(gamma0 > 0.0)) return 0;           // zero right-hand side: v = 0
    const double target = P.rtol * P.rtol * gamma0;
    if (!(delta > 0.0)) { st.status = 5; return 0; }
    double alpha = gamma / delta, beta = 0.0;
    const int maxit = fixed_iters > 0 ? fixed_iters : P.maxit;
    int k = 0;
    bool converged = false;
    while (k < maxit) {
        ++k;
        gpart = cg_update(P, sm, S, alpha, beta);
        grid_sync(S, sm);
        a0 = 0.0;
        normal_apply(P, S, sm, a0, squared);
        publish(S, sm, 0, a0, OP_SUM);
        publish(S, sm, 1, gpart, OP_SUM);
        grid_sync(S, sm);
        collect(S, sm, 2);
        double delta2 = sm.out[0], gamma2 = sm.out[1];
        if (fixed_iters <= 0 && gamma2 <= target) { converged = true; break; }
        beta = gamma2 / gamma;
        double denom = delta2 - beta * gamma2 / alpha;
        if (!(denom > 0.0) || !(gamma2 == gamma2)) {
            if (fixed_iters <= 0) { st.status = 5; break; }
            denom = delta2;   // keep the bench loop finite
        }
        alpha = gamma2 / denom;
        gamma = gamma2;
    }
    st.inner_iters += k;
    // maxit reached without meeting the target (the loop leaves by `break` when
    // it converges): rank-deficient A with a right-hand side outside its range,
    // or a system too ill-conditioned for inner_maxit -- fail loudly, as the
    // reference does when its factorisation is singular (projections.py:101-108)
    if (fixed_iters <= 0 && !converged && st.status == 0) st.status = 5;
    return k;
}

// (A~ A~')^-1 rhs: banded direct solve, else CG; and when the CG breaks down
// or does not converge -- A A' singular: rank-deficient Jacobian -- the
// minimum-norm least-squares solution by CG on the squared operator, which is
// what the reference's fallback computes (dense SVD pseudo-inverse when its
// factorisation is singular, projections.py:101-108).  Fails loudly (status 5)
// only if that does not converge either.
__device__ __forceinline__ void inner_solve(const ProjView &P, const SyncView &S,
                                            Smem &sm, ProjStats &st, int fixed_iters = 0) {
    st.n_solves += 1;
    if (st.band_ok && fixed_iters <= 0 && inner_solve_band(P, S, sm, st)) return;
    inner_solve_cg(P, S, sm, st, fixed_iters);
    if (st.status == 5 && fixed_iters <= 0 && P.q != nullptr) {
        st.status = 0;
        st.n_singular += 1;
        grid_sync(S, sm);
        inner_solve_cg(P, S, sm, st, 0, true);
    }
}

// Z with the banded factor: the solve is checked by the product the
// reference's orthogonality test needs anyway -- w = A z is D^1/2 times the
// residual of the normal equations -- so M is not touched at all:
//   rhs = A~ x;  v = sweep(rhs);  z = x - A~' v;
//   w = A z;  rhs' = D^-1/2 w;  done when ||rhs'|| <= rtol max(||rhs||, ||z||)
//   and orthogonality <= orth_tol, else  v' = sweep(rhs');  z -= A~' v' ...
// -- the refinement loop of projections.py:72-78 with the solver's own accuracy
// as a second trigger.  Returns false when the residual does not contract.
__device__ __noinline__ bool null_space_band(const ProjView &P, const SyncView &S, Smem &sm,
                                             ProjStats &st, const double *x, double *z,
                                             double &zz_out) {
    double a0 = 0.0, a1 = 0.0;
    {
        Epi e = epi_plain(P.rhs);
        e.want_sq = 1;
        spmv(P.A, x, e, sm, S, a0, a1);
    }
    const double rr0_part = a1;              // published with the first collect below
    grid_sync(S, sm);
    band_sweep(P, S, sm, P.rhs, P.v, 0);
    st.n_sweeps += 1;
    st.n_solves += 1;
    grid_sync(S, sm);
    const double *src = x;
    double rr0 = 0.0, rr_prev = INFINITY, zz = 0.0;
    int k = 0;
    bool ok = true;
    while (true) {
        double zz_part;
        {
            Epi e = epi_plain(z);
            e.sa = -1.0; e.addv = src; e.sb = 1.0; e.want_sq = 1;
            a0 = a1 = 0.0;
            spmv(P.At, P.v, e, sm, S, a0, a1);
            zz_part = a1;
        }
        grid_sync(S, sm);
        {
            // w = A z with the un-normalised A, exactly what orthogonality()
            // measures (projections.py:52); D^-1/2 w is the next right-hand side
            Epi e = epi_plain(P.rhs);
            e.want_sq = 1; e.post = P.dscale; e.want_psq = 1;
            CsrView Ar = P.A; Ar.val = P.Araw;
            a0 = a1 = 0.0;
            spmv(Ar, z, e, sm, S, a0, a1);
        }
        publish(S, sm, 0, zz_part, OP_SUM);
        publish(S, sm, 1, a1, OP_SUM);
        publish(S, sm, 2, a0, OP_SUM);
        publish(S, sm, 3, rr0_part, OP_SUM);
        set_ops(sm, OP_SUM, OP_SUM, OP_SUM, OP_SUM);
        grid_sync(S, sm);
        collect(S, sm, 4);
        zz = sm.out[0];
        const double az = sm.out[1], rr = sm.out[2];
        if (k == 0) rr0 = sm.out[3];
        const double ng = sqrt(zz);
        double orth = 0.0;
        if (ng != 0.0 && P.fro != 0.0) orth = sqrt(az) / (P.fro * ng);
        st.orth = orth;
        ++k;
        const bool solved = rr <= P.rtol * P.rtol * fmax(rr0, zz);
        const bool orth_ok = !(orth > P.orth_tol);
        if (solved && orth_ok) break;
        const bool stalled = !(rr == rr) || k > 6 || (k >= 2 && rr > 0.0625 * rr_prev);
        if (stalled) {
            // the attainable accuracy is reached: accept by the reference's own
            // criterion, otherwise give the factor up
            if (!(orth_ok && rr == rr)) { band_distrust(P, st); ok = false; }
            break;
        }
        if (solved && k > P.max_refin) break;   // only the reference's trigger is left and its rounds are spent
        rr_prev = rr;
        band_sweep(P, S, sm, P.rhs, P.v, 0);
        st.n_sweeps += 1;
        st.n_refinements += 1;
        grid_sync(S, sm);
        src = z;
    }
    st.inner_iters += k;
    if (ok && st.orth > st.max_orth) st.max_orth = st.orth;
    zz_out = zz;
    return ok;
}

// z = x - A~'(A~A~')^-1 A~ x, then iterative refinement while
// orthogonality(A, z) > orth_tol (projections.py:65-80).  x and z may alias.
// Returns ||z||^2 (identical in every thread).
__device__ __noinline__ double null_space(const ProjView &P, const SyncView &S,
                                          Smem &sm, ProjStats &st,
                                          const double *x, double *z) {
    const int gt = blockIdx.x * THREADS + threadIdx.x, nt = gridDim.x * THREADS;
    double a0, a1;
    if (P.m == 0) {                                  // Z = I
        double zz = 0.0;
        for (int i = gt; i < P.n; i += nt) { double v = x[i]; z[i] = v; zz += v * v; }
        publish(S, sm, 0, zz, OP_SUM);
        set_ops(sm, OP_SUM);
        grid_sync(S, sm);
        collect(S, sm, 1);
        st.orth = 0.0;
        return sm.out[0];
    }
    if (st.band_ok) {
        double zz = 0.0;
        if (null_space_band(P, S, sm, st, x, z, zz)) return zz;
        // z = x - A~' v for some v: Z z = Z x, project z with the CG below
        grid_sync(S, sm);
        x = z;
    }
    a0 = a1 = 0.0;
    spmv(P.A, x, epi_plain(P.rhs), sm, S, a0, a1);
    grid_sync(S, sm);
    inner_solve(P, S, sm, st);
    const double *src = x;
    double zz = 0.0;
    int k = 0;
    while (true) {
        // z = src - A~' v, ||z||^2
        double zz_part;
        {
            Epi e = epi_plain(z);
            e.sa = -1.0; e.addv = src; e.sb = 1.0; e.want_sq = 1;
            a0 = a1 = 0.0;
            spmv(P.At, P.v, e, sm, S, a0, a1);
            zz_part = a1;
        }
        grid_sync(S, sm);
        // w = A z with the un-normalised A, exactly what orthogonality()
        // measures (projections.py:52); D^-1/2 w is the next right-hand side
        {
            Epi e = epi_plain(P.rhs);
            e.want_sq = 1; e.post = P.dscale;
            CsrView Ar = P.A; Ar.val = P.Araw;
            a0 = a1 = 0.0;
            spmv(Ar, z, e, sm, S, a0, a1);
        }
        publish(S, sm, 0, zz_part, OP_SUM);
        publish(S, sm, 1, a1, OP_SUM);
        set_ops(sm, OP_SUM, OP_SUM);
        grid_sync(S, sm);
        collect(S, sm, 2);
        zz = sm.out[0];
        double az = sm.out[1];
        double orth = 0.0;
        double ng = sqrt(zz);
        if (ng != 0.0 && P.fro != 0.0) orth = sqrt(az) / (P.fro * ng);
        st.orth = orth;
        if (!(orth > P.orth_tol) || k >= P.max_refin) break;
        // refine: solve M dv = A~ z, z -= A~' dv
        inner_solve(P, S, sm, st);
        st.n_refinements += 1;
        src = z;
        ++k;
    }
    if (st.orth > st.max_orth) st.max_orth = st.orth;
    return zz;
}

// v_out = (A A')^-1 A x = D^-1/2 (A~A~')^-1 A~ x
__device__ __noinline__ void least_squares(const ProjView &P, const SyncView &S,
                                           Smem &sm, ProjStats &st,
                                           const double *x, double *vout) {
    const int gt = blockIdx.x * THREADS + threadIdx.x, nt = gridDim.x * THREADS;
    if (P.m == 0) return;
    double a0 = 0.0, a1 = 0.0;
    spmv(P.A, x, epi_plain(P.rhs), sm, S, a0, a1);
    grid_sync(S, sm);
    inner_solve(P, S, sm, st);
    for (int i = gt; i < P.m; i += nt) vout[i] = P.dscale[i] * P.v[i];
}

// y = A'(A A')^-1 (sign*b) = A~' (A~A~')^-1 D^-1/2 (sign*b); returns the
// squared norm of y.
__device__ __noinline__ double row_space(const ProjView &P, const SyncView &S,
                                         Smem &sm, ProjStats &st,
                                         const double *b, double sign, double *y) {
    const int gt = blockIdx.x * THREADS + threadIdx.x, nt = gridDim.x * THREADS;
    if (P.m == 0) {
        for (int i = gt; i < P.n; i += nt) y[i] = 0.0;
        return 0.0;
    }
    for (int i = gt; i < P.m; i += nt) P.rhs[i] = P.dscale[i] * (sign * b[i]);
    grid_sync(S, sm);
    inner_solve(P, S, sm, st);
    Epi e = epi_plain(y);
    e.want_sq = 1;
    double a0 = 0.0, a1 = 0.0;
    spmv(P.At, P.v, e, sm, S, a0, a1);
    publish(S, sm, 0, a1, OP_SUM);
    set_ops(sm, OP_SUM);
    grid_sync(S, sm);
    collect(S, sm, 1);
    return sm.out[0];
}

// ---------------------------------------------------------------- intersections
struct Interval { double ta, tb; int hit; };

// sphere_intersections (qp_subproblem.py:66-149) from the three dot products.
__device__ __forceinline__ Interval ball_interval(double dd, double zd, double zz,
                                                  double radius, int line) {
    Interval I;
    if (dd == 0.0) { I.ta = 0; I.tb = 0; I.hit = 0; return I; }   // norm(d) == 0
    if (isinf(radius)) {
        I.hit = 1;
        if (line) { I.ta = -INFINITY; I.tb = INFINITY; } else { I.ta = 0; I.tb = 1; }
        return I;
    }
    double a = dd, b = 2 * zd, c = zz - radius * radius;
    double disc = b * b - 4 * a * c;
    if (disc < 0) { I.ta = 0; I.tb = 0; I.hit = 0; return I; }
    double sq = sqrt(disc);
    double aux = b + copysign(sq, b);
    double ta = -aux / (2 * a), tb = -2 * c / aux;
    if (tb < ta) { double t = ta; ta = tb; tb = t; }
    if (line) { I.ta = ta; I.tb = tb; I.hit = 1; return I; }
    if (tb < 0 || ta > 1) { I.ta = 0; I.tb = 0; I.hit = 0; return I; }
    I.ta = fmax(0.0, ta); I.tb = fmin(1.0, tb); I.hit = 1;
    return I;
}

// box_intersections (:152-234) from the reduced extrema.
__device__ __forceinline__ Interval box_interval(double dd, double tlo, double thi,
                                                 double bad_fixed, int line) {
    Interval I;
    if (dd == 0.0 || bad_fixed > 0.0) { I.ta = 0; I.tb = 0; I.hit = 0; return I; }
    I.hit = (tlo <= thi);
    if (!line) {
        if (thi < 0 || tlo > 1) { I.ta = 0; I.tb = 0; I.hit = 0; return I; }
        tlo = fmax(0.0, tlo); thi = fmin(1.0, thi);
    }
    I.ta = tlo; I.tb = thi;
    return I;
}

// box_sphere_intersections(z, scale*dvec, lb, ub, radius, line) (:237-303).
// One fused pass over z, dvec (and lb, ub when has_box); one barrier.
__device__ __noinline__ Interval box_ball(const SyncView &S, Smem &sm, int n,
                                          const double *z, const double *dvec,
                                          double scale, const double *lb,
                                          const double *ub, int has_box,
                                          double radius, int line) {
    const int gt = blockIdx.x * THREADS + threadIdx.x, nt = gridDim.x * THREADS;
    double dd = 0, zd = 0, zz = 0, tlo = -INFINITY, thi = INFINITY, bad = 0;
    for (int i = gt; i < n; i += nt) {
        double zi = z[i], di = scale * dvec[i];
        dd += di * di; zd += zi * di; zz += zi * zi;
        if (has_box) {
            double l = lb[i], u = ub[i];
            if (di == 0.0) {
                if (zi < l || zi > u) bad = 1.0;
            } else {
                double t1 = (l - zi) / di, t2 = (u - zi) / di;
                tlo = fmax(tlo, fmin(t1, t2));
                thi = fmin(thi, fmax(t1, t2));
            }
        }
    }
    publish(S, sm, 0, dd, OP_SUM);
    publish(S, sm, 1, zd, OP_SUM);
    publish(S, sm, 2, zz, OP_SUM);
    publish(S, sm, 3, tlo, OP_MAX);
    publish(S, sm, 4, thi, OP_MIN);
    publish(S, sm, 5, bad, OP_MAX);
    set_ops(sm, OP_SUM, OP_SUM, OP_SUM, OP_MAX, OP_MIN, OP_MAX);
    grid_sync(S, sm);
    collect(S, sm, 6);
    dd = sm.out[0]; zd = sm.out[1]; zz = sm.out[2];
    // norm(d) == 0 test of the reference is on sqrt(dd)
    Interval Ib;
    if (has_box) Ib = box_interval(dd, sm.out[3], sm.out[4], sm.out[5], line);
    else {
        // unbounded box: every t is feasible unless d == 0
        if (dd == 0.0) { Ib.ta = 0; Ib.tb = 0; Ib.hit = 0; }
        else if (line) { Ib.ta = -INFINITY; Ib.tb = INFINITY; Ib.hit = 1; }
        else { Ib.ta = 0; Ib.tb = 1; Ib.hit = 1; }
    }
    Interval Is = ball_interval(dd, zd, zz, radius, line);
    Interval I;
    I.ta = fmax(Ib.ta, Is.ta);
    I.tb = fmin(Ib.tb, Is.tb);
    I.hit = (Ib.hit && Is.hit && I.ta <= I.tb);
    __syncthreads();
    return I;
}

__device__ __forceinline__ double clipd(double x, double l, double u) {
    return fmin(fmax(x, l), u);
}

// ---------------------------------------------------------------- projected CG
struct PcgArgs {
    HessView H;
    ProjView P;
    SyncView S;
    int n, m;
    const double *c, *b, *lb, *ub;     // lb/ub null when has_box == 0
    int has_box;
    double radius, tol;                // tol < 0 or NaN: default rule
    long long max_iter, max_infeasible;
    long long nm_global;               // n - m summed over the ranks
    // work vectors (n): x, xn, r, p, Hp, xfeas, tmp
    double *x, *xn, *r, *p, *Hp, *xfeas, *tmp;
    double *allvecs; long long allvecs_cap;
    // result
    struct Result {
        long long niter;
        int stop_cond, hits_boundary, status, n_allvecs;
        long long inner_iters, n_solves, n_refinements, n_spmv_h;
        double max_orth, tol_used;
        int x_slot;                    // 0: result in x, 1: xn, 2: xfeas
        int n_fallbacks;               // banded direct solves given up for CG
        long long n_sweeps;            // banded sweeps (direct solves incl. refinement)
        int fused_hp;                  // p = -g + beta p was formed inside H p
    } *result;
};

__device__ __forceinline__ void store_allvec(const PcgArgs &a, int &count,
                                             const double *v) {
    if (!a.allvecs) return;
    if (count < a.allvecs_cap) {
        const int gt = blockIdx.x * THREADS + threadIdx.x, nt = gridDim.x * THREADS;
        double *dst = a.allvecs + (size_t)count * a.n;
        for (int i = gt; i < a.n; i += nt) dst[i] = v[i];
    }
    ++count;
}

__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
k_projected_cg(const PcgArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    pipeline_init(sm, a.S);
    const SyncView &S = a.S;
    const int n = a.n;
    const int gt = blockIdx.x * THREADS + threadIdx.x, nt = gridDim.x * THREADS;
    ProjStats st = proj_stats_init(a.P);
    double *x = a.x, *xn = a.xn, *r = a.r, *p = a.p, *Hp = a.Hp;
    double *tmpv = a.tmp;              // r_next; changes places with p when H p forms p on the fly
    int fused_hp = 0;
    long long n_spmv_h = 0;
    int nvec = 0;
    int status = 0, stop_cond = 1, hits = 0;
    long long k = 0;
    double tol = a.tol;
    double acc0;
    int x_slot = 0;

    // ---- x = Y(-b)  (:502); exactly 0 when b == 0, as in the reference
    double bb = 0.0;
    for (int i = gt; i < a.m; i += nt) { double v = a.b[i]; bb += v * v; }
    publish(S, sm, 0, bb, OP_SUM);
    set_ops(sm, OP_SUM);
    grid_sync(S, sm);
    collect(S, sm, 1);
    bb = sm.out[0];
    double xx = 0.0;
    if (bb > 0.0) {
        xx = row_space(a.P, S, sm, st, a.b, -1.0, x);
    } else {
        for (int i = gt; i < n; i += nt) x[i] = 0.0;
        grid_sync(S, sm);
    }
    // ---- r = Z(H x + c), g = Z r, p = -g  (:503-505)
    if (bb > 0.0) {
        acc0 = 0.0;
        hess_apply(a.H, S, x, a.tmp, nullptr, sm, acc0);
        ++n_spmv_h;
        grid_sync(S, sm);
        for (int i = gt; i < n; i += nt) a.tmp[i] = a.tmp[i] + a.c[i];
    } else {
        for (int i = gt; i < n; i += nt) a.tmp[i] = 0.0 + a.c[i];
    }
    grid_sync(S, sm);
    null_space(a.P, S, sm, st, a.tmp, r);
    grid_sync(S, sm);
    double gg = null_space(a.P, S, sm, st, r, r);      // r now holds g
    for (int i = gt; i < n; i += nt) p[i] = -r[i];
    grid_sync(S, sm);
    store_allvec(a, nvec, x);
    // ---- H_p = H p, rt_g = ||g||^2  (:511-512)
    acc0 = 0.0;
    hess_apply(a.H, S, p, Hp, p, sm, acc0);
    ++n_spmv_h;
    publish(S, sm, 0, acc0, OP_SUM);
    set_ops(sm, OP_SUM);
    grid_sync(S, sm);
    collect(S, sm, 1);
    double pHp = sm.out[0];
    double ng = sqrt(gg);
    double rt_g = ng * ng;

    // ---- trust-region feasibility of x  (:515-526)
    double tr_distance = a.radius - sqrt(xx);
    bool done = false, early = false;
    // x = Y(-b) comes out of a different factorisation than the reference's, so
    // its norm can differ in the last bits: a violation within a few ulps of the
    // radius is "barely feasible" (reference test_qp_subproblem.py:473-489 sits
    // exactly there), not an error
    if (tr_distance < -8.0 * 2.220446049250313e-16 * a.radius) { status = 3; done = true; }
    else if (tr_distance < 1e-25) {
        stop_cond = 2; hits = 1; done = true; early = true;   // returned as it is (:521-526)
        store_allvec(a, nvec, x);
    }
    if (!(tol >= 0.0))                                   // default (:529-530)
        tol = fmax(fmin(0.01 * sqrt(rt_g), 0.1 * rt_g), 1e-25);
    long long nm = a.nm_global;                          // n - m of the whole problem
    long long max_iter = a.max_iter < 0 ? nm : a.max_iter;
    if (max_iter > nm) max_iter = nm;
    long long max_inf = a.max_infeasible < 0 ? nm : a.max_infeasible;
    long long counter = 0;

    if (!done) {
        // last_feasible_x: zeros (the reference leaves it uninitialised, :547)
        if (a.has_box) for (int i = gt; i < n; i += nt) a.xfeas[i] = 0.0;
        for (long long it = 0; it < max_iter; ++it) {
            if (rt_g < tol) { stop_cond = 4; break; }              // :551-553
            ++k;
            if (pHp <= 0) {                                          // :558-576
                if (isinf(a.radius)) { status = 4; break; }
                Interval I = box_ball(S, sm, n, x, p, 1.0, a.lb, a.ub, a.has_box,
                                      a.radius, 1);
                double alpha = I.tb;
                for (int i = gt; i < n; i += nt) {
                    double v = x[i];
                    if (I.hit) v = v + alpha * p[i];
                    if (a.has_box) v = clipd(v, a.lb[i], a.ub[i]);
                    x[i] = v;
                }
                stop_cond = 3; hits = 1;
                break;
            }
            double alpha = rt_g / pHp;                               // :579-580
            // x_next = x + alpha p; ||x_next||^2; inside-box flag;
            // r_next = r + alpha H_p (:622) in the same pass.
            double sq = 0.0, out = 0.0;
            if (!a.has_box) {
                constexpr int TS = (int)((sizeof(Stage) / (4 * sizeof(double))) & ~(size_t)31);
                TileIO<4> io = {{x, p, r, Hp}, {xn, nullptr, tmpv, nullptr}};
                tile_pass<4>(sm, S, n, io, [&](double *t, int e) {
                    const double v = t[e] + alpha * t[TS + e];
                    t[e] = v;
                    sq += v * v;
                    t[2 * TS + e] = t[2 * TS + e] + alpha * t[3 * TS + e];
                });
            } else {
                constexpr int TS = (int)((sizeof(Stage) / (6 * sizeof(double))) & ~(size_t)31);
                TileIO<6> io = {{x, p, r, Hp, a.lb, a.ub},
                                {xn, nullptr, tmpv, nullptr, nullptr, nullptr}};
                tile_pass<6>(sm, S, n, io, [&](double *t, int e) {
                    const double v = t[e] + alpha * t[TS + e];
                    t[e] = v;
                    sq += v * v;
                    if (v < t[4 * TS + e] || v > t[5 * TS + e]) out = 1.0;
                    t[2 * TS + e] = t[2 * TS + e] + alpha * t[3 * TS + e];
                });
            }
            // ||x_next|| and the box flag decide nothing when the region is the whole
            // space: no reduction (and, row-partitioned, no cross-GPU exchange) then
            const bool unbounded = !a.has_box && isinf(a.radius);
            if (!unbounded) {
                publish(S, sm, 0, sq, OP_SUM);
                publish(S, sm, 1, out, OP_MAX);
                set_ops(sm, OP_SUM, OP_MAX);
            }
            grid_sync(S, sm);
            if (!unbounded) collect(S, sm, 2);
            double nxn = unbounded ? 0.0 : sqrt(sm.out[0]);
            bool inside = unbounded || !(sm.out[1] > 0.0);
            if (nxn >= a.radius) {                                   // :583-596
                Interval I = box_ball(S, sm, n, x, p, alpha, a.lb, a.ub, a.has_box,
                                      a.radius, 0);
                double ta = I.tb * alpha;                            // theta*alpha
                for (int i = gt; i < n; i += nt) {
                    double v = x[i];
                    if (I.hit) v = v + ta * p[i];
                    if (a.has_box) v = clipd(v, a.lb[i], a.ub[i]);
                    x[i] = v;
                }
                stop_cond = 2; hits = 1;
                break;
            }
            if (inside) counter = 0; else ++counter;                 // :599-613
            if (counter > 0) {
                Interval I = box_ball(S, sm, n, x, p, alpha, a.lb, a.ub, a.has_box,
                                      a.radius, 0);
                if (I.hit) {
                    double ta = I.tb * alpha;
                    for (int i = gt; i < n; i += nt)
                        a.xfeas[i] = clipd(x[i] + ta * p[i], a.lb[i], a.ub[i]);
                    counter = 0;
                }
            }
            if (counter > max_inf) break;                            // :615-616
            store_allvec(a, nvec, xn);
            // g_next = Z r_next (:624);  r <- g_next (:632)
            double gg2 = null_space(a.P, S, sm, st, tmpv, r);
            double ng2 = sqrt(gg2);
            double rt_g_next = ng2 * ng2;                            // :626
            double beta = rt_g_next / rt_g;                          // :627
            { double *t = x; x = xn; xn = t; x_slot ^= 1; }          // :630
            rt_g = rt_g_next;                                        // :633
            acc0 = 0.0;
            if (hess_can_xform(a.H, S)) {
                // p = -g_next + beta p (:628) formed inside H p (:634): the new p goes to
                // the (now dead) r_next buffer, which then changes places with p
                hess_apply_xform(a.H, S, r, p, beta, tmpv, Hp, sm, acc0);
                { double *t = p; p = tmpv; tmpv = t; }
                fused_hp = 1;
            } else {
                {                                                    // p = -g_next + beta p  (:628)
                    constexpr int TS = (int)((sizeof(Stage) / (2 * sizeof(double))) & ~(size_t)31);
                    TileIO<2> io = {{r, p}, {nullptr, p}};
                    tile_pass<2>(sm, S, n, io, [&](double *t, int e) {
                        t[TS + e] = -t[e] + beta * t[TS + e];
                    });
                }
                grid_sync(S, sm);
                hess_apply(a.H, S, p, Hp, p, sm, acc0);              // :634
            }
            ++n_spmv_h;
            publish(S, sm, 0, acc0, OP_SUM);
            set_ops(sm, OP_SUM);
            grid_sync(S, sm);
            collect(S, sm, 1);
            pHp = sm.out[0];
        }
    }
    // ---- final feasibility fix-up (:636-638)
    int final_slot = x_slot;
    if (status == 0 && a.has_box && !early) {
        grid_sync(S, sm);
        double out = 0.0;
        for (int i = gt; i < n; i += nt) {
            double v = x[i];
            if (v < a.lb[i] || v > a.ub[i]) out = 1.0;
        }
        publish(S, sm, 0, out, OP_MAX);
        set_ops(sm, OP_MAX);
        grid_sync(S, sm);
        collect(S, sm, 1);
        if (sm.out[0] > 0.0) { final_slot = 2; hits = 1; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        PcgArgs::Result &R = *a.result;
        R.niter = k; R.stop_cond = stop_cond; R.hits_boundary = hits;
        R.status = status ? status : st.status;
        R.n_allvecs = nvec;
        R.inner_iters = st.inner_iters; R.n_solves = st.n_solves;
        R.n_refinements = st.n_refinements; R.n_spmv_h = n_spmv_h;
        R.max_orth = st.max_orth; R.tol_used = tol;
        R.x_slot = final_slot;
        R.n_fallbacks = (int)st.n_fallbacks; R.n_sweeps = st.n_sweeps;
        R.fused_hp = fused_hp;
    }
}

// ---------------------------------------------------------------- projector kernels
struct ProjArgs {
    ProjView P;
    SyncView S;
    int op;                      // 0: Z, 1: LS, 2: Y, 3: orthogonality, 4: bench
    const double *in;
    double *out;
    int fixed_iters;
    struct Result {
        long long inner_iters, n_solves, n_refinements;
        int status;
        int n_fallbacks;
        double orth;
        long long n_sweeps;
    } *result;
};

__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
k_projector(const ProjArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    pipeline_init(sm, a.S);
    ProjStats st = proj_stats_init(a.P);
    const SyncView &S = a.S;
    const ProjView &P = a.P;
    if (a.op == 0) {
        null_space(P, S, sm, st, a.in, a.out);
    } else if (a.op == 1) {
        least_squares(P, S, sm, st, a.in, a.out);
    } else if (a.op == 2) {
        row_space(P, S, sm, st, a.in, 1.0, a.out);
    } else if (a.op == 3) {
        // orthogonality(A, g): ||A g|| / (||A||_F ||g||)
        const int gt = blockIdx.x * THREADS + threadIdx.x, nt = gridDim.x * THREADS;
        double gg = 0.0, a0 = 0.0, a1 = 0.0;
        for (int i = gt; i < P.n; i += nt) { double v = a.in[i]; gg += v * v; }
        if (P.m > 0) {
            Epi e = epi_plain(P.rhs);
            e.want_sq = 1;
            CsrView Ar = P.A; Ar.val = P.Araw;
            spmv(Ar, a.in, e, sm, S, a0, a1);
        }
        publish(S, sm, 0, gg, OP_SUM);
        publish(S, sm, 1, a1, OP_SUM);
        set_ops(sm, OP_SUM, OP_SUM);
        grid_sync(S, sm);
        collect(S, sm, 2);
        double ng = sqrt(sm.out[0]);
        st.orth = (ng == 0.0 || P.fro == 0.0) ? 0.0 : sqrt(sm.out[1]) / (P.fro * ng);
    } else if (a.op == 4) {
        double a0 = 0.0, a1 = 0.0;
        spmv(P.A, a.in, epi_plain(P.rhs), sm, S, a0, a1);
        grid_sync(S, sm);
        inner_solve(P, S, sm, st, a.fixed_iters);
    } else if (a.op >= 10) {
        // one phase of the inner iteration repeated fixed_iters times, each
        // followed by a grid barrier (measurement only; results discarded)
        double a0 = 0.0, a1 = 0.0;
        for (int it = 0; it < a.fixed_iters; ++it) {
            if (a.op == 10) {
                a1 += cg_update(P, sm, S, 1e-3, 0.5);
            } else if (a.op == 11) {
                spmv(P.At, P.r, epi_plain(P.t), sm, S, a0, a1);
            } else if (a.op == 12) {
                Epi e = epi_plain(P.w); e.dotv = P.r;
                spmv(P.A, P.t, e, sm, S, a0, a1);
            } else if (a.op == 13 && P.has_M) {
                Epi e = epi_plain(P.w); e.dotv = P.r;
                spmv(P.M, P.r, e, sm, S, a0, a1);
            } else if (a.op == 15 && P.has_M) {
                spmv(P.M, P.r, epi_plain(nullptr), sm, S, a0, a1);   // bulk copies only
            } else if (a.op == 18 && P.has_M) {
                Epi e = epi_plain(P.w); e.dotv = P.r;
                CsrView Mp = P.M; Mp.space = 2;                     // row arithmetic only
                spmv(Mp, P.r, e, sm, S, a0, a1);
            } else if (a.op == 16) {
                // a two-value reduction exactly as the solvers do it
                publish(S, sm, 0, 1.0, OP_SUM);
                publish(S, sm, 1, 2.0, OP_SUM);
                set_ops(sm, OP_SUM, OP_SUM);
                grid_sync(S, sm);
                collect(S, sm, 2);
                a0 += sm.out[0] + sm.out[1];
            } else if (a.op == 17) {
                if (S.C.size > 1) halo_sync(S, sm, P.r, 1);         // + ghost push of an m-vector
            } else if (a.op == 19 && P.band.enabled) {
                a1 += band_sweep(P, S, sm, P.rhs, P.w, 0);          // banded direct solve, one sweep
            }
            grid_sync(S, sm);
        }
        if (a0 + a1 == 1.2345e-300) st.orth = a0;      // keep the sums alive
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.result->inner_iters = st.inner_iters;
        a.result->n_solves = st.n_solves;
        a.result->n_refinements = st.n_refinements;
        a.result->status = st.status;
        a.result->orth = st.orth;
        a.result->n_fallbacks = (int)st.n_fallbacks;
        a.result->n_sweeps = st.n_sweeps;
    }
}

// ---------------------------------------------------------------- modified dogleg
struct DoglegArgs {
    ProjView P;
    SyncView S;
    CsrView Araw, Atraw;         // un-normalised A and A'
    int n, m;
    const double *b, *lb, *ub;
    int has_box;
    double radius;
    double *newton, *cauchy, *g, *x1, *x2, *out;   // n
    double *mtmp;                                    // m
    ProjArgs::Result *result;
};

// modified_dogleg (qp_subproblem.py:320-413), whole routine in one launch.
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
k_dogleg(const DoglegArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    pipeline_init(sm, a.S);
    ProjStats st = proj_stats_init(a.P);
    const SyncView &S = a.S;
    const int n = a.n, m = a.m;
    const int gt = blockIdx.x * THREADS + threadIdx.x, nt = gridDim.x * THREADS;
    double a0, a1;
    // newton = -Y b  (:368)
    double nn = row_space(a.P, S, sm, st, a.b, -1.0, a.newton);
    // inside box and ball? (:370-373)
    double out = 0.0;
    if (a.has_box)
        for (int i = gt; i < n; i += nt) {
            double v = a.newton[i];
            if (v < a.lb[i] || v > a.ub[i]) out = 1.0;
        }
    publish(S, sm, 0, out, OP_MAX);
    set_ops(sm, OP_MAX);
    grid_sync(S, sm);
    collect(S, sm, 1);
    const double *res = a.newton;
    if (!(sm.out[0] > 0.0) && sqrt(nn) <= a.radius) {
        res = a.newton;
    } else {
        // g = A' b (:376); A g (:379); cauchy = -(g.g)/(Ag.Ag) g (:380)
        a0 = a1 = 0.0;
        Epi eg = epi_plain(a.g); eg.want_sq = 1;
        spmv(a.Atraw, a.b, eg, sm, S, a0, a1);
        const double gg_part = a1;
        grid_sync(S, sm);
        Epi eag = epi_plain(a.mtmp); eag.want_sq = 1;
        a0 = a1 = 0.0;
        spmv(a.Araw, a.g, eag, sm, S, a0, a1);
        publish(S, sm, 0, gg_part, OP_SUM);
        publish(S, sm, 1, a1, OP_SUM);
        set_ops(sm, OP_SUM, OP_SUM);
        grid_sync(S, sm);
        collect(S, sm, 2);
        double coef = -sm.out[0] / sm.out[1];
        // cauchy, and step = newton - cauchy kept in x2 for now
        for (int i = gt; i < n; i += nt) {
            double cp = coef * a.g[i];
            a.cauchy[i] = cp;
            a.x2[i] = a.newton[i] - cp;
        }
        grid_sync(S, sm);
        // segment cauchy -> newton (:386-391)
        Interval I = box_ball(S, sm, n, a.cauchy, a.x2, 1.0, a.lb, a.ub, a.has_box,
                              a.radius, 0);
        if (I.hit) {
            for (int i = gt; i < n; i += nt) a.x1[i] = a.cauchy[i] + I.tb * a.x2[i];
        } else {
            // segment origin -> cauchy (:395-399): z = 0
            for (int i = gt; i < n; i += nt) a.g[i] = 0.0;     // g reused as origin
            grid_sync(S, sm);
            Interval J = box_ball(S, sm, n, a.g, a.cauchy, 1.0, a.lb, a.ub, a.has_box,
                                  a.radius, 0);
            for (int i = gt; i < n; i += nt) a.x1[i] = 0.0 + J.tb * a.cauchy[i];
        }
        // segment origin -> newton (:403-407)
        for (int i = gt; i < n; i += nt) a.g[i] = 0.0;
        grid_sync(S, sm);
        Interval K = box_ball(S, sm, n, a.g, a.newton, 1.0, a.lb, a.ub, a.has_box,
                              a.radius, 0);
        for (int i = gt; i < n; i += nt) a.x2[i] = 0.0 + K.tb * a.newton[i];
        grid_sync(S, sm);
        // ||A x1 + b|| < ||A x2 + b|| ? (:410-413)
        Epi e1 = epi_plain(a.mtmp); e1.addv = a.b; e1.sb = 1.0; e1.want_sq = 1;
        a0 = a1 = 0.0;
        spmv(a.Araw, a.x1, e1, sm, S, a0, a1);
        const double r1_part = a1;
        grid_sync(S, sm);
        a0 = a1 = 0.0;
        spmv(a.Araw, a.x2, e1, sm, S, a0, a1);
        publish(S, sm, 0, r1_part, OP_SUM);
        publish(S, sm, 1, a1, OP_SUM);
        set_ops(sm, OP_SUM, OP_SUM);
        grid_sync(S, sm);
        collect(S, sm, 2);
        res = (sqrt(sm.out[0]) < sqrt(sm.out[1])) ? a.x1 : a.x2;
    }
    for (int i = gt; i < n; i += nt) a.out[i] = res[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.result->inner_iters = st.inner_iters;
        a.result->n_solves = st.n_solves;
        a.result->n_refinements = st.n_refinements;
        a.result->status = st.status;
        a.result->orth = 0.0;
        a.result->n_fallbacks = (int)st.n_fallbacks;
        a.result->n_sweeps = st.n_sweeps;
    }
    (void)m;
}

// ---------------------------------------------------------------- intersections
struct IntersectArgs {
    SyncView S;
    int n;
    const double *z, *d, *lb, *ub;
    int has_box, line;
    double radius;
    struct Result { double ta, tb; int hit; } *result;
};

// box_sphere_intersections / box_intersections / sphere_intersections
// (qp_subproblem.py:66-303) for host callers (the SQP second-order
// correction, equality_constrained_sqp.py:177).
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
k_intersect(const IntersectArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    pipeline_init(sm, a.S);
    Interval I = box_ball(a.S, sm, a.n, a.z, a.d, 1.0, a.lb, a.ub, a.has_box,
                          a.radius, a.line);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.result->ta = I.ta; a.result->tb = I.tb; a.result->hit = I.hit;
    }
}

// ---------------------------------------------------------------- plain products
struct MatvecArgs {
    int kind;                    // 0: CSR, 1: Hessian
    CsrView M;
    HessView H;
    SyncView S;
    const double *x;
    double *y;
};

__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
k_matvec(const MatvecArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    pipeline_init(sm, a.S);
    double a0 = 0.0, a1 = 0.0;
    if (a.kind == 0) spmv(a.M, a.x, epi_plain(a.y), sm, a.S, a0, a1);
    else hess_apply(a.H, a.S, a.x, a.y, nullptr, sm, a0);
}


// ---------------------------------------------------------------- SQP step glue
// Elementwise stages between the solve kernels of one SQP step (ipb_sqp_step);
// ordinary launches.  Each stage may produce up to three sums: per-CTA
// partials, then the last CTA to finish adds them up in CTA order -- a fixed
// order, so the scalars are run-to-run reproducible.
struct GlueArgs {
    int mode, len;
    const double *a, *b, *c;     // inputs (meaning per mode)
    double *o0, *o1;             // outputs
    double s0;                   // scalar parameter
    double *partials;            // [3][gridDim.x]
    unsigned *counter;
    double *sums;                // [3] results
};

__global__ void __launch_bounds__(256) k_glue(const GlueArgs g) {
    __shared__ double red[3][8];
    __shared__ bool last;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.len; i += gridDim.x * blockDim.x) {
        if (g.mode == 0) {                 // o0 = a + b (H dn + c); sums: ||c||^2 of input c
            const double dn = g.c[i];
            g.o0[i] = g.a[i] + g.b[i];
            v0 += dn * dn;
        } else if (g.mode == 1) {          // o0 = a - c, o1 = b - c  (lb - dn, ub - dn)
            const double dn = g.c[i];
            g.o0[i] = g.a[i] - dn;
            g.o1[i] = g.b[i] - dn;
        } else if (g.mode == 2) {          // o0 = d = a + b (dn + dt); sums: ||dt||^2, ||d||^2, c.d
            const double dt = g.b[i], d = g.a[i] + dt;
            g.o0[i] = d;
            v0 += dt * dt; v1 += d * d; v2 += g.c[i] * d;
        } else if (g.mode == 3) {          // sums: a.b (Hd.d)
            v0 += g.a[i] * g.b[i];
        } else if (g.mode == 4) {          // o0 = a + b (A d + b); sums: ||o0||^2, ||b||^2
            const double bb = g.b[i], l = g.a[i] + bb;
            g.o0[i] = l;
            v0 += l * l; v1 += bb * bb;
        } else if (g.mode == 5) {          // o0 = s0 * a, o1 = s0 * b  (scaled bounds)
            g.o0[i] = g.s0 * g.a[i];
            g.o1[i] = g.s0 * g.b[i];
        } else if (g.mode == 6) {          // o0 = -a (v = -LS c)
            g.o0[i] = -g.a[i];
        } else if (g.mode == 7) {          // max |a + b|  (||c + A'v||_inf) in sums[0]
            v0 = fmax(v0, fabs(g.a[i] + g.b[i]));
        }
    }
    const bool is_max = g.mode == 7;
    double v[3] = {v0, v1, v2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double t = __shfl_xor_sync(0xffffffffu, v[k], o);
            v[k] = is_max ? fmax(v[k], t) : v[k] + t;
        }
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; ++k) {
            double a = red[k][0];
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) a = is_max ? fmax(a, red[k][w]) : a + red[k][w];
            g.partials[k * gridDim.x + blockIdx.x] = a;
        }
        __threadfence();
        last = atomicAdd(g.counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x < 3) {
        __threadfence();
        const int k = threadIdx.x;
        double a = is_max ? 0.0 : 0.0;
        for (unsigned q = 0; q < gridDim.x; ++q) {
            const double t = __ldcg(&g.partials[k * gridDim.x + q]);
            a = is_max ? fmax(a, t) : a + t;
        }
        g.sums[k] = a;
        if (k == 0) *g.counter = 0u;
    }
}

// ---------------------------------------------------------------- setup kernels
// (ordinary launches, not on the per-iteration path)
__global__ void k_row_norms(int rows, const int *rowptr, const double *val,
                            double *dnorm2, double *dscale) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    double s = 0.0;
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) s += val[k] * val[k];
    dnorm2[r] = s;
    dscale[r] = s > 0.0 ? 1.0 / sqrt(s) : 0.0;
}

__global__ void k_scale_rows(int rows, const int *rowptr, const double *val,
                             const double *dscale, double *out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    double d = dscale[r];
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) out[k] = val[k] * d;
}

__global__ void k_gather(long long nnz, const int *perm, const double *src,
                         double *dst) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < nnz) dst[i] = src[perm[i]];
}

__global__ void k_gather_int(long long n, const int *perm, const int *src,
                             int *dst) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[perm[i]];
}

__global__ void k_expand_rows(int rows, const int *rowptr, int *rowidx) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) rowidx[k] = r;
}

__global__ void k_count_cols(long long nnz, const int *col, int *counts) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < nnz) atomicAdd(&counts[col[i] + 1], 1);
}

__global__ void k_iota(long long n, int *out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)i;
}

__global__ void k_narrow_i64(long long n, const long long *in, int *out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)in[i];
}

// ---- explicit M = A A' (pattern once per sparsity pattern, values per update)
// Candidate count of row i: sum over its columns j of the column's length.
__global__ void k_aat_count(int rows, const int *rowptr, const int *col,
                            const int *t_rowptr, long long *cnt) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    long long c = 0;
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
        int j = col[k];
        c += t_rowptr[j + 1] - t_rowptr[j];
    }
    cnt[r] = c;
}

// keys[cand_ptr[i] + ...] = (i << 32) | k for every row k that shares a column with row i
__global__ void k_aat_expand(int rows, const int *rowptr, const int *col,
                             const int *t_rowptr, const int *t_col,
                             const long long *cand_ptr, unsigned long long *keys) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    long long o = cand_ptr[r];
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
        int j = col[k];
        for (int q = t_rowptr[j]; q < t_rowptr[j + 1]; ++q)
            keys[o++] = ((unsigned long long)(unsigned)r << 32) | (unsigned)t_col[q];
    }
}

// CSR arrays of M from its sorted unique (row, col) keys
__global__ void k_aat_rows(long long nnz, int rows, const unsigned long long *keys,
                           int *rowptr, int *col) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i > nnz) return;
    int row = i < nnz ? (int)(keys[i] >> 32) : rows;
    int prev = i > 0 ? (int)(keys[i - 1] >> 32) : -1;
    for (int r = prev + 1; r <= row; ++r) rowptr[r] = (int)i;
    if (i < nnz) col[i] = (int)(keys[i] & 0xffffffffull);
}

// M[i,k] = sum_j a_ij a_kj, j ascending (rows of A are sorted): both M[i,k]
// and M[k,i] come out bit-identical.
__global__ void k_aat_values(int rows, const int *rowptr, const int *col,
                             const double *val, const int *m_rowptr,
                             const int *m_col, double *m_val) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int a0 = rowptr[i], a1 = rowptr[i + 1];
    for (int e = m_rowptr[i]; e < m_rowptr[i + 1]; ++e) {
        const int k = m_col[e];
        int p = a0, q = rowptr[k];
        const int q1 = rowptr[k + 1];
        double s = 0.0;
        while (p < a1 && q < q1) {
            int cp = col[p], cq = col[q];
            if (cp == cq) { s += val[p] * val[q]; ++p; ++q; }
            else if (cp < cq) ++p;
            else ++q;
        }
        m_val[e] = s;
    }
}


// ---- banded Cholesky factor of M (see BandView) ---------------------------
// largest |row - col| over the entries of a CSR matrix
__global__ void k_band_width(int rows, const int *rowptr, const int *col, int *out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    int w = 0;
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) w = max(w, abs(col[k] - r));
    if (w > 0) atomicMax(out, w);
}

// Rows [c*B, (c+1)*B) of the Cholesky factor of the banded SPD matrix M, one
// thread per chunk c.  The recurrence is started ``ovf`` rows early from a
// truncated matrix: it forgets its start as fast as M^-1 decays, so the rows
// written agree with the global factor to round-off (measured: ovf = 32 gives
// 1e-16).  Lrow[i][0] = 1 / L_ii, Lrow[i][k] = L[i][i-k].  A non-positive
// pivot clears *ok (the caller then keeps to CG).
__global__ void k_band_factor(int m, const int *rowptr, const int *col, const double *val,
                              int B, int ovf, double *Lrow, int *ok) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long i0 = c * B;
    if (i0 >= m) return;
    const long long i1 = min((long long)m, i0 + B);
    const long long a = max(0ll, i0 - ovf);
    double ring[HB][HB];
#pragma unroll
    for (int q = 0; q < HB; ++q)
#pragma unroll
        for (int k = 0; k < HB; ++k) ring[q][k] = 0.0;
    for (long long i = a; i < i1; ++i) {
        double row[HB];
#pragma unroll
        for (int k = 0; k < HB; ++k) row[k] = 0.0;
        for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            const long long k = i - col[e];
            if (k >= 0 && k < HB && col[e] >= a) row[k] = val[e];
        }
        for (int k = HB - 1; k >= 1; --k) {
            const long long j = i - k;
            if (j < a) { row[k] = 0.0; continue; }
            double s = row[k];
            for (int t = k + 1; t < HB; ++t)
                if (i - t >= a) s = __fma_rn(-row[t], ring[j & (HB - 1)][t - k], s);
            row[k] = s * ring[j & (HB - 1)][0];
        }
        double d = row[0];
        for (int t = 1; t < HB; ++t) d = __fma_rn(-row[t], row[t], d);
        if (!(d > 0.0)) { *ok = 0; d = 1.0; }
        row[0] = 1.0 / sqrt(d);
        for (int k = 0; k < HB; ++k) ring[i & (HB - 1)][k] = row[k];
        if (i >= i0)
            for (int k = 0; k < HB; ++k) Lrow[i * HB + k] = row[k];
    }
}

// Lint[g][j][k][lane] = Lrow[(g*32+lane)*B - ov + j][k]; rows outside the
// matrix are identity rows (1/L_ii = 1, no couplings).
__global__ void k_band_interleave(int m, int B, int ov, int steps, long long total,
                                  long long row_off, const double *Lrow, double *Lint) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int lane = (int)(idx & 31);
    const int k = (int)((idx >> 5) & (HB - 1));
    const long long gj = idx >> 8;                  // g * steps + j
    const long long g = gj / steps;
    const int j = (int)(gj - g * steps);
    // row_off: index of the own row 0 within the factored rows (a row-partitioned
    // rank factors its own rows plus a margin of its neighbours')
    const long long i = (g * 32 + lane) * B - ov + j + row_off;
    Lint[idx] = (i >= 0 && i < m) ? Lrow[i * HB + k] : (k == 0 ? 1.0 : 0.0);
}

// xform_ok of a square matrix: every block stages its x slice, the slice covers
// the block's own rows and leaves room for a second slice of the same length
__global__ void k_block_xform_ok(int nblk, const int4 *desc, int *bad) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const int4 a = desc[2 * (size_t)b], c = desc[2 * (size_t)b + 1];
    const bool ok = a.w <= CHUNK && c.y > 0 && c.y <= OPND_MAX && c.y <= XMAX - OPND_MAX &&
                    a.x >= c.x && a.y <= c.x + c.y;
    if (!ok) atomicExch(bad, 1);
}

// column range of every row block -> {c0, xcnt} (second int4 of its descriptor)
__global__ void k_block_colrange(int nblk, int4 *desc, const int *col, int cols) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nblk) return;
    const int4 a = desc[2 * (size_t)warp];
    int lo = 0x7fffffff, hi = -0x7fffffff;
    for (int q = a.z + lane; q < a.z + a.w; q += 32) {
        const int c = col[q];
        lo = min(lo, c); hi = max(hi, c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) {
        int c0 = 0, xcnt = 0;
        if (hi >= lo && a.w <= CHUNK) {
            c0 = lo & ~1;
            xcnt = (hi - c0 + 2) & ~1;
            if (xcnt > XMAX) { c0 = 0; xcnt = 0; }
        }
        // columns outside [0, cols) are ghost entries of a row-partitioned run
        const int ghost = (hi >= lo && (lo < 0 || hi >= cols)) ? 1 : 0;
        desc[2 * (size_t)warp + 1] = make_int4(c0, xcnt, ghost, 0);
    }
}

}  // namespace ipb
