// Host side of the C ABI declared in include/ipsolver_b200.h.
// Owns device memory, builds the transposed / row-normalised copies of A,
// and launches the persistent cooperative kernels of device.cuh.
#include "device.cuh"
#include "../../include/ipsolver_b200.h"

#include <cub/cub.cuh>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>

using namespace ipb;

// ------------------------------------------------------------------ objects
struct ipb_vec {
    double *d = nullptr;
    int64_t len = 0;
};

struct ipb_csr {
    int space = 0;                  // vector space of the columns (0: n, 1: m)
    int rows = 0, cols = 0;
    int64_t nnz = 0;
    int *rowptr = nullptr, *col = nullptr;
    double *val = nullptr;
    int4 *desc = nullptr;
    int nblk = 0;
    int xform_ok = 0;               // see CsrView::xform_ok (square matrices only)
    bool has_t = false;
    int *t_rowptr = nullptr, *t_col = nullptr, *perm = nullptr;
    double *t_val = nullptr;
    int4 *t_desc = nullptr;
    int t_nblk = 0;
};

struct ipb_hess {
    int64_t n = 0, nx = 0;
    ipb_csr *S = nullptr;
    double *D = nullptr;
    double *tail = nullptr;
};

struct ipb_proj {
    ipb_csr *A = nullptr;
    int64_t m = 0, n = 0;                          // own rows of A / own columns
    // "parts" mode (row-partitioned multi-GPU): the five local matrices are
    // given explicitly, already in local column numbering
    bool parts = false;
    ipb_csr *pA_raw = nullptr, *pA_norm = nullptr, *pAt_raw = nullptr, *pAt_norm = nullptr,
            *pM = nullptr;
    double *nval = nullptr, *nt_val = nullptr;     // row-normalised A and A'
    double *dscale = nullptr, *dnorm2 = nullptr;
    double fro = 0.0;
    double orth_tol = 1e-12, rtol = 1e-13;
    int max_refin = 3, maxit = 1000;
    double *v = nullptr, *r = nullptr, *w = nullptr, *p = nullptr, *s = nullptr,
           *rhs = nullptr, *t = nullptr, *q = nullptr;
    // explicit M = A~ A~' (pattern built once, values per refresh)
    int aat_mode = 0;                  // 0 auto, 1 never, 2 always
    bool has_M = false;
    int64_t m_nnz = 0;
    int *m_rowptr = nullptr, *m_col = nullptr;
    double *m_val = nullptr;
    int4 *m_desc = nullptr;
    int m_nblk = 0;
    // banded Cholesky factor of M, applied in overlapping chunks (BandView)
    int band_mode = 0;                 // 0 auto (when M is banded), 1 never
    int band_B_req = 0, band_ov_req = 0;   // 0: defaults
    bool band = false;
    int band_bw = -1, band_B = 0, band_ov = 0, band_steps = 0, band_ngroup = 0;
    double *band_Lrow = nullptr, *band_L = nullptr, *band_y = nullptr;
    int *band_ok = nullptr;
    int band_in_lo = 0, band_in_hi = 0;   // rows of a sweep's input that exist (own-relative)
};

struct ipb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    int sms = 0, blocks = 0;
    unsigned *barrier = nullptr;
    double *partials = nullptr;
    int *abort_flag = nullptr;
    void *d_result = nullptr;          // device result struct (256 B)
    void *h_result = nullptr;          // pinned mirror
    // workspace for n- and m- sized temporaries of the solve kernels
    double *ws = nullptr;
    size_t ws_len = 0;
    double *stage = nullptr;           // device staging for host-vector calls
    size_t stage_len = 0;
    void *cub_tmp = nullptr;
    size_t cub_tmp_bytes = 0;
    int64_t launches = 0;
    std::string err;
    unsigned long long *trace = nullptr;   // barrier timeline of the next launches (ipb_trace_*)
    int trace_cap = 0;
    // vectors of the fused SQP step (ipb_sqp_step): resident across outer iterations
    double *step = nullptr;
    size_t step_len = 0;
    double *step_nv[16] = {nullptr}, *step_mv[8] = {nullptr};   // the step's n- / m-vectors
    int64_t step_n = -1, step_m = -1;
    bool step_has_c = false, step_has_b = false, step_has_box = false;
    double *glue_partials = nullptr, *glue_sums = nullptr;
    unsigned *glue_counter = nullptr;
    // ---- multi-GPU (ipb_comm_*): symmetric arena + peer mappings
    int rank = 0, size = 1;
    char *arena = nullptr;             // [mailbox | flags | seq | vectors]
    size_t arena_bytes = 0, arena_used = 0, arena_vec0 = 0;
    char *peer_base[MAX_RANKS] = {nullptr};
    int64_t own[2] = {0, 0}, own_max[2] = {0, 0}, glo[2] = {0, 0}, ghi[2] = {0, 0};
    int64_t n_lo[2] = {0, 0}, n_hi[2] = {0, 0};       // ghost entries this rank really holds
    size_t gmbox_off[2] = {0, 0};                      // ghost mailboxes inside the arena
    int4 *send_ent[2] = {nullptr, nullptr};
    int send_cnt[2] = {0, 0};
    int64_t nm_global = -1;
    double *dws_n[7] = {nullptr}, *dws_m = nullptr;          // workspace vectors (dist)
    double *dst_n[4] = {nullptr}, *dst_m = nullptr;          // host-staging vectors (dist)
};

static const size_t MBOX_BYTES = 4 * MAX_RANKS * NSLOT * 2 * sizeof(unsigned long long);
static const size_t FLAG_OFF = MBOX_BYTES, SEQ_OFF = FLAG_OFF + 256, GFLAG_OFF = SEQ_OFF + 256,
                    VEC_OFF = GFLAG_OFF + 256;

#define IPB_FAIL(ctx, code, ...)                                              \
    do {                                                                      \
        char _b[512];                                                         \
        snprintf(_b, sizeof _b, __VA_ARGS__);                                 \
        (ctx)->err = _b;                                                      \
        return (code);                                                        \
    } while (0)

#define CU(ctx, call)                                                         \
    do {                                                                      \
        cudaError_t _e = (call);                                              \
        if (_e != cudaSuccess)                                                \
            IPB_FAIL(ctx, IPB_ERR_CUDA, "%s failed: %s (%s:%d)", #call,       \
                     cudaGetErrorString(_e), __FILE__, __LINE__);             \
    } while (0)

#define CHECK(expr)                                                           \
    do {                                                                      \
        int _s = (expr);                                                      \
        if (_s != IPB_OK) return _s;                                          \
    } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
// stride of the sub-vectors carved out of one allocation: keeps every one of
// them 256-byte aligned (TMA bulk copies need 16) with slack behind it
static inline size_t vstride(int64_t len) { return ((size_t)len + PAD + 31) & ~(size_t)31; }

template <typename T>
static int dalloc(ipb_ctx *ctx, T **p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc((void **)p, count * sizeof(T));
    if (e != cudaSuccess)
        IPB_FAIL(ctx, IPB_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s",
                 count * sizeof(T), cudaGetErrorString(e));
    return IPB_OK;
}

static SyncView sync_view(const ipb_ctx *ctx) {
    SyncView s;
    memset(&s, 0, sizeof s);
    s.barrier = ctx->barrier;
    s.partials = ctx->partials;
    s.abort_flag = ctx->abort_flag;
    s.trace = ctx->trace; s.trace_cap = ctx->trace_cap;
    s.C.rank = ctx->rank; s.C.size = ctx->size;
    if (ctx->size > 1) {
        s.C.arena = (double *)ctx->arena;
        s.C.mbox = (double *)ctx->arena;
        s.C.seq = (unsigned *)(ctx->arena + SEQ_OFF);
        for (int q = 0; q < ctx->size; ++q) {
            s.C.peer_arena[q] = (double *)ctx->peer_base[q];
            s.C.peer_mbox[q] = (double *)ctx->peer_base[q];
        }
        s.C.ghost_flag = (unsigned *)(ctx->arena + GFLAG_OFF);
        for (int k = 0; k < 2; ++k) {
            s.C.send[k].ent = ctx->send_ent[k]; s.C.send[k].count = ctx->send_cnt[k];
            s.C.gmbox[k] = (unsigned long long *)(ctx->arena + ctx->gmbox_off[k]);
            for (int q = 0; q < ctx->size; ++q)
                s.C.peer_gmbox[k][q] = (unsigned long long *)(ctx->peer_base[q] + ctx->gmbox_off[k]);
            s.C.gcap[k] = (int)(ctx->glo[k] + ctx->ghi[k]); s.C.gcap_lo[k] = (int)ctx->glo[k];
            s.C.n_lo[k] = (int)ctx->n_lo[k]; s.C.n_hi[k] = (int)ctx->n_hi[k];
            s.C.own_len[k] = (int)ctx->own[k];
        }
    }
    return s;
}

// A solver vector of ``len`` own entries.  Single GPU: its own allocation.
// Multi-GPU: carved from the symmetric arena as [ghost_lo | own | ghost_hi]
// (class = n- or m-space by its length); the returned pointer is the own part.
static int valloc(ipb_ctx *ctx, double **p, int64_t len) {
    if (ctx->size == 1) return dalloc(ctx, p, (size_t)len + PAD);
    int k = (len == ctx->own[0]) ? 0 : (len == ctx->own[1] ? 1 : -1);
    if (k < 0)
        IPB_FAIL(ctx, IPB_ERR_INVALID, "distributed vectors must have the rank's own n (%lld) or m (%lld) "
                 "length, got %lld", (long long)ctx->own[0], (long long)ctx->own[1], (long long)len);
    size_t stride = (((size_t)(ctx->glo[k] + ctx->own_max[k] + ctx->ghi[k]) + PAD + 31) & ~(size_t)31) * 8;
    if (ctx->arena_used + stride > ctx->arena_bytes)
        IPB_FAIL(ctx, IPB_ERR_NOMEM, "symmetric arena exhausted (%zu of %zu bytes)", ctx->arena_used,
                 ctx->arena_bytes);
    *p = (double *)(ctx->arena + ctx->arena_used) + ctx->glo[k];
    ctx->arena_used += stride;
    return IPB_OK;
}
static void vfree(ipb_ctx *ctx, double *p) {
    if (ctx && ctx->size > 1) return;          // arena memory is released with the context
    cudaFree(p);
}

static CsrView csr_view(const ipb_csr *m, const double *val_override = nullptr) {
    CsrView v;
    v.rows = m->rows; v.cols = m->cols; v.nblk = m->nblk; v.desc = m->desc; v.space = m->space;
    v.xform_ok = m->xform_ok;
    v.rowptr = m->rowptr; v.col = m->col;
    v.val = val_override ? val_override : m->val;
    return v;
}

static CsrView csr_view_t(const ipb_csr *m, const double *val_override = nullptr) {
    CsrView v;
    v.rows = m->cols; v.cols = m->rows; v.nblk = m->t_nblk; v.desc = m->t_desc; v.space = 1 - m->space;
    v.xform_ok = 0;
    v.rowptr = m->t_rowptr; v.col = m->t_col;
    v.val = val_override ? val_override : m->t_val;
    return v;
}

// Cut the rows of a CSR matrix into blocks of at most CHUNK nnz and RMAX rows
// (greedy, in row order); a row longer than CHUNK becomes a block of its own.
// rowptr is read back from the device (once per sparsity pattern).
static int build_blocks(ipb_ctx *ctx, const int *d_rowptr, const int *d_col, int rows, int cols,
                        int4 **d_desc, int *nblk);

// number of CTAs for a launch working on ``work`` items: enough to fill the
// machine for large problems, a handful for tiny ones (cheap barriers).
static int grid_for(const ipb_ctx *ctx, int64_t work) {
    int64_t want = (work + THREADS * 4 - 1) / (THREADS * 4);
    return (int)std::max<int64_t>(1, std::min<int64_t>(ctx->blocks, want));
}

static int build_blocks(ipb_ctx *ctx, const int *d_rowptr, const int *d_col, int rows, int cols,
                        int4 **d_desc, int *nblk) {
    std::vector<int> rp((size_t)rows + 1);
    CU(ctx, cudaMemcpyAsync(rp.data(), d_rowptr, ((size_t)rows + 1) * sizeof(int),
                            cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<int4> desc;
    desc.reserve((size_t)rows / 256 + 16);
    int r0 = 0;
    while (r0 < rows) {
        int r1 = r0 + 1;
        const int base = rp[r0];
        if (rp[r1] - base <= CHUNK) {
            const int rlim = std::min(rows, r0 + RMAX);
            // largest r1 with rp[r1] - base <= CHUNK (rp is non-decreasing)
            int hi = (int)(std::upper_bound(rp.begin() + r1, rp.begin() + rlim + 1,
                                            base + CHUNK) - rp.begin()) - 1;
            r1 = std::max(r1, hi);
        }
        desc.push_back(make_int4(r0, r1, base, rp[r1] - base));
        desc.push_back(make_int4(0, 0, 0, 0));      // {c0, xcnt}: filled on the device
        r0 = r1;
    }
    *nblk = (int)(desc.size() / 2);
    CHECK(dalloc(ctx, d_desc, desc.size()));
    if (!desc.empty()) {
        CU(ctx, cudaMemcpyAsync(*d_desc, desc.data(), desc.size() * sizeof(int4),
                                cudaMemcpyHostToDevice, ctx->stream));
        k_block_colrange<<<cdiv((int64_t)*nblk * 32, 256), 256, 0, ctx->stream>>>(*nblk, *d_desc, d_col, cols);
        ctx->launches += 1;
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IPB_OK;
}

template <typename Args>
static int launch_coop(ipb_ctx *ctx, void (*kernel)(const Args), const Args &args,
                       int grid, float *ms) {
    CU(ctx, cudaMemsetAsync(ctx->barrier, 0, sizeof(unsigned), ctx->stream));
    if (ctx->trace) CU(ctx, cudaMemsetAsync(ctx->trace, 0, sizeof(unsigned long long), ctx->stream));
    void *params[] = {(void *)&args};
    if (ms) CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    CU(ctx, cudaLaunchCooperativeKernel((void *)kernel, dim3(grid), dim3(THREADS),
                                        params, sizeof(Smem), ctx->stream));
    ctx->launches += 1;
    if (ms) {
        CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CU(ctx, cudaEventSynchronize(ctx->ev1));
        CU(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    }
    return IPB_OK;
}

static int ensure_ws(ipb_ctx *ctx, size_t len) {
    if (ctx->ws_len >= len) return IPB_OK;
    if (ctx->ws) cudaFree(ctx->ws);
    ctx->ws = nullptr; ctx->ws_len = 0;
    CHECK(dalloc(ctx, &ctx->ws, len));
    ctx->ws_len = len;
    return IPB_OK;
}

static int ensure_stage(ipb_ctx *ctx, size_t len) {
    if (ctx->stage_len >= len) return IPB_OK;
    if (ctx->stage) cudaFree(ctx->stage);
    ctx->stage = nullptr; ctx->stage_len = 0;
    CHECK(dalloc(ctx, &ctx->stage, len));
    ctx->stage_len = len;
    return IPB_OK;
}

static int ensure_cub(ipb_ctx *ctx, size_t bytes) {
    if (ctx->cub_tmp_bytes >= bytes) return IPB_OK;
    if (ctx->cub_tmp) cudaFree(ctx->cub_tmp);
    ctx->cub_tmp = nullptr; ctx->cub_tmp_bytes = 0;
    char *p;
    CHECK(dalloc(ctx, &p, bytes));
    ctx->cub_tmp = p; ctx->cub_tmp_bytes = bytes;
    return IPB_OK;
}

// ------------------------------------------------------------------ context
extern "C" int ipb_ctx_create(int device, ipb_ctx **out) {
    if (!out) return IPB_ERR_INVALID;
    *out = nullptr;
    ipb_ctx *ctx = new ipb_ctx();
    *out = ctx;          // returned even on failure so the message is readable
    ctx->device = device;
    CU(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(ctx, cudaGetDeviceProperties(&prop, device));
    if (!prop.cooperativeLaunch)
        IPB_FAIL(ctx, IPB_ERR_CUDA, "device %d lacks cooperative launch", device);
    ctx->sms = prop.multiProcessorCount;
    CU(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CU(ctx, cudaEventCreate(&ctx->ev0));
    CU(ctx, cudaEventCreate(&ctx->ev1));
    CU(ctx, cudaEventCreate(&ctx->ev2));
    CU(ctx, cudaEventCreate(&ctx->ev3));
    int occ = 1 << 30, o;
    const void *kernels[] = {(const void *)k_projected_cg, (const void *)k_projector,
                             (const void *)k_dogleg, (const void *)k_matvec,
                             (const void *)k_intersect};
    for (const void *k : kernels) {
        CU(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(Smem)));
        CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k, THREADS, sizeof(Smem)));
        occ = std::min(occ, o);
    }
    if (occ < 1) IPB_FAIL(ctx, IPB_ERR_CUDA, "kernels do not fit on an SM");
    int per_sm = occ;
    const char *env = getenv("IPB_BLOCKS_PER_SM");
    if (env && atoi(env) > 0) per_sm = std::min(occ, atoi(env));
    ctx->blocks = std::min(MAX_GRID, per_sm * ctx->sms);
    CHECK(dalloc(ctx, &ctx->barrier, 1));
    CHECK(dalloc(ctx, &ctx->partials, (size_t)2 * NSLOT * MAX_GRID));
    CHECK(dalloc(ctx, &ctx->abort_flag, 1));
    CU(ctx, cudaMemset(ctx->barrier, 0, sizeof(unsigned)));
    CU(ctx, cudaMemset(ctx->abort_flag, 0, sizeof(int)));
    CU(ctx, cudaMalloc(&ctx->d_result, 256));
    CU(ctx, cudaMallocHost(&ctx->h_result, 256));
    return IPB_OK;
}

extern "C" int ipb_ctx_destroy(ipb_ctx *ctx) {
    if (!ctx) return IPB_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->barrier); cudaFree(ctx->partials); cudaFree(ctx->abort_flag);
    cudaFree(ctx->d_result); cudaFreeHost(ctx->h_result);
    cudaFree(ctx->ws); cudaFree(ctx->stage); cudaFree(ctx->cub_tmp); cudaFree(ctx->trace);
    cudaFree(ctx->step); cudaFree(ctx->glue_partials); cudaFree(ctx->glue_sums); cudaFree(ctx->glue_counter);
    for (int q = 0; q < ctx->size; ++q)
        if (q != ctx->rank && ctx->peer_base[q]) cudaIpcCloseMemHandle(ctx->peer_base[q]);
    cudaFree(ctx->arena); cudaFree(ctx->send_ent[0]); cudaFree(ctx->send_ent[1]);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->ev3) cudaEventDestroy(ctx->ev3);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return IPB_OK;
}

extern "C" const char *ipb_last_error(const ipb_ctx *ctx) {
    return ctx ? ctx->err.c_str() : "null context";
}

extern "C" int ipb_ctx_grid(const ipb_ctx *ctx, int *blocks, int *threads, int *sms) {
    if (!ctx) return IPB_ERR_INVALID;
    if (blocks) *blocks = ctx->blocks;
    if (threads) *threads = THREADS;
    if (sms) *sms = ctx->sms;
    return IPB_OK;
}

extern "C" int64_t ipb_ctx_launches(const ipb_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int ipb_ctx_sync(ipb_ctx *ctx) {
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IPB_OK;
}

extern "C" int ipb_timer_start(ipb_ctx *ctx) {
    if (!ctx) return IPB_ERR_INVALID;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev2, ctx->stream));
    return IPB_OK;
}

extern "C" int ipb_timer_stop(ipb_ctx *ctx, double *ms) {
    if (!ctx || !ms) return IPB_ERR_INVALID;
    float f = 0.f;
    CU(ctx, cudaEventRecord(ctx->ev3, ctx->stream));
    CU(ctx, cudaEventSynchronize(ctx->ev3));
    CU(ctx, cudaEventElapsedTime(&f, ctx->ev2, ctx->ev3));
    *ms = f;
    return IPB_OK;
}

// ------------------------------------------------------------------ barrier timeline
extern "C" int ipb_trace_enable(ipb_ctx *ctx, int capacity) {
    if (!ctx || capacity < 0) return IPB_ERR_INVALID;
    cudaFree(ctx->trace);
    ctx->trace = nullptr; ctx->trace_cap = 0;
    if (capacity > 0) {
        CHECK(dalloc(ctx, &ctx->trace, (size_t)capacity + 1));
        CU(ctx, cudaMemset(ctx->trace, 0, ((size_t)capacity + 1) * sizeof(unsigned long long)));
        ctx->trace_cap = capacity + 1;
    }
    return IPB_OK;
}
extern "C" int ipb_trace_read(ipb_ctx *ctx, uint64_t *out, int capacity, int *count) {
    if (!ctx || !out || !count || !ctx->trace) return IPB_ERR_INVALID;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned long long k = 0;
    CU(ctx, cudaMemcpy(&k, ctx->trace, sizeof k, cudaMemcpyDeviceToHost));
    const int have = (int)std::min<unsigned long long>(k, (unsigned long long)std::min(capacity, ctx->trace_cap - 1));
    if (have > 0)
        CU(ctx, cudaMemcpy(out, ctx->trace + 1, (size_t)have * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    *count = (int)k;
    return IPB_OK;
}

// ------------------------------------------------------------------ vectors
extern "C" int ipb_vec_create(ipb_ctx *ctx, int64_t len, ipb_vec **out) {
    if (!ctx || !out || len < 0) return IPB_ERR_INVALID;
    ipb_vec *v = new ipb_vec();
    v->len = len;
    int s = valloc(ctx, &v->d, len);           // slack behind it: bulk copies round sizes up
    if (s != IPB_OK) { delete v; return s; }
    *out = v;
    return IPB_OK;
}
extern "C" int ipb_vec_destroy(ipb_ctx *ctx, ipb_vec *v) {
    if (v) { vfree(ctx, v->d); delete v; }
    return IPB_OK;
}
extern "C" int ipb_vec_upload(ipb_ctx *ctx, ipb_vec *v, const double *host) {
    CU(ctx, cudaMemcpyAsync(v->d, host, v->len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IPB_OK;
}
extern "C" int ipb_vec_download(ipb_ctx *ctx, const ipb_vec *v, double *host) {
    CU(ctx, cudaMemcpyAsync(host, v->d, v->len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IPB_OK;
}
extern "C" int64_t ipb_vec_len(const ipb_vec *v) { return v ? v->len : 0; }
extern "C" void *ipb_vec_ptr(ipb_vec *v) { return v ? v->d : nullptr; }

// ------------------------------------------------------------------ CSR
static int build_transpose(ipb_ctx *ctx, ipb_csr *m) {
    const int64_t nnz = m->nnz;
    const int TB = 256;
    CHECK(dalloc(ctx, &m->t_rowptr, (size_t)m->cols + 1 + PAD));
    CHECK(dalloc(ctx, &m->t_col, (size_t)nnz + PAD));
    CHECK(dalloc(ctx, &m->t_val, (size_t)nnz + PAD));
    CHECK(dalloc(ctx, &m->perm, (size_t)nnz));
    CU(ctx, cudaMemsetAsync(m->t_rowptr, 0, ((size_t)m->cols + 1) * sizeof(int), ctx->stream));
    if (nnz > 0) {
        int *rowidx, *iota, *keys_out;
        CHECK(dalloc(ctx, &rowidx, (size_t)nnz));
        CHECK(dalloc(ctx, &iota, (size_t)nnz));
        CHECK(dalloc(ctx, &keys_out, (size_t)nnz));
        k_expand_rows<<<cdiv(m->rows, TB), TB, 0, ctx->stream>>>(m->rows, m->rowptr, rowidx);
        k_iota<<<cdiv(nnz, TB), TB, 0, ctx->stream>>>(nnz, iota);
        k_count_cols<<<cdiv(nnz, TB), TB, 0, ctx->stream>>>(nnz, m->col, m->t_rowptr);
        ctx->launches += 3;
        int bits = 1;
        while ((1ll << bits) < (int64_t)m->cols && bits < 31) ++bits;
        size_t tmp = 0;
        CU(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, m->col, keys_out, iota, m->perm,
                                                (int)nnz, 0, bits, ctx->stream));
        CHECK(ensure_cub(ctx, tmp));
        CU(ctx, cub::DeviceRadixSort::SortPairs(ctx->cub_tmp, tmp, m->col, keys_out, iota, m->perm,
                                                (int)nnz, 0, bits, ctx->stream));
        // t_col[k] = row of the k-th entry in column order
        k_gather_int<<<cdiv(nnz, TB), TB, 0, ctx->stream>>>(nnz, m->perm, rowidx, m->t_col);
        k_gather<<<cdiv(nnz, TB), TB, 0, ctx->stream>>>(nnz, m->perm, m->val, m->t_val);
        ctx->launches += 2;
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(rowidx); cudaFree(iota); cudaFree(keys_out);
    }
    // rowptr of the transpose: inclusive scan of the shifted column counts
    {
        size_t tmp = 0;
        CU(ctx, cub::DeviceScan::InclusiveSum(nullptr, tmp, m->t_rowptr, m->t_rowptr,
                                              m->cols + 1, ctx->stream));
        CHECK(ensure_cub(ctx, tmp));
        CU(ctx, cub::DeviceScan::InclusiveSum(ctx->cub_tmp, tmp, m->t_rowptr, m->t_rowptr,
                                              m->cols + 1, ctx->stream));
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CHECK(build_blocks(ctx, m->t_rowptr, m->t_col, m->cols, m->rows, &m->t_desc, &m->t_nblk));
    m->has_t = true;
    return IPB_OK;
}

extern "C" int ipb_csr_create(ipb_ctx *ctx, int64_t rows, int64_t cols, int64_t nnz,
                              const void *indptr, const void *indices, int index_bytes,
                              const double *data, int with_transpose, ipb_csr **out) {
    if (!ctx || !out) return IPB_ERR_INVALID;
    if (rows < 0 || cols < 0 || nnz < 0 || (index_bytes != 4 && index_bytes != 8))
        IPB_FAIL(ctx, IPB_ERR_INVALID, "ipb_csr_create: bad shape or index width");
    if (rows >= (1ll << 31) - 1 || cols >= (1ll << 31) - 1 || nnz >= (1ll << 31) - 1)
        IPB_FAIL(ctx, IPB_ERR_INVALID, "ipb_csr_create: dimensions exceed int32 per-GPU limit");
    CU(ctx, cudaSetDevice(ctx->device));
    ipb_csr *m = new ipb_csr();
    m->rows = (int)rows; m->cols = (int)cols; m->nnz = nnz;
    int s;
    if ((s = dalloc(ctx, &m->rowptr, (size_t)rows + 1 + PAD)) != IPB_OK) return s;
    if ((s = dalloc(ctx, &m->col, (size_t)nnz + PAD)) != IPB_OK) return s;
    if ((s = dalloc(ctx, &m->val, (size_t)nnz + PAD)) != IPB_OK) return s;
    if (index_bytes == 4) {
        CU(ctx, cudaMemcpyAsync(m->rowptr, indptr, ((size_t)rows + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (nnz) CU(ctx, cudaMemcpyAsync(m->col, indices, (size_t)nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        long long *tmp;
        size_t cnt = (size_t)std::max<int64_t>(rows + 1, nnz);
        if ((s = dalloc(ctx, &tmp, cnt)) != IPB_OK) return s;
        CU(ctx, cudaMemcpyAsync(tmp, indptr, ((size_t)rows + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        k_narrow_i64<<<cdiv(rows + 1, 256), 256, 0, ctx->stream>>>(rows + 1, tmp, m->rowptr);
        ctx->launches += 1;
        if (nnz) {
            CU(ctx, cudaMemcpyAsync(tmp, indices, (size_t)nnz * 8, cudaMemcpyHostToDevice, ctx->stream));
            k_narrow_i64<<<cdiv(nnz, 256), 256, 0, ctx->stream>>>(nnz, tmp, m->col);
            ctx->launches += 1;
        }
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(tmp);
    }
    if (nnz) CU(ctx, cudaMemcpyAsync(m->val, data, (size_t)nnz * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (rows == 0) CU(ctx, cudaMemsetAsync(m->rowptr, 0, sizeof(int), ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if ((s = build_blocks(ctx, m->rowptr, m->col, m->rows, m->cols, &m->desc, &m->nblk)) != IPB_OK) return s;
    if (rows == cols && m->nblk > 0) {
        int *d_bad = nullptr, bad = 0;
        if ((s = dalloc(ctx, &d_bad, 1)) != IPB_OK) return s;
        CU(ctx, cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
        k_block_xform_ok<<<cdiv(m->nblk, 256), 256, 0, ctx->stream>>>(m->nblk, m->desc, d_bad);
        ctx->launches += 1;
        CU(ctx, cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(d_bad);
        m->xform_ok = bad ? 0 : 1;
    }
    if (with_transpose) {
        s = build_transpose(ctx, m);
        if (s != IPB_OK) return s;
    }
    *out = m;
    return IPB_OK;
}

extern "C" int ipb_csr_update_values(ipb_ctx *ctx, ipb_csr *m, const double *data) {
    if (!ctx || !m) return IPB_ERR_INVALID;
    if (m->nnz) {
        CU(ctx, cudaMemcpyAsync(m->val, data, (size_t)m->nnz * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (m->has_t) {
            k_gather<<<cdiv(m->nnz, 256), 256, 0, ctx->stream>>>(m->nnz, m->perm, m->val, m->t_val);
            ctx->launches += 1;
        }
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IPB_OK;
}

extern "C" int ipb_csr_destroy(ipb_ctx *ctx, ipb_csr *m) {
    (void)ctx;
    if (!m) return IPB_OK;
    cudaFree(m->rowptr); cudaFree(m->col); cudaFree(m->val); cudaFree(m->desc);
    cudaFree(m->t_desc); cudaFree(m->t_rowptr); cudaFree(m->t_col); cudaFree(m->t_val); cudaFree(m->perm);
    delete m;
    return IPB_OK;
}

static int run_matvec(ipb_ctx *ctx, const CsrView &M, const double *dx, double *dy) {
    MatvecArgs a;
    memset(&a, 0, sizeof a);
    a.kind = 0; a.M = M; a.S = sync_view(ctx); a.x = dx; a.y = dy;
    int grid = grid_for(ctx, std::max<int64_t>(M.rows, 1) * 4);
    return launch_coop(ctx, k_matvec, a, grid, nullptr);
}

#define SINGLE_GPU_ONLY(ctx)                                                         \
    do {                                                                             \
        if ((ctx)->size > 1)                                                         \
            IPB_FAIL(ctx, IPB_ERR_INVALID, "%s takes host vectors and is single-GPU; " \
                     "row-partitioned runs use the *_dev entry points", __func__);   \
    } while (0)

static int host_in_out(ipb_ctx *ctx, int64_t nin, int64_t nout, const double *hin,
                       double **din, double **dout) {
    SINGLE_GPU_ONLY(ctx);
    CHECK(ensure_stage(ctx, vstride(nin) + vstride(nout)));
    *din = ctx->stage; *dout = ctx->stage + vstride(nin);
    if (nin) CU(ctx, cudaMemcpyAsync(*din, hin, (size_t)nin * 8, cudaMemcpyHostToDevice, ctx->stream));
    return IPB_OK;
}

static int host_out(ipb_ctx *ctx, int64_t nout, const double *dout, double *hout) {
    if (nout) CU(ctx, cudaMemcpyAsync(hout, dout, (size_t)nout * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IPB_OK;
}

extern "C" int ipb_csr_matvec(ipb_ctx *ctx, const ipb_csr *m, const double *x, double *y) {
    if (!ctx || !m) return IPB_ERR_INVALID;
    double *dx, *dy;
    CHECK(host_in_out(ctx, m->cols, m->rows, x, &dx, &dy));
    CHECK(run_matvec(ctx, csr_view(m), dx, dy));
    return host_out(ctx, m->rows, dy, y);
}

extern "C" int ipb_csr_rmatvec(ipb_ctx *ctx, const ipb_csr *m, const double *y, double *x) {
    if (!ctx || !m) return IPB_ERR_INVALID;
    if (!m->has_t) IPB_FAIL(ctx, IPB_ERR_INVALID, "matrix was created without its transpose");
    double *dy, *dx;
    CHECK(host_in_out(ctx, m->rows, m->cols, y, &dy, &dx));
    CHECK(run_matvec(ctx, csr_view_t(m), dy, dx));
    return host_out(ctx, m->cols, dx, x);
}

extern "C" int ipb_csr_matvec_dev(ipb_ctx *ctx, const ipb_csr *m, const ipb_vec *x, ipb_vec *y) {
    if (!ctx || !m || !x || !y || x->len != m->cols || y->len != m->rows)
        return IPB_ERR_INVALID;
    return run_matvec(ctx, csr_view(m), x->d, y->d);
}

extern "C" int ipb_csr_rmatvec_dev(ipb_ctx *ctx, const ipb_csr *m, const ipb_vec *y, ipb_vec *x) {
    if (!ctx || !m || !x || !y || x->len != m->cols || y->len != m->rows || !m->has_t)
        return IPB_ERR_INVALID;
    return run_matvec(ctx, csr_view_t(m), y->d, x->d);
}

static int device_sum(ipb_ctx *ctx, const double *d, int64_t n, double *out) {
    double *dres = (double *)ctx->d_result;
    size_t tmp = 0;
    CU(ctx, cub::DeviceReduce::Sum(nullptr, tmp, d, dres, (int)n, ctx->stream));
    CHECK(ensure_cub(ctx, tmp));
    CU(ctx, cub::DeviceReduce::Sum(ctx->cub_tmp, tmp, d, dres, (int)n, ctx->stream));
    CU(ctx, cudaMemcpyAsync(out, dres, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IPB_OK;
}

extern "C" int ipb_csr_fro_norm(ipb_ctx *ctx, const ipb_csr *m, double *out) {
    if (!ctx || !m || !out) return IPB_ERR_INVALID;
    double *d2, *ds;
    CHECK(dalloc(ctx, &d2, (size_t)m->rows));
    CHECK(dalloc(ctx, &ds, (size_t)m->rows));
    double s = 0.0;
    if (m->rows > 0) {
        k_row_norms<<<cdiv(m->rows, 256), 256, 0, ctx->stream>>>(m->rows, m->rowptr, m->val, d2, ds);
        ctx->launches += 1;
        CHECK(device_sum(ctx, d2, m->rows, &s));
    }
    cudaFree(d2); cudaFree(ds);
    *out = std::sqrt(s);
    return IPB_OK;
}

// ------------------------------------------------------------------ Hessian
extern "C" int ipb_hess_create(ipb_ctx *ctx, int64_t n, int64_t n_x, ipb_hess **out) {
    if (!ctx || !out || n < 0 || n_x < 0 || n_x > n) return IPB_ERR_INVALID;
    ipb_hess *h = new ipb_hess();
    h->n = n; h->nx = n_x;
    *out = h;
    return IPB_OK;
}
extern "C" int ipb_hess_set_csr(ipb_ctx *ctx, ipb_hess *h, ipb_csr *s) {
    if (!ctx || !h) return IPB_ERR_INVALID;
    if (s && (s->rows != h->nx || (ctx->size == 1 && s->cols != h->nx)))
        IPB_FAIL(ctx, IPB_ERR_INVALID, "Hessian CSR term must be n_x by n_x");
    h->S = s;
    return IPB_OK;
}
extern "C" int ipb_hess_set_dense(ipb_ctx *ctx, ipb_hess *h, const double *d) {
    if (!ctx || !h) return IPB_ERR_INVALID;
    if (!d) { cudaFree(h->D); h->D = nullptr; return IPB_OK; }
    size_t cnt = (size_t)h->nx * (size_t)h->nx;
    if (!h->D) CHECK(dalloc(ctx, &h->D, cnt));
    CU(ctx, cudaMemcpyAsync(h->D, d, cnt * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IPB_OK;
}
extern "C" int ipb_hess_set_diag_tail(ipb_ctx *ctx, ipb_hess *h, const double *diag) {
    if (!ctx || !h) return IPB_ERR_INVALID;
    if (!diag) { cudaFree(h->tail); h->tail = nullptr; return IPB_OK; }
    size_t cnt = (size_t)(h->n - h->nx);
    if (!h->tail) CHECK(dalloc(ctx, &h->tail, cnt));
    if (cnt) CU(ctx, cudaMemcpyAsync(h->tail, diag, cnt * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IPB_OK;
}
extern "C" int ipb_hess_destroy(ipb_ctx *ctx, ipb_hess *h) {
    (void)ctx;
    if (!h) return IPB_OK;
    cudaFree(h->D); cudaFree(h->tail);
    delete h;
    return IPB_OK;
}

static HessView hess_view(const ipb_hess *h) {
    HessView v;
    memset(&v, 0, sizeof v);
    v.n = (int)h->n; v.nx = (int)h->nx;
    v.has_csr = h->S != nullptr;
    v.has_dense = h->D != nullptr;
    v.has_tail = h->tail != nullptr;
    if (h->S) v.S = csr_view(h->S);
    v.D = h->D; v.tail = h->tail;
    return v;
}

static int run_hess(ipb_ctx *ctx, const ipb_hess *h, const double *dx, double *dy) {
    MatvecArgs a;
    memset(&a, 0, sizeof a);
    a.kind = 1; a.H = hess_view(h); a.S = sync_view(ctx); a.x = dx; a.y = dy;
    int64_t work = std::max<int64_t>(h->n * 4, h->D ? h->nx * h->nx / 8 : 0);
    return launch_coop(ctx, k_matvec, a, grid_for(ctx, work), nullptr);
}

extern "C" int ipb_hess_matvec(ipb_ctx *ctx, const ipb_hess *h, const double *x, double *y) {
    if (!ctx || !h) return IPB_ERR_INVALID;
    double *dx, *dy;
    CHECK(host_in_out(ctx, h->n, h->n, x, &dx, &dy));
    CHECK(run_hess(ctx, h, dx, dy));
    return host_out(ctx, h->n, dy, y);
}
extern "C" int ipb_hess_matvec_dev(ipb_ctx *ctx, const ipb_hess *h, const ipb_vec *x, ipb_vec *y) {
    if (!ctx || !h || !x || !y || x->len != h->n || y->len != h->n) return IPB_ERR_INVALID;
    return run_hess(ctx, h, x->d, y->d);
}

// ------------------------------------------------------------------ projector
static void free_aat(ipb_proj *p) {
    cudaFree(p->m_rowptr); cudaFree(p->m_col); cudaFree(p->m_val); cudaFree(p->m_desc);
    p->m_rowptr = p->m_col = nullptr; p->m_val = nullptr; p->m_desc = nullptr;
    p->has_M = false; p->m_nnz = 0; p->m_nblk = 0;
}

// Sparsity pattern of M = A A' on the device: expand every pair of rows that
// share a column into a 64-bit (row, col) key, radix-sort, unique.  The
// explicit product is kept only when it moves fewer bytes per application
// than A and A' do (12 nnz(M) against 24 nnz(A) + 16 n), unless forced.
static int build_aat_pattern(ipb_ctx *ctx, ipb_proj *p) {
    free_aat(p);
    ipb_csr *A = p->A;
    const int m = A->rows, n = A->cols;
    if (p->aat_mode == 1 || m == 0 || A->nnz == 0) return IPB_OK;
    const int TB = 256;
    long long *cnt = nullptr, *ptr = nullptr;
    CHECK(dalloc(ctx, &cnt, (size_t)m + 1));
    CHECK(dalloc(ctx, &ptr, (size_t)m + 1));
    CU(ctx, cudaMemsetAsync(cnt, 0, ((size_t)m + 1) * sizeof(long long), ctx->stream));
    k_aat_count<<<cdiv(m, TB), TB, 0, ctx->stream>>>(m, A->rowptr, A->col, A->t_rowptr, cnt);
    ctx->launches += 1;
    size_t tmp = 0;
    CU(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, cnt, ptr, m + 1, ctx->stream));
    CHECK(ensure_cub(ctx, tmp));
    CU(ctx, cub::DeviceScan::ExclusiveSum(ctx->cub_tmp, tmp, cnt, ptr, m + 1, ctx->stream));
    long long ncand = 0;
    CU(ctx, cudaMemcpyAsync(&ncand, ptr + m, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    const long long budget = 2 * (long long)A->nnz + n;     // nnz(M) that still pays
    const long long cap = p->aat_mode == 2 ? (1ll << 31) - 2 : 8 * budget;
    if (ncand > cap || ncand >= (1ll << 31) - 1) {
        cudaFree(cnt); cudaFree(ptr);
        if (p->aat_mode == 2)
            IPB_FAIL(ctx, IPB_ERR_NOMEM, "explicit A A' would have %lld candidate entries", ncand);
        return IPB_OK;
    }
    unsigned long long *keys = nullptr, *keys2 = nullptr;
    int *d_num = nullptr;
    CHECK(dalloc(ctx, &keys, (size_t)ncand));
    CHECK(dalloc(ctx, &keys2, (size_t)ncand));
    CHECK(dalloc(ctx, &d_num, 1));
    k_aat_expand<<<cdiv(m, TB), TB, 0, ctx->stream>>>(m, A->rowptr, A->col, A->t_rowptr,
                                                      A->t_col, ptr, keys);
    ctx->launches += 1;
    int bits = 1;
    while ((1ll << bits) < (long long)m && bits < 31) ++bits;
    CU(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tmp, keys, keys2, (int)ncand, 0, 32 + bits,
                                           ctx->stream));
    CHECK(ensure_cub(ctx, tmp));
    CU(ctx, cub::DeviceRadixSort::SortKeys(ctx->cub_tmp, tmp, keys, keys2, (int)ncand, 0, 32 + bits,
                                           ctx->stream));
    CU(ctx, cub::DeviceSelect::Unique(nullptr, tmp, keys2, keys, d_num, (int)ncand, ctx->stream));
    CHECK(ensure_cub(ctx, tmp));
    CU(ctx, cub::DeviceSelect::Unique(ctx->cub_tmp, tmp, keys2, keys, d_num, (int)ncand, ctx->stream));
    int nnzM = 0;
    CU(ctx, cudaMemcpyAsync(&nnzM, d_num, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    int status = IPB_OK;
    if (p->aat_mode == 2 || (long long)nnzM <= budget) {
        p->m_nnz = nnzM;
        if ((status = dalloc(ctx, &p->m_rowptr, (size_t)m + 1 + PAD)) == IPB_OK &&
            (status = dalloc(ctx, &p->m_col, (size_t)nnzM + PAD)) == IPB_OK &&
            (status = dalloc(ctx, &p->m_val, (size_t)nnzM + PAD)) == IPB_OK) {
            k_aat_rows<<<cdiv((long long)nnzM + 1, TB), TB, 0, ctx->stream>>>(
                nnzM, m, keys, p->m_rowptr, p->m_col);
            ctx->launches += 1;
            cudaError_t e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); status = IPB_ERR_CUDA; }
            if (status == IPB_OK)
                status = build_blocks(ctx, p->m_rowptr, p->m_col, m, m, &p->m_desc, &p->m_nblk);
            if (status == IPB_OK) p->has_M = true;
        }
        if (status != IPB_OK) free_aat(p);
    }
    cudaFree(cnt); cudaFree(ptr); cudaFree(keys); cudaFree(keys2); cudaFree(d_num);
    return status;
}

// ------------------------------------------------------------------ banded factor of M
static void free_band(ipb_proj *p) {
    cudaFree(p->band_Lrow); cudaFree(p->band_L); cudaFree(p->band_y); cudaFree(p->band_ok);
    p->band_Lrow = p->band_L = p->band_y = nullptr; p->band_ok = nullptr;
    p->band = false; p->band_bw = -1; p->band_ngroup = 0;
}

// Owned rows per chunk.  A sweep costs max(bytes moved / HBM bandwidth, steps x
// ~150 ns): 176 rows is the measured optimum when the bandwidth binds (m = 5e6:
// 128 -> 186 us, 176 -> 171, 256 -> 177: longer chunks leave too few warps to
// cover the per-step latency), and below ~3 groups of 32 chunks per SM the
// chain length (2 B + 3 overlap steps) binds, so small problems (or the ranks
// of a row-partitioned one) get shorter chunks, down to 64 rows.
static int band_chunk_rows(const ipb_ctx *ctx, int64_t m) {
    const int64_t groups = 3 * (int64_t)ctx->sms;
    int B = (int)((m + 32 * groups - 1) / (32 * groups));
    B = std::max(64, std::min(176, B));
    return (B + 7) & ~7;
}

// Decide whether M is banded (half-bandwidth < HB) and size the chunked
// factor: B owned rows per chunk such that the machine gets ~8 groups of 32
// chunks per SM (enough bulk copies in flight to saturate HBM) without the
// overlap (ov rows each side, read redundantly) dominating.
static int build_band(ipb_ctx *ctx, ipb_proj *p) {
    free_band(p);
    if (p->parts || !p->has_M || p->band_mode == 1 || p->m < 1) return IPB_OK;
    const int m = (int)p->m;
    int *d_bw = nullptr;
    CHECK(dalloc(ctx, &d_bw, 1));
    CU(ctx, cudaMemsetAsync(d_bw, 0, sizeof(int), ctx->stream));
    k_band_width<<<cdiv(m, 256), 256, 0, ctx->stream>>>(m, p->m_rowptr, p->m_col, d_bw);
    ctx->launches += 1;
    int bw = 0;
    CU(ctx, cudaMemcpyAsync(&bw, d_bw, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_bw);
    p->band_bw = bw;
    if (bw < 1 || bw >= HB) return IPB_OK;      // diagonal M: CG takes one step; wide M: CG
    int ov = p->band_ov_req > 0 ? p->band_ov_req : 48;
    ov = (ov + 3) & ~3;
    int B = p->band_B_req;
    if (B <= 0) {
        B = band_chunk_rows(ctx, m);
    }
    B = (B + 3) & ~3;
    const int64_t nchunk = (m + (int64_t)B - 1) / B;
    const int64_t ngroup = (nchunk + 31) / 32;
    p->band_B = B; p->band_ov = ov; p->band_steps = B + 2 * ov; p->band_ngroup = (int)ngroup;
    CHECK(dalloc(ctx, &p->band_Lrow, (size_t)m * HB));
    CHECK(dalloc(ctx, &p->band_L, (size_t)ngroup * p->band_steps * HB * 32));
    CHECK(dalloc(ctx, &p->band_y, (size_t)ngroup * (B + ov) * 32));
    CHECK(dalloc(ctx, &p->band_ok, 1));
    p->band_in_lo = 0; p->band_in_hi = m;
    p->band = true;
    return IPB_OK;
}

static int refresh_band(ipb_ctx *ctx, ipb_proj *p) {
    if (!p->band) return IPB_OK;
    const int m = (int)p->m, B = p->band_B;
    const int one = 1;
    CU(ctx, cudaMemcpyAsync(p->band_ok, &one, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    const int64_t nchunk = (m + (int64_t)B - 1) / B;
    k_band_factor<<<cdiv(nchunk, 64), 64, 0, ctx->stream>>>(m, p->m_rowptr, p->m_col, p->m_val, B,
                                                            48, p->band_Lrow, p->band_ok);
    const long long total = (long long)p->band_ngroup * p->band_steps * HB * 32;
    k_band_interleave<<<cdiv(total, 256), 256, 0, ctx->stream>>>(m, B, p->band_ov, p->band_steps,
                                                                 total, 0, p->band_Lrow, p->band_L);
    ctx->launches += 2;
    return IPB_OK;
}

extern "C" int ipb_proj_set_band(ipb_ctx *ctx, ipb_proj *p, int mode, int chunk_rows, int overlap) {
    if (!ctx || !p || mode < 0 || mode > 1 || chunk_rows < 0 || overlap < 0) return IPB_ERR_INVALID;
    p->band_mode = mode; p->band_B_req = chunk_rows; p->band_ov_req = overlap;
    CHECK(build_band(ctx, p));
    CHECK(refresh_band(ctx, p));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IPB_OK;
}

// Row-partitioned projector: the banded factor of the rank's own rows of M.
// M_ext holds rows [own_lo - ext_off, ...) of M as a standalone square CSR
// (columns relative to its first row, entries outside dropped): the own rows
// plus a margin on each side, so that the factor recurrence, started at the
// margin's edge, has forgotten its start where the sweeps use it.  The sweeps
// read ghost_below / ghost_above rows of their input beyond the own block:
// the m-space ghost zones must hold exactly those rows (contiguous).
extern "C" int ipb_proj_set_band_parts(ipb_ctx *ctx, ipb_proj *p, const ipb_csr *M_ext,
                                       int64_t ext_off, int ghost_below, int ghost_above,
                                       int chunk_rows, int overlap) {
    if (!ctx || !p || !p->parts || !M_ext || ext_off < 0 || ghost_below < 0 || ghost_above < 0 ||
        chunk_rows < 0 || overlap < 0)
        return IPB_ERR_INVALID;
    free_band(p);
    const int m = (int)p->m, m_ext = M_ext->rows;
    if (m < 1) return IPB_OK;
    if (!p->pM) IPB_FAIL(ctx, IPB_ERR_INVALID, "the banded solve needs the own rows of M (residual check)");
    int ov = overlap > 0 ? overlap : 48;
    ov = (ov + 3) & ~3;
    if (ghost_below > ov || ghost_above > ov || ext_off < ghost_below || ext_off + m + ghost_above > m_ext)
        IPB_FAIL(ctx, IPB_ERR_INVALID, "ipb_proj_set_band_parts: margins do not cover the ghost rows");
    if (ctx->size > 1 && (ghost_below > ctx->n_lo[1] || ghost_above > ctx->n_hi[1]))
        IPB_FAIL(ctx, IPB_ERR_INVALID, "m-space ghost zones are smaller than the sweep's halo");
    int *d_bw = nullptr;
    CHECK(dalloc(ctx, &d_bw, 1));
    CU(ctx, cudaMemsetAsync(d_bw, 0, sizeof(int), ctx->stream));
    k_band_width<<<cdiv(m_ext, 256), 256, 0, ctx->stream>>>(m_ext, M_ext->rowptr, M_ext->col, d_bw);
    int bw = 0;
    CU(ctx, cudaMemcpyAsync(&bw, d_bw, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_bw);
    p->band_bw = bw;
    if (bw < 1 || bw >= HB) return IPB_OK;
    int B = chunk_rows;
    if (B <= 0) {
        B = band_chunk_rows(ctx, m);
    }
    B = (B + 3) & ~3;
    const int64_t nchunk = (m + (int64_t)B - 1) / B;
    const int64_t ngroup = (nchunk + 31) / 32;
    p->band_B = B; p->band_ov = ov; p->band_steps = B + 2 * ov; p->band_ngroup = (int)ngroup;
    CHECK(dalloc(ctx, &p->band_Lrow, (size_t)m_ext * HB));
    CHECK(dalloc(ctx, &p->band_L, (size_t)ngroup * p->band_steps * HB * 32));
    CHECK(dalloc(ctx, &p->band_y, (size_t)ngroup * (B + ov) * 32));
    CHECK(dalloc(ctx, &p->band_ok, 1));
    const int one = 1;
    CU(ctx, cudaMemcpyAsync(p->band_ok, &one, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    const int FB = 128;       // rows per factor thread
    k_band_factor<<<cdiv(cdiv(m_ext, FB), 64), 64, 0, ctx->stream>>>(
        m_ext, M_ext->rowptr, M_ext->col, M_ext->val, FB, 48, p->band_Lrow, p->band_ok);
    const long long total = (long long)ngroup * p->band_steps * HB * 32;
    k_band_interleave<<<cdiv(total, 256), 256, 0, ctx->stream>>>(m_ext, B, ov, p->band_steps, total,
                                                                 (long long)ext_off, p->band_Lrow, p->band_L);
    ctx->launches += 3;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    p->band_in_lo = -ghost_below; p->band_in_hi = m + ghost_above;
    p->band = true;
    return IPB_OK;
}

extern "C" int ipb_proj_band_info(ipb_ctx *ctx, const ipb_proj *p, int32_t *out) {
    if (!ctx || !p || !out) return IPB_ERR_INVALID;
    int ok = 0;
    if (p->band) {
        CU(ctx, cudaMemcpyAsync(&ok, p->band_ok, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    out[0] = p->band ? 1 : 0; out[1] = p->band_bw; out[2] = p->band_B; out[3] = p->band_ov;
    out[4] = p->band_ngroup; out[5] = ok;
    return IPB_OK;
}

extern "C" int ipb_proj_refresh(ipb_ctx *ctx, ipb_proj *p) {
    if (!ctx || !p) return IPB_ERR_INVALID;
    if (p->parts) return IPB_OK;       // the host supplies refreshed parts
    ipb_csr *A = p->A;
    double s = 0.0;
    if (A->rows > 0) {
        k_row_norms<<<cdiv(A->rows, 256), 256, 0, ctx->stream>>>(A->rows, A->rowptr, A->val,
                                                                  p->dnorm2, p->dscale);
        k_scale_rows<<<cdiv(A->rows, 256), 256, 0, ctx->stream>>>(A->rows, A->rowptr, A->val,
                                                                   p->dscale, p->nval);
        ctx->launches += 2;
        if (A->nnz) {
            k_gather<<<cdiv(A->nnz, 256), 256, 0, ctx->stream>>>(A->nnz, A->perm, p->nval, p->nt_val);
            ctx->launches += 1;
        }
        if (p->has_M) {
            k_aat_values<<<cdiv(A->rows, 128), 128, 0, ctx->stream>>>(
                A->rows, A->rowptr, A->col, p->nval, p->m_rowptr, p->m_col, p->m_val);
            ctx->launches += 1;
            CHECK(refresh_band(ctx, p));
        }
        CHECK(device_sum(ctx, p->dnorm2, A->rows, &s));
    }
    p->fro = std::sqrt(s);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IPB_OK;
}

static int proj_vectors(ipb_ctx *ctx, ipb_proj *p) {
    CHECK(valloc(ctx, &p->v, p->m)); CHECK(valloc(ctx, &p->r, p->m));
    CHECK(valloc(ctx, &p->w, p->m)); CHECK(valloc(ctx, &p->p, p->m));
    CHECK(valloc(ctx, &p->s, p->m)); CHECK(valloc(ctx, &p->rhs, p->m));
    CHECK(valloc(ctx, &p->q, p->m));
    CHECK(valloc(ctx, &p->t, p->n));
    return IPB_OK;
}

extern "C" int ipb_proj_create(ipb_ctx *ctx, ipb_csr *A, double orth_tol, int max_refin,
                               double inner_rtol, int inner_maxit, ipb_proj **out) {
    if (!ctx || !A || !out) return IPB_ERR_INVALID;
    if (!A->has_t) IPB_FAIL(ctx, IPB_ERR_INVALID, "projector needs A created with_transpose");
    ipb_proj *p = new ipb_proj();
    p->A = A;
    p->m = A->rows; p->n = A->cols;
    if (ctx->size > 1)
        IPB_FAIL(ctx, IPB_ERR_INVALID, "row-partitioned runs create the projector with ipb_proj_create_parts");
    p->orth_tol = orth_tol; p->max_refin = max_refin;
    p->rtol = inner_rtol > 0 ? inner_rtol : 1e-13;
    p->maxit = inner_maxit > 0 ? inner_maxit : 1000;
    size_t m = A->rows, nnz = A->nnz;
    CHECK(dalloc(ctx, &p->nval, nnz + PAD));
    CHECK(dalloc(ctx, &p->nt_val, nnz + PAD));
    CHECK(dalloc(ctx, &p->dscale, m));
    CHECK(dalloc(ctx, &p->dnorm2, m));
    CHECK(proj_vectors(ctx, p));
    const char *env = getenv("IPB_EXPLICIT_AAT");      // 0 auto, 1 never, 2 always
    if (env) p->aat_mode = atoi(env);
    CHECK(build_aat_pattern(ctx, p));
    const char *benv = getenv("IPB_BAND");             // 0: never use the banded factor
    if (benv && atoi(benv) == 0) p->band_mode = 1;
    if ((benv = getenv("IPB_BAND_B"))) p->band_B_req = atoi(benv);
    if ((benv = getenv("IPB_BAND_OV"))) p->band_ov_req = atoi(benv);
    CHECK(build_band(ctx, p));
    CHECK(ipb_proj_refresh(ctx, p));
    *out = p;
    return IPB_OK;
}

extern "C" int ipb_proj_set_aat_mode(ipb_ctx *ctx, ipb_proj *p, int mode) {
    if (!ctx || !p || mode < 0 || mode > 2) return IPB_ERR_INVALID;
    if (p->parts) IPB_FAIL(ctx, IPB_ERR_INVALID, "a projector built from parts holds M iff it was given");
    p->aat_mode = mode;
    CHECK(build_aat_pattern(ctx, p));
    CHECK(build_band(ctx, p));
    return ipb_proj_refresh(ctx, p);
}

extern "C" int64_t ipb_proj_aat_nnz(const ipb_proj *p) {
    if (p && p->parts) return p->pM ? p->pM->nnz : -1;
    return (p && p->has_M) ? p->m_nnz : -1;
}

extern "C" int ipb_proj_destroy(ipb_ctx *ctx, ipb_proj *p) {
    if (!p) return IPB_OK;
    cudaFree(p->nval); cudaFree(p->nt_val); cudaFree(p->dscale); cudaFree(p->dnorm2);
    vfree(ctx, p->v); vfree(ctx, p->r); vfree(ctx, p->w); vfree(ctx, p->p); vfree(ctx, p->s);
    vfree(ctx, p->rhs); vfree(ctx, p->t); vfree(ctx, p->q);
    free_aat(p);
    free_band(p);
    delete p;
    return IPB_OK;
}

static ProjView proj_view(const ipb_proj *p) {
    ProjView v;
    memset(&v, 0, sizeof v);
    v.m = (int)p->m; v.n = (int)p->n;
    v.dscale = p->dscale; v.dnorm2 = p->dnorm2;
    v.fro = p->fro; v.orth_tol = p->orth_tol; v.max_refin = p->max_refin;
    v.rtol = p->rtol; v.maxit = p->maxit;
    v.v = p->v; v.r = p->r; v.w = p->w; v.p = p->p; v.s = p->s; v.rhs = p->rhs; v.t = p->t;
    v.q = p->q;
    if (p->band) {
        v.band.enabled = 1; v.band.B = p->band_B; v.band.ov = p->band_ov;
        v.band.steps = p->band_steps; v.band.ngroup = p->band_ngroup;
        v.band.L = p->band_L; v.band.y = p->band_y; v.band.ok = p->band_ok;
        v.band.in_lo = p->band_in_lo; v.band.in_hi = p->band_in_hi;
    }
    if (p->parts) {
        v.A = csr_view(p->pA_norm);
        v.At = csr_view(p->pAt_norm);
        v.has_M = p->pM ? 1 : 0;
        if (p->pM) v.M = csr_view(p->pM);
        v.Araw = p->pA_raw->val;
        return v;
    }
    v.A = csr_view(p->A, p->nval);
    v.At = csr_view_t(p->A, p->nt_val);
    v.has_M = p->has_M ? 1 : 0;
    if (p->has_M) {
        v.M.space = 1; v.M.xform_ok = 0;
        v.M.rows = v.m; v.M.cols = v.m; v.M.nblk = p->m_nblk; v.M.desc = p->m_desc;
        v.M.rowptr = p->m_rowptr; v.M.col = p->m_col; v.M.val = p->m_val;
    }
    v.Araw = p->A->val;
    v.dscale = p->dscale; v.dnorm2 = p->dnorm2;
    v.fro = p->fro; v.orth_tol = p->orth_tol; v.max_refin = p->max_refin;
    v.rtol = p->rtol; v.maxit = p->maxit;
    v.v = p->v; v.r = p->r; v.w = p->w; v.p = p->p; v.s = p->s; v.rhs = p->rhs; v.t = p->t;
    return v;
}

static int proj_grid(const ipb_ctx *ctx, const ipb_proj *p) {
    return grid_for(ctx, std::max<int64_t>(p->n, 1) * 4);
}

static int run_proj(ipb_ctx *ctx, ipb_proj *p, int op, const double *din, double *dout,
                    int fixed_iters, ipb_proj_info *info) {
    ProjArgs a;
    memset(&a, 0, sizeof a);
    a.P = proj_view(p); a.S = sync_view(ctx); a.op = op; a.in = din; a.out = dout;
    a.fixed_iters = fixed_iters;
    a.result = (ProjArgs::Result *)ctx->d_result;
    float ms = 0.f;
    CHECK(launch_coop(ctx, k_projector, a, proj_grid(ctx, p), &ms));
    ProjArgs::Result *h = (ProjArgs::Result *)ctx->h_result;
    CU(ctx, cudaMemcpyAsync(h, ctx->d_result, sizeof(*h), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (info) {
        info->inner_iters = h->inner_iters; info->n_solves = h->n_solves;
        info->n_refinements = h->n_refinements; info->status = h->status;
        info->n_fallbacks = h->n_fallbacks; info->orth = h->orth; info->kernel_ms = ms;
    }
    if (h->status != 0)
        IPB_FAIL(ctx, h->status, "no solver for (A A') v = rhs converged (status %d)", h->status);
    return IPB_OK;
}

extern "C" int ipb_proj_null_space_dev(ipb_ctx *ctx, ipb_proj *p, const ipb_vec *x, ipb_vec *z,
                                       ipb_proj_info *info) {
    if (!ctx || !p || !x || !z || x->len != p->n || z->len != p->n) return IPB_ERR_INVALID;
    return run_proj(ctx, p, 0, x->d, z->d, 0, info);
}
extern "C" int ipb_proj_least_squares_dev(ipb_ctx *ctx, ipb_proj *p, const ipb_vec *x, ipb_vec *v,
                                          ipb_proj_info *info) {
    if (!ctx || !p || !x || !v || x->len != p->n || v->len != p->m) return IPB_ERR_INVALID;
    return run_proj(ctx, p, 1, x->d, v->d, 0, info);
}
extern "C" int ipb_proj_row_space_dev(ipb_ctx *ctx, ipb_proj *p, const ipb_vec *b, ipb_vec *y,
                                      ipb_proj_info *info) {
    if (!ctx || !p || !b || !y || b->len != p->m || y->len != p->n) return IPB_ERR_INVALID;
    return run_proj(ctx, p, 2, b->d, y->d, 0, info);
}

extern "C" int ipb_proj_null_space(ipb_ctx *ctx, ipb_proj *p, const double *x, double *z,
                                   ipb_proj_info *info) {
    if (!ctx || !p) return IPB_ERR_INVALID;
    int64_t n = p->n;
    double *din, *dout;
    CHECK(host_in_out(ctx, n, n, x, &din, &dout));
    CHECK(run_proj(ctx, p, 0, din, dout, 0, info));
    return host_out(ctx, n, dout, z);
}
extern "C" int ipb_proj_least_squares(ipb_ctx *ctx, ipb_proj *p, const double *x, double *v,
                                      ipb_proj_info *info) {
    if (!ctx || !p) return IPB_ERR_INVALID;
    int64_t n = p->n, m = p->m;
    double *din, *dout;
    CHECK(host_in_out(ctx, n, m, x, &din, &dout));
    CHECK(run_proj(ctx, p, 1, din, dout, 0, info));
    return host_out(ctx, m, dout, v);
}
extern "C" int ipb_proj_row_space(ipb_ctx *ctx, ipb_proj *p, const double *b, double *y,
                                  ipb_proj_info *info) {
    if (!ctx || !p) return IPB_ERR_INVALID;
    int64_t n = p->n, m = p->m;
    double *din, *dout;
    CHECK(host_in_out(ctx, m, n, b, &din, &dout));
    CHECK(run_proj(ctx, p, 2, din, dout, 0, info));
    return host_out(ctx, n, dout, y);
}
extern "C" int ipb_orthogonality(ipb_ctx *ctx, ipb_proj *p, const double *g, double *out) {
    if (!ctx || !p || !out) return IPB_ERR_INVALID;
    int64_t n = p->n;
    double *din, *dout;
    CHECK(host_in_out(ctx, n, 0, g, &din, &dout));
    ipb_proj_info info;
    CHECK(run_proj(ctx, p, 3, din, nullptr, 0, &info));
    *out = info.orth;
    return IPB_OK;
}

extern "C" int ipb_bench_inner_iterations(ipb_ctx *ctx, ipb_proj *p, const ipb_vec *x, int iters,
                                          double *ms) {
    if (!ctx || !p || !x || iters <= 0 || x->len != p->n) return IPB_ERR_INVALID;
    ipb_proj_info info;
    CHECK(run_proj(ctx, p, 4, x->d, nullptr, iters, &info));
    if (ms) *ms = info.kernel_ms;
    return IPB_OK;
}

extern "C" int ipb_bench_phase(ipb_ctx *ctx, ipb_proj *p, int phase, int reps, double *ms) {
    if (!ctx || !p || reps <= 0 || phase < 0 || phase > 9) return IPB_ERR_INVALID;
    ipb_proj_info info;
    CHECK(run_proj(ctx, p, 10 + phase, p->rhs, nullptr, reps, &info));
    if (ms) *ms = info.kernel_ms;
    return IPB_OK;
}

// workspace vectors of a row-partitioned context (allocated once, same order on every rank)
static int dist_workspace(ipb_ctx *ctx) {
    if (ctx->dws_m) return IPB_OK;
    if (ctx->nm_global < 0)
        IPB_FAIL(ctx, IPB_ERR_INVALID, "call ipb_comm_init before solving on a row-partitioned context");
    for (int k = 0; k < 7; ++k) CHECK(valloc(ctx, &ctx->dws_n[k], ctx->own[0]));
    CHECK(valloc(ctx, &ctx->dws_m, ctx->own[1]));
    return IPB_OK;
}

// ------------------------------------------------------------------ projected CG
static int run_pcg(ipb_ctx *ctx, const ipb_hess *H, ipb_proj *P, const double *dc,
                   const double *db, double radius, const double *dlb, const double *dub,
                   double tol, int64_t max_iter, int64_t max_inf, double *dx_out,
                   double *d_allvecs, int64_t allvecs_cap, ipb_pcg_info *info) {
    const int64_t n = P->n, m = P->m;
    if (H->n != n) IPB_FAIL(ctx, IPB_ERR_INVALID, "H is %lld x %lld but A has %lld columns",
                            (long long)H->n, (long long)H->n, (long long)n);
    const size_t ns = vstride(n);
    if (ctx->size == 1) CHECK(ensure_ws(ctx, 7 * ns));
    else CHECK(dist_workspace(ctx));
    PcgArgs a;
    memset(&a, 0, sizeof a);
    a.H = hess_view(H); a.P = proj_view(P); a.S = sync_view(ctx);
    a.n = (int)n; a.m = (int)m;
    a.c = dc; a.b = db; a.lb = dlb; a.ub = dub;
    a.has_box = (dlb != nullptr && dub != nullptr);
    a.radius = radius; a.tol = tol;
    a.max_iter = max_iter; a.max_infeasible = max_inf;
    double *w = ctx->ws;
    if (ctx->size == 1) {
        a.x = w; a.xn = w + ns; a.r = w + 2 * ns; a.p = w + 3 * ns; a.Hp = w + 4 * ns;
        a.xfeas = w + 5 * ns; a.tmp = w + 6 * ns;
    } else {
        double **d = ctx->dws_n;
        a.x = d[0]; a.xn = d[1]; a.r = d[2]; a.p = d[3]; a.Hp = d[4]; a.xfeas = d[5]; a.tmp = d[6];
    }
    a.nm_global = ctx->size > 1 ? ctx->nm_global : (long long)n - m;
    a.allvecs = d_allvecs; a.allvecs_cap = allvecs_cap;
    a.result = (PcgArgs::Result *)ctx->d_result;
    float ms = 0.f;
    int64_t work = std::max<int64_t>(n * 4, H->D ? H->nx * H->nx / 8 : 0);
    CHECK(launch_coop(ctx, k_projected_cg, a, grid_for(ctx, work), &ms));
    PcgArgs::Result *h = (PcgArgs::Result *)ctx->h_result;
    CU(ctx, cudaMemcpyAsync(h, ctx->d_result, sizeof(*h), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    const double *src = h->x_slot == 0 ? a.x : (h->x_slot == 1 ? a.xn : a.xfeas);
    if (n) CU(ctx, cudaMemcpyAsync(dx_out, src, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    if (info) {
        info->niter = h->niter; info->stop_cond = h->stop_cond;
        info->hits_boundary = h->hits_boundary; info->status = h->status;
        info->n_allvecs = h->n_allvecs; info->inner_iters = h->inner_iters;
        info->n_solves = h->n_solves; info->n_refinements = h->n_refinements;
        info->n_spmv_h = h->n_spmv_h; info->max_orth = h->max_orth;
        info->tol_used = h->tol_used; info->kernel_ms = ms;
        info->n_sweeps = h->n_sweeps; info->n_fallbacks = h->n_fallbacks;
        info->reserved = h->fused_hp;          // 1: p = -g + beta p was formed inside H p
    }
    if (h->status == IPB_ERR_TR_INFEASIBLE)
        IPB_FAIL(ctx, IPB_ERR_TR_INFEASIBLE, "Trust region problem does not have a solution.");
    if (h->status == IPB_ERR_NEG_CURVATURE)
        IPB_FAIL(ctx, IPB_ERR_NEG_CURVATURE, "Negative curvature not allowed for unrestrited problems.");
    if (h->status != 0)
        IPB_FAIL(ctx, h->status, "projected CG: inner solve failed (status %d)", h->status);
    return IPB_OK;
}

extern "C" int ipb_projected_cg_dev(ipb_ctx *ctx, const ipb_hess *H, ipb_proj *P,
                                    const ipb_vec *c, const ipb_vec *b, double trust_radius,
                                    const ipb_vec *lb, const ipb_vec *ub, double tol,
                                    int64_t max_iter, int64_t max_infeasible_iter,
                                    ipb_vec *x, ipb_pcg_info *info) {
    if (!ctx || !H || !P || !c || !b || !x) return IPB_ERR_INVALID;
    const int64_t n = P->n, m = P->m;
    if (c->len != n || b->len != m || x->len != n || (lb && lb->len != n) || (ub && ub->len != n))
        IPB_FAIL(ctx, IPB_ERR_INVALID, "projected_cg: vector lengths do not match A");
    int s = run_pcg(ctx, H, P, c->d, b->d, trust_radius, lb ? lb->d : nullptr,
                    ub ? ub->d : nullptr, tol, max_iter, max_infeasible_iter, x->d, nullptr, 0, info);
    cudaStreamSynchronize(ctx->stream);
    return s;
}

extern "C" int ipb_projected_cg(ipb_ctx *ctx, const ipb_hess *H, ipb_proj *P, const double *c,
                                const double *b, double trust_radius, const double *lb,
                                const double *ub, double tol, int64_t max_iter,
                                int64_t max_infeasible_iter, double *x, double *allvecs,
                                int64_t allvecs_cap, ipb_pcg_info *info) {
    if (!ctx || !H || !P || !c || !b || !x) return IPB_ERR_INVALID;
    SINGLE_GPU_ONLY(ctx);
    const int64_t n = P->n, m = P->m;
    const bool box = lb && ub;
    if (!allvecs) allvecs_cap = 0;
    const size_t ns = vstride(n), ms = vstride(m);
    size_t need = ns * (2 + (box ? 2 : 0)) + ms + (size_t)allvecs_cap * n;
    CHECK(ensure_stage(ctx, need));
    double *dc = ctx->stage, *db = dc + ns, *dx = db + ms, *dlb = dx + ns, *dub = dlb + (box ? ns : 0);
    double *dall = allvecs_cap ? dub + (box ? ns : 0) : nullptr;
    if (n) CU(ctx, cudaMemcpyAsync(dc, c, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (m) CU(ctx, cudaMemcpyAsync(db, b, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (box && n) {
        CU(ctx, cudaMemcpyAsync(dlb, lb, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(dub, ub, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    ipb_pcg_info local;
    if (!info) info = &local;
    int s = run_pcg(ctx, H, P, dc, db, trust_radius, box ? dlb : nullptr, box ? dub : nullptr, tol,
                    max_iter, max_infeasible_iter, dx, dall, allvecs_cap, info);
    if (s != IPB_OK) { cudaStreamSynchronize(ctx->stream); return s; }
    if (n) CU(ctx, cudaMemcpyAsync(x, dx, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    int64_t got = std::min<int64_t>(info->n_allvecs, allvecs_cap);
    if (got > 0 && n)
        CU(ctx, cudaMemcpyAsync(allvecs, dall, (size_t)got * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IPB_OK;
}

// ------------------------------------------------------------------ dogleg
// modified_dogleg on device vectors; the result is left in ctx->ws (returned pointer).
static int run_dogleg(ipb_ctx *ctx, ipb_proj *P, const double *db, double trust_radius,
                      const double *dlb, const double *dub, const double **d_out,
                      ipb_proj_info *info) {
    const int64_t n = P->n, m = P->m;
    const bool box = dlb && dub;
    const size_t ns = vstride(n), mstr = vstride(m);
    if (ctx->size == 1) CHECK(ensure_ws(ctx, 7 * ns + mstr));
    else CHECK(dist_workspace(ctx));        // arena vectors: products exchange their ghosts
    DoglegArgs a;
    memset(&a, 0, sizeof a);
    a.P = proj_view(P); a.S = sync_view(ctx);
    if (P->parts) { a.Araw = csr_view(P->pA_raw); a.Atraw = csr_view(P->pAt_raw); }
    else { a.Araw = csr_view(P->A); a.Atraw = csr_view_t(P->A); }
    a.n = (int)n; a.m = (int)m; a.b = db; a.lb = box ? dlb : nullptr; a.ub = box ? dub : nullptr;
    a.has_box = box; a.radius = trust_radius;
    if (ctx->size == 1) {
        double *w = ctx->ws;
        a.newton = w; a.cauchy = w + ns; a.g = w + 2 * ns; a.x1 = w + 3 * ns; a.x2 = w + 4 * ns;
        a.out = w + 5 * ns; a.mtmp = w + 7 * ns;
    } else {
        double **d = ctx->dws_n;
        a.newton = d[0]; a.cauchy = d[1]; a.g = d[2]; a.x1 = d[3]; a.x2 = d[4]; a.out = d[5];
        a.mtmp = ctx->dws_m;
    }
    a.result = (ProjArgs::Result *)ctx->d_result;
    float ms = 0.f;
    CHECK(launch_coop(ctx, k_dogleg, a, proj_grid(ctx, P), &ms));
    ProjArgs::Result *h = (ProjArgs::Result *)ctx->h_result;
    CU(ctx, cudaMemcpyAsync(h, ctx->d_result, sizeof(*h), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (info) {
        info->inner_iters = h->inner_iters; info->n_solves = h->n_solves;
        info->n_refinements = h->n_refinements; info->status = h->status;
        info->n_fallbacks = h->n_fallbacks; info->orth = 0.0; info->kernel_ms = ms;
    }
    if (h->status != 0)
        IPB_FAIL(ctx, h->status, "modified dogleg: inner solve failed (status %d)", h->status);
    *d_out = a.out;
    return IPB_OK;
}

extern "C" int ipb_modified_dogleg(ipb_ctx *ctx, ipb_proj *P, const double *b, double trust_radius,
                                   const double *lb, const double *ub, double *x,
                                   ipb_proj_info *info) {
    if (!ctx || !P || !b || !x) return IPB_ERR_INVALID;
    SINGLE_GPU_ONLY(ctx);
    const int64_t n = P->n, m = P->m;
    const bool box = lb && ub;
    const size_t ns = vstride(n), mstr = vstride(m);
    CHECK(ensure_stage(ctx, mstr + (box ? 2 * ns : 0)));
    double *db = ctx->stage, *dlb = db + mstr, *dub = dlb + (box ? ns : 0);
    if (m) CU(ctx, cudaMemcpyAsync(db, b, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (box && n) {
        CU(ctx, cudaMemcpyAsync(dlb, lb, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(dub, ub, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    const double *dout = nullptr;
    CHECK(run_dogleg(ctx, P, db, trust_radius, box ? dlb : nullptr, box ? dub : nullptr, &dout, info));
    if (n) CU(ctx, cudaMemcpyAsync(x, dout, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IPB_OK;
}

// ------------------------------------------------------------------ device-resident SQP step
// One outer iteration of the Byrd-Omojokun SQP between two user callbacks
// (reference equality_constrained_sqp.py:113-140) with its vectors resident in
// HBM, in two calls because the host decides the tangential radius:
//   normal:      dn = modified_dogleg(A, Y, b, zeta*radius, box*lb, box*ub)      -> dn
//   (host)       radius_t = sqrt(radius^2 - norm(dn)^2)        (numpy, as the reference)
//   tangential:  c_t = H dn + c;  dt = projected_cg(H, c_t, Z, Y, 0, radius_t, lb - dn, ub - dn)
//                d = dn + dt;  H d;  A d                                         -> d, H d, A d
// c, b and finite bounds go up only when they changed; nothing the device
// already holds (dn, c_t, the shifted bounds, d) is uploaded again.  All SCALARS
// of the iteration (norms, the quadratic model) are left to the host's numpy on
// the downloaded vectors: runs that end by xtol are decided by round-off in
// exactly these scalars, and a device reduction order would change their tail.
static int glue(ipb_ctx *ctx, int mode, int64_t len, const double *a, const double *b,
                const double *c, double *o0, double *o1, double s0, int slot) {
    GlueArgs g;
    g.mode = mode; g.len = (int)len; g.a = a; g.b = b; g.c = c; g.o0 = o0; g.o1 = o1; g.s0 = s0;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(1024, (len + 255) / 256));
    g.partials = ctx->glue_partials; g.counter = ctx->glue_counter;
    g.sums = ctx->glue_sums + 3 * slot;
    k_glue<<<grid, 256, 0, ctx->stream>>>(g);
    ctx->launches += 1;
    return IPB_OK;
}

enum { SV_C = 0, SV_LB, SV_UB, SV_LBN, SV_UBN, SV_LBT, SV_UBT, SV_DN, SV_CT, SV_DT, SV_D, SV_HD,
       SV_ATV, SV_NCOUNT };
enum { SM_B = 0, SM_LIN, SM_ZERO, SM_V, SM_MCOUNT };

static int ensure_step(ipb_ctx *ctx, int64_t n, int64_t m) {
    if (!ctx->glue_partials) {
        CHECK(dalloc(ctx, &ctx->glue_partials, 3 * 1024));
        CHECK(dalloc(ctx, &ctx->glue_sums, 3 * 8));
        CHECK(dalloc(ctx, &ctx->glue_counter, 1));
        CU(ctx, cudaMemset(ctx->glue_counter, 0, sizeof(unsigned)));
    }
    if (ctx->step_n == n && ctx->step_m == m) return IPB_OK;
    static_assert(SV_NCOUNT <= 16 && SM_MCOUNT <= 8, "step vector tables");
    if (ctx->size > 1) {
        // row-partitioned: n, m are the rank's own lengths; every vector comes from the
        // symmetric arena (same order on every rank) so that products find its ghost zones
        if (ctx->step_n >= 0)
            IPB_FAIL(ctx, IPB_ERR_INVALID, "a row-partitioned context serves one problem");
        if (n != ctx->own[0] || m != ctx->own[1])
            IPB_FAIL(ctx, IPB_ERR_INVALID, "SQP step: the projector does not belong to this partition");
        for (int k = 0; k < SV_NCOUNT; ++k) CHECK(valloc(ctx, &ctx->step_nv[k], n));
        for (int k = 0; k < SM_MCOUNT; ++k) CHECK(valloc(ctx, &ctx->step_mv[k], m));
    } else {
        const size_t need = SV_NCOUNT * vstride(n) + SM_MCOUNT * vstride(m);
        if (ctx->step_len < need) {
            cudaFree(ctx->step);
            ctx->step = nullptr; ctx->step_len = 0;
            CHECK(dalloc(ctx, &ctx->step, need));
            ctx->step_len = need;
        }
        for (int k = 0; k < SV_NCOUNT; ++k) ctx->step_nv[k] = ctx->step + (size_t)k * vstride(n);
        for (int k = 0; k < SM_MCOUNT; ++k)
            ctx->step_mv[k] = ctx->step + SV_NCOUNT * vstride(n) + (size_t)k * vstride(m);
    }
    ctx->step_n = n; ctx->step_m = m;
    ctx->step_has_c = ctx->step_has_b = ctx->step_has_box = false;
    if (m) CU(ctx, cudaMemsetAsync(ctx->step_mv[SM_ZERO], 0, (size_t)m * 8, ctx->stream));
    return IPB_OK;
}
static inline double *step_n(ipb_ctx *ctx, int k) { return ctx->step_nv[k]; }
static inline double *step_m(ipb_ctx *ctx, int k) { return ctx->step_mv[k]; }

extern "C" int ipb_sqp_normal_step(ipb_ctx *ctx, ipb_proj *P, const double *c, const double *b,
                                   double normal_radius, const double *lb, const double *ub,
                                   int bounds_changed, double box_factor, double *dn,
                                   ipb_proj_info *info) {
    if (!ctx || !P || !dn) return IPB_ERR_INVALID;
    const int64_t n = P->n, m = P->m;
    CHECK(ensure_step(ctx, n, m));
    const bool box = lb && ub;
    double *dc = step_n(ctx, SV_C), *db = step_m(ctx, SM_B);
    if (c) {
        if (n) CU(ctx, cudaMemcpyAsync(dc, c, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        ctx->step_has_c = true;
    }
    if (b) {
        if (m) CU(ctx, cudaMemcpyAsync(db, b, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
        ctx->step_has_b = true;
    }
    if (!ctx->step_has_c || !ctx->step_has_b)
        IPB_FAIL(ctx, IPB_ERR_INVALID, "ipb_sqp_normal_step: c / b were never uploaded for this problem");
    if (box && (bounds_changed || !ctx->step_has_box)) {
        CU(ctx, cudaMemcpyAsync(step_n(ctx, SV_LB), lb, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(step_n(ctx, SV_UB), ub, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        CHECK(glue(ctx, 5, n, step_n(ctx, SV_LB), step_n(ctx, SV_UB), nullptr, step_n(ctx, SV_LBN),
                   step_n(ctx, SV_UBN), box_factor, 7));
    }
    ctx->step_has_box = box;
    const double *dn_ws = nullptr;
    CHECK(run_dogleg(ctx, P, db, normal_radius, box ? step_n(ctx, SV_LBN) : nullptr,
                     box ? step_n(ctx, SV_UBN) : nullptr, &dn_ws, info));
    double *ddn = step_n(ctx, SV_DN);
    if (n) {
        CU(ctx, cudaMemcpyAsync(ddn, dn_ws, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(dn, ddn, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IPB_OK;
}

extern "C" int ipb_sqp_tangential_step(ipb_ctx *ctx, const ipb_hess *H, ipb_proj *P,
                                       double tangential_radius, double *d, double *Hd,
                                       double *Ad, ipb_pcg_info *info) {
    if (!ctx || !H || !P || !d || !Hd || !info) return IPB_ERR_INVALID;
    const int64_t n = P->n, m = P->m;
    if (H->n != n) IPB_FAIL(ctx, IPB_ERR_INVALID, "H and A do not match");
    if (ctx->step_n != n || ctx->step_m != m || !ctx->step_has_c)
        IPB_FAIL(ctx, IPB_ERR_INVALID, "ipb_sqp_tangential_step needs ipb_sqp_normal_step first");
    const bool box = ctx->step_has_box;
    double *dc = step_n(ctx, SV_C), *dn = step_n(ctx, SV_DN);
    // c_t = H dn + c  (:125)
    CHECK(run_hess(ctx, H, dn, step_n(ctx, SV_HD)));
    CHECK(glue(ctx, 0, n, step_n(ctx, SV_HD), dc, dn, step_n(ctx, SV_CT), nullptr, 0.0, 0));
    if (box)                                                   // lb - dn, ub - dn  (:128-129)
        CHECK(glue(ctx, 1, n, step_n(ctx, SV_LB), step_n(ctx, SV_UB), dn, step_n(ctx, SV_LBT),
                   step_n(ctx, SV_UBT), 0.0, 7));
    int st = run_pcg(ctx, H, P, step_n(ctx, SV_CT), step_m(ctx, SM_ZERO), tangential_radius,
                     box ? step_n(ctx, SV_LBT) : nullptr, box ? step_n(ctx, SV_UBT) : nullptr, -1.0, -1,
                     -1, step_n(ctx, SV_DT), nullptr, 0, info);
    if (st != IPB_OK) { cudaStreamSynchronize(ctx->stream); return st; }
    // d = dn + dt, H d, A d  (:135-140)
    CHECK(glue(ctx, 2, n, dn, step_n(ctx, SV_DT), dc, step_n(ctx, SV_D), nullptr, 0.0, 1));
    if (n) CU(ctx, cudaMemcpyAsync(d, step_n(ctx, SV_D), (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CHECK(run_hess(ctx, H, step_n(ctx, SV_D), step_n(ctx, SV_HD)));
    if (n) CU(ctx, cudaMemcpyAsync(Hd, step_n(ctx, SV_HD), (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (m > 0 && Ad) {
        CHECK(run_matvec(ctx, P->parts ? csr_view(P->pA_raw) : csr_view(P->A), step_n(ctx, SV_D),
                         step_m(ctx, SM_LIN)));
        CU(ctx, cudaMemcpyAsync(Ad, step_m(ctx, SM_LIN), (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return IPB_OK;
}

// v = -LS c and ||c + A'v||_inf after an accepted step (reference :227, :238);
// c becomes the resident gradient of the following ipb_sqp_step calls.
extern "C" int ipb_sqp_multipliers(ipb_ctx *ctx, ipb_proj *P, const double *c, double *v,
                                   double *optimality, ipb_proj_info *info) {
    if (!ctx || !P || !c || !v || !optimality) return IPB_ERR_INVALID;
    const int64_t n = P->n, m = P->m;
    CHECK(ensure_step(ctx, n, m));
    double *dc = step_n(ctx, SV_C);
    if (n) CU(ctx, cudaMemcpyAsync(dc, c, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    ctx->step_has_c = true;
    double *dv = step_m(ctx, SM_V), *atv = step_n(ctx, SV_ATV);
    if (m > 0) {
        CHECK(run_proj(ctx, P, 1, dc, step_m(ctx, SM_LIN), 0, info));
        CHECK(glue(ctx, 6, m, step_m(ctx, SM_LIN), nullptr, nullptr, dv, nullptr, 0.0, 7));
        CHECK(run_matvec(ctx, P->parts ? csr_view(P->pAt_raw) : csr_view_t(P->A), dv, atv));
    } else {
        if (n) CU(ctx, cudaMemsetAsync(atv, 0, (size_t)n * 8, ctx->stream));
    }
    CHECK(glue(ctx, 7, n, dc, atv, nullptr, nullptr, nullptr, 0.0, 4));
    double opt = 0.0;
    CU(ctx, cudaMemcpyAsync(&opt, ctx->glue_sums + 12, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (m) CU(ctx, cudaMemcpyAsync(v, dv, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    *optimality = opt;
    return IPB_OK;
}

// ------------------------------------------------------------------ intersections
extern "C" int ipb_box_sphere_intersections(ipb_ctx *ctx, int64_t n, const double *z,
                                            const double *d, const double *lb, const double *ub,
                                            double trust_radius, int entire_line, double *ta,
                                            double *tb, int *intersect) {
    if (!ctx || !z || !d || !ta || !tb || !intersect || n < 0) return IPB_ERR_INVALID;
    SINGLE_GPU_ONLY(ctx);
    const bool box = lb && ub;
    CHECK(ensure_stage(ctx, (size_t)n * (box ? 4 : 2)));
    double *dz = ctx->stage, *dd = dz + n, *dlb = dd + n, *dub = dlb + (box ? n : 0);
    if (n) {
        CU(ctx, cudaMemcpyAsync(dz, z, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemcpyAsync(dd, d, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (box) {
            CU(ctx, cudaMemcpyAsync(dlb, lb, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
            CU(ctx, cudaMemcpyAsync(dub, ub, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    IntersectArgs a;
    memset(&a, 0, sizeof a);
    a.S = sync_view(ctx); a.n = (int)n; a.z = dz; a.d = dd;
    a.lb = box ? dlb : nullptr; a.ub = box ? dub : nullptr;
    a.has_box = box; a.line = entire_line; a.radius = trust_radius;
    a.result = (IntersectArgs::Result *)ctx->d_result;
    CHECK(launch_coop(ctx, k_intersect, a, grid_for(ctx, n), nullptr));
    IntersectArgs::Result *h = (IntersectArgs::Result *)ctx->h_result;
    CU(ctx, cudaMemcpyAsync(h, ctx->d_result, sizeof(*h), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    *ta = h->ta; *tb = h->tb; *intersect = h->hit;
    return IPB_OK;
}

// The same on device vectors (row-partitioned contexts: the rank's own blocks; the
// interval is reduced over the ranks inside the kernel).
extern "C" int ipb_box_sphere_intersections_dev(ipb_ctx *ctx, const ipb_vec *z, const ipb_vec *d,
                                                const ipb_vec *lb, const ipb_vec *ub,
                                                double trust_radius, int entire_line, double *ta,
                                                double *tb, int *intersect) {
    if (!ctx || !z || !d || !ta || !tb || !intersect || z->len != d->len) return IPB_ERR_INVALID;
    const bool box = lb && ub;
    if (box && (lb->len != z->len || ub->len != z->len)) return IPB_ERR_INVALID;
    IntersectArgs a;
    memset(&a, 0, sizeof a);
    a.S = sync_view(ctx); a.n = (int)z->len; a.z = z->d; a.d = d->d;
    a.lb = box ? lb->d : nullptr; a.ub = box ? ub->d : nullptr;
    a.has_box = box; a.line = entire_line; a.radius = trust_radius;
    a.result = (IntersectArgs::Result *)ctx->d_result;
    CHECK(launch_coop(ctx, k_intersect, a, grid_for(ctx, std::max<int64_t>(z->len, 1) * 4), nullptr));
    IntersectArgs::Result *h = (IntersectArgs::Result *)ctx->h_result;
    CU(ctx, cudaMemcpyAsync(h, ctx->d_result, sizeof(*h), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    *ta = h->ta; *tb = h->tb; *intersect = h->hit;
    return IPB_OK;
}

// ------------------------------------------------------------------ pinned host memory
extern "C" int ipb_host_alloc(ipb_ctx *ctx, int64_t bytes, void **out) {
    if (!ctx || !out || bytes < 0) return IPB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMallocHost(out, (size_t)std::max<int64_t>(bytes, 8)));
    return IPB_OK;
}
extern "C" int ipb_host_free(ipb_ctx *ctx, void *p) {
    (void)ctx;
    if (p) cudaFreeHost(p);
    return IPB_OK;
}

// ------------------------------------------------------------------ multi-GPU
extern "C" int ipb_csr_set_space(ipb_ctx *ctx, ipb_csr *m, int space) {
    if (!ctx || !m || space < 0 || space > 1) return IPB_ERR_INVALID;
    m->space = space;
    return IPB_OK;
}

extern "C" int ipb_comm_init(ipb_ctx *ctx, int rank, int size, const int64_t *own,
                             const int64_t *own_max, const int64_t *ghost_lo,
                             const int64_t *ghost_hi, const int64_t *my_ghost_lo,
                             const int64_t *my_ghost_hi, int64_t n_minus_m_global,
                             int64_t arena_bytes, void *ipc_handle_out) {
    if (!ctx || size < 1 || size > MAX_RANKS || rank < 0 || rank >= size || !own || !own_max ||
        !ghost_lo || !ghost_hi || !my_ghost_lo || !my_ghost_hi || !ipc_handle_out)
        return IPB_ERR_INVALID;
    if (ctx->arena) IPB_FAIL(ctx, IPB_ERR_INVALID, "ipb_comm_init called twice");
    CU(ctx, cudaSetDevice(ctx->device));
    for (int k = 0; k < 2; ++k) {
        if (ghost_lo[k] & 1) IPB_FAIL(ctx, IPB_ERR_INVALID, "ghost_lo must be even (16-byte alignment)");
        ctx->own[k] = own[k]; ctx->own_max[k] = own_max[k];
        ctx->glo[k] = ghost_lo[k]; ctx->ghi[k] = ghost_hi[k];
        if (my_ghost_lo[k] > ghost_lo[k] || my_ghost_hi[k] > ghost_hi[k])
            IPB_FAIL(ctx, IPB_ERR_INVALID, "a rank's ghost count exceeds the common capacity");
        ctx->n_lo[k] = my_ghost_lo[k]; ctx->n_hi[k] = my_ghost_hi[k];
    }
    ctx->rank = rank; ctx->size = size; ctx->nm_global = n_minus_m_global;
    // ghost mailboxes: 4 generations x capacity x 2 words, per space (same offsets on every rank)
    size_t off = VEC_OFF;
    for (int k = 0; k < 2; ++k) {
        ctx->gmbox_off[k] = off;
        off += ((size_t)4 * (size_t)(ghost_lo[k] + ghost_hi[k]) * 16 + 255) & ~(size_t)255;
    }
    ctx->arena_bytes = ((size_t)arena_bytes + off + 4095) & ~(size_t)4095;
    CU(ctx, cudaMalloc((void **)&ctx->arena, ctx->arena_bytes));
    CU(ctx, cudaMemset(ctx->arena, 0, ctx->arena_bytes));
    ctx->arena_used = off;
    cudaIpcMemHandle_t h;
    CU(ctx, cudaIpcGetMemHandle(&h, ctx->arena));
    memcpy(ipc_handle_out, &h, sizeof h);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (size == 1) ctx->peer_base[0] = ctx->arena;
    return IPB_OK;
}

extern "C" int ipb_comm_connect(ipb_ctx *ctx, const void *handles) {
    if (!ctx || !handles || !ctx->arena) return IPB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    for (int q = 0; q < ctx->size; ++q) {
        if (q == ctx->rank) { ctx->peer_base[q] = ctx->arena; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * (size_t)q, sizeof h);
        void *p = nullptr;
        CU(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_base[q] = (char *)p;
    }
    return IPB_OK;
}

extern "C" int ipb_comm_set_send_list(ipb_ctx *ctx, int space, int64_t count, const int32_t *entries) {
    if (!ctx || space < 0 || space > 1 || count < 0 || (count && !entries)) return IPB_ERR_INVALID;
    cudaFree(ctx->send_ent[space]);
    ctx->send_ent[space] = nullptr; ctx->send_cnt[space] = 0;
    if (count) {
        CHECK(dalloc(ctx, &ctx->send_ent[space], (size_t)count));
        CU(ctx, cudaMemcpy(ctx->send_ent[space], entries, (size_t)count * sizeof(int4), cudaMemcpyHostToDevice));
        ctx->send_cnt[space] = (int)count;
    }
    return IPB_OK;
}

extern "C" int ipb_proj_create_parts(ipb_ctx *ctx, ipb_csr *A_raw, ipb_csr *A_norm, ipb_csr *At_raw,
                                     ipb_csr *At_norm, ipb_csr *M, const double *dscale,
                                     double fro, double orth_tol, int max_refin,
                                     double inner_rtol, int inner_maxit, ipb_proj **out) {
    if (!ctx || !A_raw || !A_norm || !At_raw || !At_norm || !out || (A_raw->rows && !dscale))
        return IPB_ERR_INVALID;
    if (A_norm->rows != A_raw->rows || A_norm->nnz != A_raw->nnz || At_norm->rows != At_raw->rows ||
        (M && M->rows != A_raw->rows))
        IPB_FAIL(ctx, IPB_ERR_INVALID, "ipb_proj_create_parts: inconsistent parts");
    ipb_proj *p = new ipb_proj();
    p->parts = true;
    p->pA_raw = A_raw; p->pA_norm = A_norm; p->pAt_raw = At_raw; p->pAt_norm = At_norm; p->pM = M;
    A_raw->space = 0; A_norm->space = 0; At_raw->space = 1; At_norm->space = 1;
    if (M) M->space = 1;
    p->m = A_raw->rows; p->n = At_raw->rows;
    p->fro = fro; p->orth_tol = orth_tol; p->max_refin = max_refin;
    p->rtol = inner_rtol > 0 ? inner_rtol : 1e-13;
    p->maxit = inner_maxit > 0 ? inner_maxit : 1000;
    CHECK(dalloc(ctx, &p->dscale, (size_t)p->m));
    CHECK(dalloc(ctx, &p->dnorm2, (size_t)p->m));
    if (p->m) CU(ctx, cudaMemcpy(p->dscale, dscale, (size_t)p->m * 8, cudaMemcpyHostToDevice));
    CHECK(proj_vectors(ctx, p));
    *out = p;
    return IPB_OK;
}
