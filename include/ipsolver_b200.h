/*
 * ipsolver_b200 -- C ABI of the B200-native projected-CG / null-space-projection
 * hot path (fp64, sm_100a).
 *
 * The reference (antonior92/ip-nonlinear-solver) is pure Python and has no FFI;
 * its seam for this path is two Python signatures plus the scipy
 * LinearOperator protocol.  Every entry point below
 * names the reference interface it stands in for; the ctypes binding a
 * maintainer would add is shown in INTEGRATION.md and shipped in
 * ip-nonlinear-solver_b200/ipsolver_b200/_cabi.py.
 *
 * Conventions
 *   - plain pointers and sizes only; all host vectors are contiguous fp64;
 *   - every function returns an ipb_status (0 = ok); ipb_last_error() gives
 *     the message of the last failure on that context;
 *   - host buffers are never retained past the call that received them;
 *   - handles (ipb_csr, ipb_hess, ipb_proj, ipb_vec) are opaque device objects
 *     owned by the context that created them.
 */
#ifndef IPSOLVER_B200_H
#define IPSOLVER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ipb_ctx ipb_ctx;
typedef struct ipb_csr ipb_csr;
typedef struct ipb_hess ipb_hess;
typedef struct ipb_proj ipb_proj;
typedef struct ipb_vec ipb_vec;

typedef enum {
    IPB_OK = 0,
    IPB_ERR_INVALID = 1,          /* bad argument                                  */
    IPB_ERR_CUDA = 2,             /* CUDA runtime failure                          */
    IPB_ERR_TR_INFEASIBLE = 3,    /* qp_subproblem.py:516-517 (ValueError)         */
    IPB_ERR_NEG_CURVATURE = 4,    /* qp_subproblem.py:559-562 (ValueError)         */
    IPB_ERR_INNER_SOLVE = 5,      /* no solver for (A A') v = rhs converged         */
    IPB_ERR_NOMEM = 6,
    IPB_ERR_COMM = 7              /* multi-GPU exchange failure / timeout          */
} ipb_status;

/* Result of one projected-CG solve: the reference's ``info`` dict
 * (qp_subproblem.py:639-643) plus counters of the inner work. */
typedef struct {
    int64_t niter;            /* info['niter']                                   */
    int32_t stop_cond;        /* info['stop_cond'] 1..4                          */
    int32_t hits_boundary;    /* info['hits_boundary']                           */
    int32_t status;           /* ipb_status raised inside the kernel             */
    int32_t n_allvecs;        /* vectors stored when return_all was requested    */
    int64_t inner_iters;      /* inner iterations on A A' summed over all solves:
                               * CG iterations, or residual-checked sweeps of the
                               * banded direct solve                            */
    int64_t n_solves;         /* number of (A A')^-1 solves                      */
    int64_t n_refinements;    /* iterative-refinement rounds (projections.py:72) */
    int64_t n_spmv_h;         /* H products                                      */
    double  max_orth;         /* largest orthogonality() accepted                */
    double  tol_used;         /* tolerance actually used (default rule :530)     */
    double  kernel_ms;        /* device time of the solve kernel (CUDA events)   */
    int64_t n_sweeps;         /* sweeps of the banded direct solve (0: CG only)  */
    int32_t n_fallbacks;      /* banded solves given up for the CG (0 normally)  */
    int32_t reserved;         /* 1: the update p = -g + beta p ran inside H p     */
} ipb_pcg_info;

/* Counters of one projector application (Z, LS or Y). */
typedef struct {
    int64_t inner_iters;
    int64_t n_solves;
    int64_t n_refinements;
    int32_t status;
    int32_t n_fallbacks;      /* banded direct solves given up for CG (0 normally) */
    double  orth;             /* orthogonality(A, z) of the returned z (Z only)  */
    double  kernel_ms;
} ipb_proj_info;

/* ---- context ---------------------------------------------------------- */
/* One context per process and GPU.  device = CUDA ordinal. */
int ipb_ctx_create(int device, ipb_ctx **out);
int ipb_ctx_destroy(ipb_ctx *ctx);
const char *ipb_last_error(const ipb_ctx *ctx);
/* grid of the persistent kernels (blocks, threads/block, SM count). */
int ipb_ctx_grid(const ipb_ctx *ctx, int *blocks, int *threads, int *sms);
/* cumulative count of kernels launched by this library on this context. */
int64_t ipb_ctx_launches(const ipb_ctx *ctx);
int ipb_ctx_sync(ipb_ctx *ctx);

/* ---- CSR matrices (A, the sparse part of H) ---------------------------- */
/* Replaces scipy.sparse.csr_matrix as consumed by A.dot / A.T.dot / H.dot
 * (csr_matvec call sites: projections.py:52,66,67,76,77,84,88;
 * qp_subproblem.py:376,379,410; equality_constrained_sqp.py:86,125,138,140).
 * index_bytes is 4 or 8 (the barrier Jacobian arrives with int64 indices,
 * tr_interior_point.py:184-193).  with_transpose builds the explicit CSR of
 * the transpose on the device (deterministic A'.y, no atomics). */
int ipb_csr_create(ipb_ctx *ctx, int64_t rows, int64_t cols, int64_t nnz,
                   const void *indptr, const void *indices, int index_bytes,
                   const double *data, int with_transpose, ipb_csr **out);
/* same sparsity pattern, new values (one upload per outer iteration). */
int ipb_csr_update_values(ipb_ctx *ctx, ipb_csr *m, const double *data);
int ipb_csr_destroy(ipb_ctx *ctx, ipb_csr *m);
/* y = A x (host vectors).  */
int ipb_csr_matvec(ipb_ctx *ctx, const ipb_csr *m, const double *x, double *y);
/* x = A' y (host vectors); needs with_transpose. */
int ipb_csr_rmatvec(ipb_ctx *ctx, const ipb_csr *m, const double *y, double *x);
/* device-resident variants */
int ipb_csr_matvec_dev(ipb_ctx *ctx, const ipb_csr *m, const ipb_vec *x, ipb_vec *y);
int ipb_csr_rmatvec_dev(ipb_ctx *ctx, const ipb_csr *m, const ipb_vec *y, ipb_vec *x);
/* ||A||_F  (scipy.sparse.linalg.norm(A, 'fro'), projections.py:44) */
int ipb_csr_fro_norm(ipb_ctx *ctx, const ipb_csr *m, double *out);

/* ---- Lagrangian Hessian operator --------------------------------------- */
/* Replaces the closure-backed LinearOperator of
 * _canonical_constraint.py:119-139 and the barrier version
 * tr_interior_point.py:222-241:
 *     H v = [ (S + D) v_x ; diag_tail * v_s ]
 * S: CSR n_x x n_x (may be NULL), D: dense row-major n_x x n_x (may be NULL),
 * diag_tail: n - n_x entries (S Hs S, tr_interior_point.py:206-220). */
int ipb_hess_create(ipb_ctx *ctx, int64_t n, int64_t n_x, ipb_hess **out);
int ipb_hess_set_csr(ipb_ctx *ctx, ipb_hess *h, ipb_csr *s);       /* borrowed */
int ipb_hess_set_dense(ipb_ctx *ctx, ipb_hess *h, const double *d); /* copied  */
int ipb_hess_set_diag_tail(ipb_ctx *ctx, ipb_hess *h, const double *diag);
int ipb_hess_destroy(ipb_ctx *ctx, ipb_hess *h);
int ipb_hess_matvec(ipb_ctx *ctx, const ipb_hess *h, const double *x, double *y);
int ipb_hess_matvec_dev(ipb_ctx *ctx, const ipb_hess *h, const ipb_vec *x, ipb_vec *y);

/* ---- projectors --------------------------------------------------------- */
/* Replaces projections(A, method, orth_tol, max_refin, tol) for sparse A
 * (projections.py:290-406; NormalEquation :58-90, AugmentedSystem :93-172).
 * (A A')^-1 is applied by a conjugate-gradient iteration on the row-normalised
 * A (Jacobi-preconditioned CG on A A'), or -- when A A' is banded -- by the
 * banded direct solve below with the CG as its fallback.  A is borrowed and
 * must have been created with_transpose. */
int ipb_proj_create(ipb_ctx *ctx, ipb_csr *A, double orth_tol, int max_refin,
                    double inner_rtol, int inner_maxit, ipb_proj **out);
/* How (A A') is applied inside the inner CG: mode 0 (default) keeps the explicit
 * product M = A A' in HBM when that moves fewer bytes per application than A and
 * A' do (12 nnz(M) < 24 nnz(A) + 16 n), mode 1 always applies A then A', mode 2
 * always uses M.  The pattern of M is built once per sparsity pattern, its
 * values on every refresh.  ipb_proj_aat_nnz: nnz(M), or -1 when not held. */
int ipb_proj_set_aat_mode(ipb_ctx *ctx, ipb_proj *p, int mode);
int64_t ipb_proj_aat_nnz(const ipb_proj *p);
/* Banded direct solve.  When M = A~ A~' is banded (half-bandwidth 1..7 -- a
 * Jacobian whose rows touch a narrow window of columns), its Cholesky factor
 * is kept in HBM and (A A')^-1 is applied by overlapping-chunk substitution
 * sweeps: exact to round-off for these matrices (the reference applies a
 * factorisation too: CHOLMOD projections.py:62, SuperLU :102), every solve is
 * verified by its residual and refined, and a solve whose residual does not
 * contract falls back to the CG (and switches the factor off).
 *   mode 0: use it when M is banded (default), 1: never;
 *   chunk_rows / overlap: 0 = defaults (chunk sized to the GPU, overlap 48).
 * ipb_proj_band_info: out[6] = {in use, half-bandwidth of M (-1 unknown),
 *   chunk rows, overlap, groups of 32 chunks, factor still trusted}. */
int ipb_proj_set_band(ipb_ctx *ctx, ipb_proj *p, int mode, int chunk_rows, int overlap);
int ipb_proj_band_info(ipb_ctx *ctx, const ipb_proj *p, int32_t *out);
/* call after ipb_csr_update_values(A): refreshes row norms, ||A||_F, A', M and
 * the banded factor. */
int ipb_proj_refresh(ipb_ctx *ctx, ipb_proj *p);
int ipb_proj_destroy(ipb_ctx *ctx, ipb_proj *p);
/* Z: z = x - A'(AA')^-1 A x with refinement (null_space closures :65-80, :114-142) */
int ipb_proj_null_space(ipb_ctx *ctx, ipb_proj *p, const double *x, double *z,
                        ipb_proj_info *info);
/* LS: v = (AA')^-1 A x  (:83-84, :148-156) */
int ipb_proj_least_squares(ipb_ctx *ctx, ipb_proj *p, const double *x, double *v,
                           ipb_proj_info *info);
/* Y: y = A'(AA')^-1 b  (:87-88, :162-170) */
int ipb_proj_row_space(ipb_ctx *ctx, ipb_proj *p, const double *b, double *y,
                       ipb_proj_info *info);
int ipb_proj_null_space_dev(ipb_ctx *ctx, ipb_proj *p, const ipb_vec *x, ipb_vec *z,
                            ipb_proj_info *info);
int ipb_proj_least_squares_dev(ipb_ctx *ctx, ipb_proj *p, const ipb_vec *x, ipb_vec *v,
                               ipb_proj_info *info);
int ipb_proj_row_space_dev(ipb_ctx *ctx, ipb_proj *p, const ipb_vec *b, ipb_vec *y,
                           ipb_proj_info *info);
/* orthogonality(A, g) = ||A g|| / (||A||_F ||g||)  (projections.py:23-55) */
int ipb_orthogonality(ipb_ctx *ctx, ipb_proj *p, const double *g, double *out);

/* ---- the trust-region EQP solve ----------------------------------------- */
/* Replaces projected_cg(H, c, Z, Y, b, trust_radius, lb, ub, tol, max_iter,
 * max_infeasible_iter, return_all) (qp_subproblem.py:416-643).
 *   lb / ub may be NULL (= -inf / +inf); trust_radius may be +inf;
 *   tol < 0 or NaN selects the default rule (:529-530);
 *   max_iter / max_infeasible_iter < 0 select the default n - m (:537-542);
 *   allvecs (may be NULL) receives up to allvecs_cap iterates of n doubles
 *   (return_all; :508-509, :618-619), info->n_allvecs says how many.
 * Returns IPB_ERR_TR_INFEASIBLE / IPB_ERR_NEG_CURVATURE where the reference
 * raises ValueError. */
int ipb_projected_cg(ipb_ctx *ctx, const ipb_hess *H, ipb_proj *P,
                     const double *c, const double *b, double trust_radius,
                     const double *lb, const double *ub, double tol,
                     int64_t max_iter, int64_t max_infeasible_iter,
                     double *x, double *allvecs, int64_t allvecs_cap,
                     ipb_pcg_info *info);
/* inputs and result resident in HBM (only ``info`` crosses PCIe). */
int ipb_projected_cg_dev(ipb_ctx *ctx, const ipb_hess *H, ipb_proj *P,
                         const ipb_vec *c, const ipb_vec *b, double trust_radius,
                         const ipb_vec *lb, const ipb_vec *ub, double tol,
                         int64_t max_iter, int64_t max_infeasible_iter,
                         ipb_vec *x, ipb_pcg_info *info);

/* Replaces modified_dogleg(A, Y, b, trust_radius, lb, ub)
 * (qp_subproblem.py:320-413).  lb / ub may be NULL. */
int ipb_modified_dogleg(ipb_ctx *ctx, ipb_proj *P, const double *b,
                        double trust_radius, const double *lb, const double *ub,
                        double *x, ipb_proj_info *info);

/* One outer iteration of equality_constrained_sqp between two user callbacks
 * (equality_constrained_sqp.py:113-140) with its vectors resident in HBM, in
 * two calls because the host decides the tangential radius:
 *   ipb_sqp_normal_step:     dn = modified_dogleg(A, Y, b, normal_radius,
 *                                 box_factor*lb, box_factor*ub)          (:113)  -> dn
 *   host:                    radius_t = sqrt(radius^2 - norm(dn)^2)      (:127)
 *   ipb_sqp_tangential_step: c_t = H dn + c                              (:125)
 *                            dt = projected_cg(H, c_t, Z, Y, 0, radius_t,
 *                                              lb - dn, ub - dn)         (:128-132)
 *                            d = dn + dt, H d, A d                       (:135-140)
 * c / b: host vectors to upload, or NULL = unchanged since the last call on
 * this context (they stay resident); lb / ub: NULL = unbounded, uploaded only
 * when bounds_changed or never seen.  Vectors the device already holds (dn,
 * c_t, the shifted bounds, d) are never uploaded again; the scalars of the
 * iteration are left to the host on the downloaded d, H d, A d (Ad may be NULL
 * when m = 0).  ipb_sqp_multipliers: v = -LS c, ||c + A'v||_inf (:227, :238);
 * its c becomes the resident gradient of the following steps.
 * Row-partitioned contexts (ipb_comm_init): the same three calls, made by every
 * rank with ITS OWN BLOCKS of c, b, lb, ub, dn, d, Hd, Ad, v (lengths = the
 * projector's own n, m); products exchange ghosts and the solves reduce their
 * scalars across the ranks inside the kernels.  ipb_sqp_multipliers then
 * returns the rank's share of ||c + A'v||_inf -- the caller takes the maximum
 * over the ranks.  ipb_modified_dogleg / ipb_projected_cg with host vectors stay
 * single-GPU (use these or the *_dev entry points). */
int ipb_sqp_normal_step(ipb_ctx *ctx, ipb_proj *P, const double *c, const double *b,
                        double normal_radius, const double *lb, const double *ub,
                        int bounds_changed, double box_factor, double *dn, ipb_proj_info *info);
int ipb_sqp_tangential_step(ipb_ctx *ctx, const ipb_hess *H, ipb_proj *P,
                            double tangential_radius, double *d, double *Hd, double *Ad,
                            ipb_pcg_info *info);
int ipb_sqp_multipliers(ipb_ctx *ctx, ipb_proj *P, const double *c, double *v,
                        double *optimality, ipb_proj_info *info);

/* Replaces box_sphere_intersections(z, d, lb, ub, trust_radius, entire_line)
 * (qp_subproblem.py:237-303); with lb = ub = NULL it is sphere_intersections
 * (:66-149), with trust_radius = +inf it is box_intersections (:152-234, the
 * form the SQP second-order correction calls, equality_constrained_sqp.py:177). */
int ipb_box_sphere_intersections(ipb_ctx *ctx, int64_t n, const double *z,
                                 const double *d, const double *lb, const double *ub,
                                 double trust_radius, int entire_line, double *ta,
                                 double *tb, int *intersect);
/* The same on device vectors.  On a row-partitioned context the vectors are
 * the rank's own blocks and the interval is reduced over the ranks inside the
 * kernel (every rank gets the same ta, tb, intersect). */
int ipb_box_sphere_intersections_dev(ipb_ctx *ctx, const ipb_vec *z, const ipb_vec *d,
                                     const ipb_vec *lb, const ipb_vec *ub, double trust_radius,
                                     int entire_line, double *ta, double *tb, int *intersect);

/* ---- row-partitioned multi-GPU runs (one process per GPU) ---------------- */
/* The reference is single-process (no communication
 * backend); this is the B200-side extension north_star asks for.  Rank g owns
 * a contiguous block of variables (n-space) and of constraints (m-space); all
 * solver vectors of a rank live in a symmetric arena laid out
 * [ghost_lo | own | ghost_hi], matrices are given per rank with column indices
 * relative to the own block (negative / past the end = ghost entries).  Inside
 * the kernels the owners send ghost entries into their peers' ghost mailboxes
 * over NVLink peer memory before each product (no barrier between the ranks:
 * only the row blocks that read ghosts wait for them), and reductions go
 * through per-rank mailboxes in peer memory, summed in rank order
 * (bit-identical on all ranks).
 *   own / own_max / ghost_lo / ghost_hi : 2 entries each, [0] n-space, [1] m-space
 *     (own: this rank's block length; own_max, ghost_*: maxima over the ranks,
 *     identical on all ranks; ghost_lo even; my_ghost_*: the entries this rank
 *     really holds below / above its own block)
 *   n_minus_m_global : n - m of the whole problem (iteration cap, qp_subproblem.py:537)
 *   ipc_handle_out   : 64 bytes to be all-gathered and passed to ipb_comm_connect
 * Host-vector entry points are single-GPU; partitioned runs use *_dev calls. */
int ipb_comm_init(ipb_ctx *ctx, int rank, int size, const int64_t *own,
                  const int64_t *own_max, const int64_t *ghost_lo, const int64_t *ghost_hi,
                  const int64_t *my_ghost_lo, const int64_t *my_ghost_hi,
                  int64_t n_minus_m_global, int64_t arena_bytes, void *ipc_handle_out);
int ipb_comm_connect(ipb_ctx *ctx, const void *handles /* size x 64 bytes, rank order */);
/* entries: count x {own index, peer rank, index on the peer (own-relative), slot in the
 * peer's ghost mailbox = ghost_lo capacity + that index if it is negative, else ghost_lo
 * capacity + (index - the peer's own length)} (int32) */
int ipb_comm_set_send_list(ipb_ctx *ctx, int space, int64_t count, const int32_t *entries);
/* vector space of a matrix's columns (0: n-space, 1: m-space) -- selects the halo list */
int ipb_csr_set_space(ipb_ctx *ctx, ipb_csr *m, int space);
/* projector from this rank's parts (created with with_transpose = 0): rows of A
 * (raw and row-normalised values), rows of A' for the own variables (raw and
 * normalised), optionally the own rows of M = A~ A~'; dscale = 1/||a_i|| of
 * the own rows, fro = ||A||_F of the whole matrix. */
int ipb_proj_create_parts(ipb_ctx *ctx, ipb_csr *A_raw, ipb_csr *A_norm, ipb_csr *At_raw,
                          ipb_csr *At_norm, ipb_csr *M, const double *dscale, double fro,
                          double orth_tol, int max_refin, double inner_rtol, int inner_maxit,
                          ipb_proj **out);

/* Banded direct solve for a projector made from parts (see ipb_proj_set_band):
 * M_ext = rows [own_lo - ext_off, own_hi + margin) of M = A~ A~' as a standalone
 * square CSR (created with_transpose = 0; columns relative to its first row,
 * entries outside dropped).  The margins (>= overlap + 48 rows where the matrix
 * has them) let the factor recurrence forget its truncated start.  ghost_below /
 * ghost_above (<= overlap): rows of the neighbours the sweeps read next to the
 * own block -- the m-space ghost zones passed to ipb_comm_init must hold exactly
 * these rows, contiguously. */
int ipb_proj_set_band_parts(ipb_ctx *ctx, ipb_proj *p, const ipb_csr *M_ext, int64_t ext_off,
                            int ghost_below, int ghost_above, int chunk_rows, int overlap);

/* ---- device vectors ------------------------------------------------------ */
int ipb_vec_create(ipb_ctx *ctx, int64_t len, ipb_vec **out);
int ipb_vec_destroy(ipb_ctx *ctx, ipb_vec *v);
int ipb_vec_upload(ipb_ctx *ctx, ipb_vec *v, const double *host);
int ipb_vec_download(ipb_ctx *ctx, const ipb_vec *v, double *host);
int64_t ipb_vec_len(const ipb_vec *v);
void *ipb_vec_ptr(ipb_vec *v);   /* raw device pointer (for torch interop) */

/* ---- pinned host memory --------------------------------------------------- */
/* Page-locked buffers for results: the host mirror hands large outputs (x of
 * projected_cg, projector results) back in such buffers, so the device-to-host
 * copy runs at PCIe rate instead of through a pageable bounce. */
int ipb_host_alloc(ipb_ctx *ctx, int64_t bytes, void **out);
int ipb_host_free(ipb_ctx *ctx, void *p);

/* ---- measurement helpers (bench.py) -------------------------------------- */
/* CUDA-event stopwatch on the context's stream (the stream every kernel of
 * this library is launched on; torch.cuda.Event would not see it). */
int ipb_timer_start(ipb_ctx *ctx);
int ipb_timer_stop(ipb_ctx *ctx, double *ms);
/* Runs ``iters`` inner CG iterations on A A' (fixed count, no convergence
 * test) starting from rhs = A x; returns device milliseconds.  This is the
 * dominant loop of the path, timed in isolation. */
int ipb_bench_inner_iterations(ipb_ctx *ctx, ipb_proj *p, const ipb_vec *x,
                               int iters, double *ms);

/* Barrier timeline: with capacity > 0 every following launch records the
 * %globaltimer (ns) at which CTA 0 leaves each grid barrier; ipb_trace_read
 * returns the stamps of the LAST launch (count = barriers passed, of which at
 * most ``capacity`` are stored).  Consecutive differences are the durations of
 * the phases inside the persistent kernel.  capacity = 0 switches it off. */
int ipb_trace_enable(ipb_ctx *ctx, int capacity);
int ipb_trace_read(ipb_ctx *ctx, uint64_t *out, int capacity, int *count);

/* Repeats ONE phase of the inner iteration ``reps`` times inside one launch,
 * a grid barrier after each (0: fused vector update, 1: A' r, 2: A t, 3: M r,
 * 4: barrier only, 5: the bulk-copy pipeline of M r without the arithmetic,
 * 6: barrier + cross-GPU all-reduce, 7: barrier + ghost push of an m-vector,
 * 8: the row arithmetic of M r without the copy pipeline, 9: one sweep of the
 * banded direct solve);
 * returns device milliseconds.  Results are discarded and
 * the projector's work vectors are overwritten. */
int ipb_bench_phase(ipb_ctx *ctx, ipb_proj *p, int phase, int reps, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* IPSOLVER_B200_H */
