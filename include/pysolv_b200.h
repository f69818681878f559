/* libpysolv_b200 -- C ABI of the B200-native PySolvers solve-phase hot path.
 *
 * This is the drop-in boundary: the Python classes in
 * pysolvers_b200/ (same names and semantics as PySolvers.Linear / .Nonlinear)
 * call these entry points through ctypes; a maintainer of the reference would
 * bind exactly these from PySolvers/Linear/*.py (INTEGRATION.md shows the
 * stubs).  Plain pointers and sizes only -- no torch, no C++ types.
 *
 * Conventions
 *  - every pointer named d_* is a CUDA DEVICE pointer owned by the caller
 *    (torch tensors in the Python host layer) and must stay alive until the
 *    call -- or, for psb_csr_create / psb_trsv_create, the handle -- is done;
 *    fp64 vectors must be 16-byte aligned (torch allocations are);
 *  - values are fp64, indices int32 (scipy CSR as the reference produces it);
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *    calls only ENQUEUE work unless the comment says they synchronise;
 *  - every function returns PSB_OK (0) or a negative error code;
 *    psb_last_error() returns the text of the calling thread's last failure;
 *  - one host thread per device/rank; handles are not thread-safe.
 */
#ifndef PYSOLV_B200_H
#define PYSOLV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSB_OK            0
#define PSB_ERR_CUDA     -1   /* a CUDA runtime call or kernel launch failed   */
#define PSB_ERR_ARG      -2   /* bad argument (null, negative size, alignment) */
#define PSB_ERR_UNSUPP   -3   /* valid request the GPU path does not support   */
#define PSB_ERR_NCCL     -4   /* NCCL missing or a collective failed           */

typedef struct psb_csr*   psb_csr_t;    /* CSR matrix + chosen SpMV kernel      */
typedef struct psb_trsv*  psb_trsv_t;   /* level-analysed triangular factor     */
typedef struct psb_prec*  psb_prec_t;   /* preconditioner: z = M^-1 r on device */
typedef struct psb_comm*  psb_comm_t;   /* NCCL communicator + side stream      */
typedef struct psb_dist*  psb_dist_t;   /* row-partitioned matrix + halo plan   */

/* ------------------------------------------------------------------ misc -- */
int         psb_version(void);
const char* psb_last_error(void);
/* kernels launched by this library since load (bench.py's gpu_launches). */
long long   psb_launch_count(void);

/* ------------------------------------------------------------------- CSR -- */
/* SpMV kernel kinds chosen from the row-length histogram at create time. */
#define PSB_SPMV_STREAM      1  /* tiles of consecutive rows staged into shared memory
                                  by bulk async copies (TMA engine), double-buffered;
                                  one thread per row sums in STORED order (bit-equal to
                                  scipy csr_matvec)                                  */
#define PSB_SPMV_VECTOR      2  /* sub-warp per row, shuffle reduction               */
#define PSB_SPMV_STREAM_LSU  3  /* first-generation STREAM: coalesced LDG.128 of the
                                  tile, products parked in shared memory            */
#define PSB_SPMV_MERGE       4  /* merge-path: the n_rows + nnz work items are split evenly
                                  over CTAs and threads wherever the row boundaries fall
                                  (a few very long rows among short ones); deterministic,
                                  equal to scipy to rounding (not bit for bit)        */
#define PSB_SPMV_TILE512    16  /* OR-ed into a STREAM kind for psb_csr_set_kind:
                                  512-row tiles instead of 256                       */

/* Wraps caller-owned device arrays (no copy).  Synchronises `stream` once to
 * read back the row-length statistics that pick the kernel.
 * Replaces: the scipy csr_matrix operand of mvmult(),
 * PySolvers/Linear/IterativeLinearSolver.py:94-106. */
int psb_csr_create(int64_t n_rows, int64_t n_cols, int64_t nnz,
                   const int32_t* d_rowptr, const int32_t* d_colind,
                   const double* d_vals, void* stream, psb_csr_t* out);
/* Process-wide option for the matrices created afterwards: keep a second copy of the column
 * indices of banded matrices as 16-bit distances from the diagonal (10 instead of 12 bytes per
 * entry in the STREAM kernels; bit-identical results).  Off by default: measured gain 4 % on the
 * stand-alone SpMV, none in the persistent PCG kernel (profiles/round1f_notes.md). */
int psb_csr_set_cols16(int enable);
int psb_csr_destroy(psb_csr_t A);
/* info[0]=kernel kind, [1]=max row length, [2]=max nnz per 256-row tile,
 * [3]=rows per tile, [4]=vector width (VECTOR kind), [5]=max grid size,
 * [6]=arrays 16-byte aligned, [7]=max nnz per 512-row tile. */
int psb_csr_info(psb_csr_t A, int64_t info[8]);
/* Force a kernel kind (testing / A-B timing); PSB_ERR_UNSUPP if impossible. */
int psb_csr_set_kind(psb_csr_t A, int kind);

/* y = A x.   Replaces `A*x`, IterativeLinearSolver.py:104. */
int psb_spmv(psb_csr_t A, const double* d_x, double* d_y, void* stream);
/* y = A x and *d_dot = x . y (deterministic two-stage fp64 reduction).
 * Replaces PCGSolver.py:111-113 (Ap = A*p; pTAp = dot(p, Ap)). */
int psb_spmv_dot(psb_csr_t A, const double* d_x, double* d_y, double* d_dot,
                 void* stream);
/* y = f - A x   (VCycleManager.py:45, VCycleSolver.py:84, GMRESSolver.py:163) */
int psb_spmv_residual(psb_csr_t A, const double* d_x, const double* d_f,
                      double* d_y, void* stream);
/* y += A x      (prolongation + correction, VCycleManager.py:55) */
int psb_spmv_add(psb_csr_t A, const double* d_x, double* d_y, void* stream);
/* x_new = x + omega * dinv .* (f - A x)   (ClassicSmoothers.py:12-14; the
 * reference has omega = 1).  x_new must not alias x. */
int psb_jacobi_sweep(psb_csr_t A, const double* d_dinv, double omega,
                     const double* d_f, const double* d_x, double* d_xnew,
                     void* stream);

/* ---------------------------------------------------------- vector ops -- */
/* *d_out = x . y ; deterministic.  (np.dot / numpy.linalg.norm call sites,
 * PCGSolver.py:86,102,125,134) */
int psb_dot(int64_t n, const double* d_x, const double* d_y, double* d_out,
            void* stream);

/* ------------------------------------------------ setup helpers (host) -- */
/* Phase 1 of the reference's smoothed-aggregation coarsening (SmoothedAggregation.py:84-89), a
 * sequential sweep over the nodes, as host code: a free node whose whole strong neighbourhood
 * (CSR lists h_s_ptr / h_s_cols, int64) is still free founds the next aggregate.  h_agg_of (int64,
 * -1 = free) comes in with the isolated nodes assigned and *n_agg with their count; new roots are
 * appended to h_roots (capacity n).  No device work. */
int psb_sa_phase1(int64_t n, const int64_t* h_s_ptr, const int64_t* h_s_cols, int64_t* h_agg_of,
                  int64_t* h_roots, int64_t* n_agg);

/* --------------------------------------------- input generators (device) -- */
/* The finite-difference Laplacians of the benchmark configurations assembled directly in HBM
 * (examples/FDLaplacian2D.py:5-23 and its 7-point 3-D extension): CSR rows [row_lo, row_hi) of the
 * m^dim grid matrix with the reference's stored column order and global column numbers; values
 * `diag` on the diagonal and `off` elsewhere.  psb_stencil_nnz sizes colind / vals (rowptr needs
 * row_hi - row_lo + 1 entries); -1 on bad arguments. */
int64_t psb_stencil_nnz(int dim, int64_t m, int64_t row_lo, int64_t row_hi);
int psb_stencil_fill(int dim, int64_t m, int64_t row_lo, int64_t row_hi, double diag, double off,
                     int32_t* d_rowptr, int32_t* d_colind, double* d_vals, void* stream);

/* ---------------------------------------------- sparse triangular solves -- */
/* Level analysis + repacking of a triangular CSR factor given in HOST memory (the
 * factors come from the reference's SuperLU setup on the CPU): rows are sorted
 * level-major, the strictly triangular part is stored SELL-32 in that order, the
 * diagonal is split out, everything is uploaded once.  `lower` != 0: entries with
 * col < row are the dependencies (col > row ignored); otherwise the reverse.
 * `unit_diag` != 0: the diagonal is taken as 1 (stored diagonal entries ignored).
 * Synchronises `stream`.  Replaces the factor objects of
 * PySolvers/Linear/ICPreconditioner.py:45-56 / ILUTPreconditioner.py:51-53. */
int psb_trsv_create(int64_t n, const int32_t* h_rowptr, const int32_t* h_colind,
                    const double* h_vals, int lower, int unit_diag, void* stream,
                    psb_trsv_t* out);
int psb_trsv_destroy(psb_trsv_t T);
/* info[0]=n, [1]=levels, [2]=off-diagonal nnz, [3]=packed (padded) nnz,
 * [4]=lower, [5]=unit_diag, [6]=32-row groups. */
int psb_trsv_info(psb_trsv_t T, int64_t info[8]);
/* Which solve kernel the analysis chose, and the data of the shared-memory window kernels:
 * info[0]=kernel (0 grid-wide, hand-over through L2; 1 one CTA, hand-over through shared
 * memory; 2 a cluster of 4 CTAs, window replicated through distributed shared memory),
 * [1]=window slots, [2]=dependencies older than the window (read from the global
 * vector), [3]=largest distance of a dependency in processing order, [4]=forced kernel or -1,
 * [5]=entries per lane of a staging buffer, [6]=the cluster kernel may be forced. */
int psb_trsv_info2(psb_trsv_t T, int64_t info[8]);
/* Force a kernel for this factor (0 / 1 / 2), or -1 to return to the analysis' choice.  All
 * kernels produce bit-identical results. */
int psb_trsv_set_kernel(psb_trsv_t T, int kernel);
/* Debugging aid of the one-CTA kernel: when d_trace (device, 12 * groups int64) is not NULL every
 * chunk records clock64 at its start, at its first missing dependency, after sleeping, at its
 * last batch of entries, when its last dependency arrived and after its store, plus the entry
 * index of the first miss and the entries per lane.  With the grid kernel the buffer needs
 * 3 * groups int64 and every chunk records {%globaltimer when claimed, when done, first item}
 * (items are level-major: psb_trsv_get_levels gives the level boundaries). */
int psb_trsv_set_trace(psb_trsv_t T, long long* d_trace);
/* Copies out the level sets (host arrays of levels+1 and n int32): level of a row
 * = 1 + max level of its dependencies; rows level-major, ascending in a level. */
int psb_trsv_get_levels(psb_trsv_t T, int32_t* h_level_ptr, int32_t* h_level_rows);
/* x = T^-1 b in one persistent launch (x must not alias b).  Replaces
 * scipy spsolve_triangular, ICPreconditioner.py:61,63. */
int psb_trsv_solve(psb_trsv_t T, const double* d_b, double* d_x, void* stream);
/* *h_flag != 0 if a solve gave up waiting for a dependency (synchronises). */
int psb_trsv_error(psb_trsv_t T, int32_t* h_flag);

/* ------------------------------------------------------ preconditioners -- */
/* z = L^-T (L^-1 r)   (ICRightPreconditioner.applyRight, ICPreconditioner.py:58-63).
 * The factors stay owned by the caller and must outlive the preconditioner. */
int psb_ic_create(psb_trsv_t L, psb_trsv_t Lt, psb_prec_t* out);
/* z = Pc U^-1 L^-1 Pr r with Pr[perm_r[i], i] = 1, Pc[i, perm_c[i]] = 1 and unit
 * lower L -- SuperLU.solve (ILUTPreconditioner.py:67,78).  perm arrays: HOST int32. */
int psb_ilu_create(psb_trsv_t L, psb_trsv_t U, const int32_t* h_perm_r,
                   const int32_t* h_perm_c, void* stream, psb_prec_t* out);
/* The same exact LU solve with the trailing n - n1 rows of L and U treated as DENSE blocks
 * whose inverses (row-major (n-n1)^2 fp64, device, kept by the caller) are applied as triangular
 * matrix-vector products: L = [L11 0; L21 L22], U = [U11 U12; 0 U22]; L11/U11 are factors of
 * order n1 (unit lower / upper), L21 is (n-n1) x n1 and U12 n1 x (n-n1) CSR; n1 may be 0
 * (then those four may be NULL).  Replaces the coarsest-level spsolve of the V-cycle
 * (PySolvers/Linear/VCycleManager.py:34-37): the dense tail of the factors is one dependency
 * level per row for a sparse triangular solve, and plain HBM streaming as a GEMV. */
int psb_splitlu_create(int64_t n, int64_t n1, psb_trsv_t L11, psb_trsv_t U11, psb_csr_t L21,
                       psb_csr_t U12, const double* d_invL22, const double* d_invU22,
                       const int32_t* h_perm_r, const int32_t* h_perm_c, void* stream,
                       psb_prec_t* out);
/* General form: L and U are split independently, each in its own symmetric permutation chosen
 * by the host (the dense block of L = the rows of its last dependency levels, the one of U = the
 * rows of its first ones).  L-stage position p takes v[h_map_in[p]] (n1L sparse rows, then
 * n - n1L dense ones); U-stage element p takes the L-stage result at h_map_mid[p] (NULL: same
 * order); result[h_map_out[p]] = x[p].  Maps: HOST int32[n]. */
int psb_splitlu2_create(int64_t n, int64_t n1L, int64_t n1U, psb_trsv_t L11, psb_trsv_t U11,
                        psb_csr_t L21, psb_csr_t U12, const double* d_invL22,
                        const double* d_invU22, const int32_t* h_map_in,
                        const int32_t* h_map_mid, const int32_t* h_map_out, void* stream,
                        psb_prec_t* out);
/* Optional block-diagonal stage of a split LU whose sparse leading blocks had their supernodes
 * collapsed on the host (L11 = L~ blockdiag(D), U11 = blockdiag(D) U~ with identity diagonal blocks in
 * L~ / U~; pysolvers_b200/Linear/supernodes.py): y = B x with B = identity outside the supernodes;
 * row r reads x[h_row0[r] + c] for c in [h_c_lo[r], h_c_hi[r]) with the weights
 * h_vals[h_off[r] + c - h_c_lo[r]].  upper = 0: applied after the L11 solve; 1: before the U11 solve.
 * All arrays HOST, one entry per row of the leading block. */
int psb_splitlu_set_blockdiag(psb_prec_t P, int upper, int64_t n_rows, const int32_t* h_row0,
                              const int32_t* h_c_lo, const int32_t* h_c_hi, const int64_t* h_off,
                              const double* h_vals, int64_t n_vals, void* stream);
/* Dependency levels of a triangular CSR matrix in HOST memory (no device work): level(i) = 1 + max
 * level of the rows row i depends on, 0 if none. */
int psb_tri_levels(int64_t n, const int32_t* h_rowptr, const int32_t* h_colind, int lower,
                   int32_t* h_level);
/* Heights of the rows of an UPPER triangular CSR matrix (HOST) in its elimination tree = the levels
 * of U^T, computed without transposing: h(i) = 1 + max h(k) over rows k < i with U[k, i] != 0. */
int psb_tri_heights_upper(int64_t n, const int32_t* h_rowptr, const int32_t* h_colind, int32_t* h_height);
/* z = M^-1 r (z must not alias r).  Preconditioner.applyRight / applyLeft. */
int psb_prec_apply(psb_prec_t P, const double* d_r, double* d_z, void* stream);
int psb_prec_destroy(psb_prec_t P);

/* ------------------------------------------------------------------ PCG -- */
/* status codes written to psb_solve_result.status */
#define PSB_CONVERGED        0  /* ||r|| <= tau ||b|| (or maxiter reached with
                                   fail_on_maxiter == 0: PCGSolver.py:129-131)   */
#define PSB_MAXITER          1  /* maxiter iterations without convergence       */
#define PSB_BREAKDOWN_UR     2  /* dot(u,r) == 0 before the loop, :104-105      */
#define PSB_BREAKDOWN_PAP    3  /* dot(p,Ap) == 0 at iteration k, :114-115      */
#define PSB_TRIVIAL          4  /* b == 0 -> x = 0, :87-88                      */
#define PSB_GMRES_FALSE_CONV 5  /* recursive residual met tau but the true one
                                   did not (GMRESSolver.py:167-174)             */

typedef struct psb_solve_result {
  int32_t status;      /* one of the codes above                               */
  int32_t k;           /* 0-based index of the last executed iteration         */
  int32_t n_hist;      /* number of residual norms written to d_hist           */
  int32_t lucky;       /* GMRES: Arnoldi breakdown flag                        */
  double  norm_r;      /* last residual norm (true residual for GMRES)         */
  double  norm_b;
  double  norm_r_rec;  /* GMRES: last recursive (implicit) residual            */
} psb_solve_result;

/* bytes of fp64 workspace psb_pcg_solve needs for an n-row system */
int64_t psb_pcg_workspace_bytes(int64_t n, int has_prec);

/* Whole PCG solve on the device: x0 = 0, loop of PySolvers/Linear/PCGSolver.py:97-142.
 * prec == NULL is the identity (u aliases r, Preconditioner.py:58-68).
 * d_hist receives ||r_k|| for k = 0.. (needs maxiter doubles).  Synchronises
 * `stream` while polling the convergence flag and before returning. */
int psb_pcg_solve(psb_csr_t A, psb_prec_t prec, const double* d_b, double* d_x,
                  void* d_work, int64_t work_bytes, int32_t maxiter, double tau,
                  int32_t fail_on_maxiter, double* d_hist,
                  psb_solve_result* result, void* stream);

/* Profiling hook of the persistent PCG kernel (csrc/pcg_mega.cu; no reference counterpart):
 * solves launched afterwards record %globaltimer stamps of iterations [first_iter,
 * first_iter + n_iters) into d_buf, laid out [iteration][CTA][6] uint64: 0 phase A starts,
 * 1 phase A done, 2 p.Ap reduced, 3 phase B done, 4 r.r reduced.  d_buf == NULL disables. */
int psb_debug_mega_timeline(void* d_buf, int32_t first_iter, int32_t n_iters);

/* ---------------------------------------------------------------- GMRES -- */
#define PSB_ORTH_CGS2  1  /* classical Gram-Schmidt twice, batched dots (default)   */
#define PSB_ORTH_MGS   2  /* modified Gram-Schmidt in the reference's order
                             (GMRESSolver.py:110-112); for un-preconditioned parity */

int64_t psb_gmres_workspace_bytes(int64_t n, int32_t maxiter);

/* Whole right-preconditioned, un-restarted GMRES solve on the device
 * (PySolvers/Linear/GMRESSolver.py:75-174): Arnoldi + Givens loop, then
 * x = M^-1 (Q y) and the true residual check.  d_hist receives |g_{k+1}| per
 * iteration.  result->norm_r is the TRUE residual norm, norm_r_rec the recursive
 * one; status PSB_GMRES_FALSE_CONV when only the latter met tau.  At maxiter the
 * reference raises NameError (:180); here x_k of the last iteration is returned
 * with status PSB_MAXITER.  Synchronises `stream`. */
int psb_gmres_solve(psb_csr_t A, psb_prec_t prec, const double* d_b, double* d_x,
                    void* d_work, int64_t work_bytes, int32_t maxiter, double tau,
                    int32_t fail_on_maxiter, int32_t orth, double* d_hist,
                    psb_solve_result* result, void* stream);

/* -------------------------------------------------------- AMG V-cycle -- */
#define PSB_SMOOTH_JACOBI 0   /* x += omega D^-1 (f - A x), ClassicSmoothers.py:10-16 (omega=1) */
#define PSB_SMOOTH_GS     1   /* x += triu(A)^-1 (f - A x), ClassicSmoothers.py:28-36          */

/* Solve phase of the smoothed-aggregation V-cycle on a hierarchy built on the host
 * (reference setup) and already uploaded: A[l] level matrices (0 = coarsest), P[l]
 * level l -> l+1, R[l] level l+1 -> l (whatever MLHierarchy.update/downdate return).
 * d_dinv[l] (Jacobi) / gsU[l] (Gauss-Seidel: psb_trsv of triu(A[l])) for l >= 1.
 * `coarse`: exact LU of A[0] as a psb_ilu preconditioner (factored once on the host).
 * The handle is a preconditioner: psb_prec_apply runs n_iters V-cycles from x0 = b with
 * the early exit ||r|| < tau ||b|| (AMGPreconditioner.py:39-51, VCycleSolver.py:68-91).
 * All handles passed in stay owned by the caller. */
int psb_amg_create(int32_t n_levels, const psb_csr_t* A, const psb_csr_t* P,
                   const psb_csr_t* R, const double* const* d_dinv,
                   const psb_trsv_t* gsU, psb_prec_t coarse, int32_t smoother,
                   double omega, int32_t nu_pre, int32_t nu_post, int32_t n_iters,
                   double tau, psb_prec_t* out);
/* AMGVCycleSolver.solve (VCycleSolver.py:52-95): up to maxiter cycles, d_hist[k] =
 * ||b - A x_k||; status PSB_CONVERGED / PSB_MAXITER / PSB_TRIVIAL.  Synchronises. */
/* x += dx (ClassicSmoothers.py:34: the update of the Gauss-Seidel smoother when it is used as a
 * stand-alone plug-in object; inside the V-cycle the same kernel runs from psb_amg_create's handle) */
int psb_vec_add(int64_t n, const double* d_dx, double* d_x, void* stream);
int psb_amg_solve(psb_prec_t amg, const double* d_b, double* d_x, int32_t maxiter,
                  double tau, double* d_hist, psb_solve_result* result, void* stream);

/* ------------------------------------- Bratu residual / Jacobian on the device -- */
/* F = Au - alpha exp(-u)   (examples/FDBratu2D.py:20-21; Au from psb_spmv) */
int psb_bratu_residual(int64_t n, const double* d_Au, const double* d_u, double alpha,
                       double* d_F, void* stream);
/* vals[diag_pos[i]] = a_diag[i] + alpha exp(-u[i]): the Jacobian of FDBratu2D.py:23-29
 * written into the VALUES of a device CSR whose structure is A's. */
int psb_bratu_jacobian(int64_t n, const int64_t* d_diag_pos, const double* d_a_diag,
                       const double* d_u, double alpha, double* d_vals, void* stream);

/* ------------------------------------------------- multi-GPU (one rank per GPU) -- */
/* The reference has no distributed code; contract: SURVEY.md section 8e.  NCCL is
 * taken from the libnccl.so.2 already loaded in the process (torch's). */
int psb_nccl_unique_id(void* h_id128);                 /* rank 0; 128-byte HOST buffer  */
int psb_comm_create(const void* h_id128, int32_t rank, int32_t nranks, psb_comm_t* out);
int psb_comm_destroy(psb_comm_t comm);
int psb_comm_allreduce_sum(psb_comm_t comm, double* d_buf, int64_t count, void* stream);

/* Plan for this rank's row block.  A_local: n_loc x (n_loc + n_halo) CSR with local column
 * numbering (owned columns first, halo columns after, in sorted-global order).  Rows
 * [r0, r1) reference no halo column (r0, r1 multiples of 4) and are multiplied while the
 * halo is in flight.  Peer i: send `send_cnt[i]` entries to rank peer_rank[i] -- either
 * the contiguous slice starting at local offset send_off[i] (d_send_idx == NULL or
 * d_send_idx[i] == NULL) or gathered through the device index list d_send_idx[i] -- and
 * receive recv_cnt[i] entries into halo positions [recv_off[i], recv_off[i]+recv_cnt[i]). */
int psb_dist_create(psb_comm_t comm, psb_csr_t A_local, int64_t n_loc, int64_t n_halo,
                    int64_t r0, int64_t r1, int32_t n_peers, const int32_t* h_peer_rank,
                    const int64_t* h_send_off, const int64_t* h_send_cnt,
                    const int32_t* const* d_send_idx, const int64_t* h_recv_off,
                    const int64_t* h_recv_cnt, psb_dist_t* out);
int psb_dist_destroy(psb_dist_t D);
/* NVLink peer-memory mode (optional; without it the solve uses NCCL collectives).  Each rank
 * allocates a region holding its two p buffers, reduction slots and halo flags and exports
 * it (cudaIpc handle, 64 bytes; layout = {p buffer 0, p buffer 1, r buffer byte offsets,
 * n_loc, n_halo, 0}); after the handles have been exchanged every rank maps its peers' regions.  Push
 * i: the contiguous slice [send_off, send_off+send_cnt) of p goes to rank push_rank[i] at
 * byte offset remote_off{0,1}[i] of its region (one per p buffer; remote_off_r[i] for the r
 * halo used by the single-launch persistent solve), and its halo flag number
 * remote_flag_index[i] is raised.  Then psb_dist_pcg_solve fuses the collectives into the
 * compute kernels: scalar all-reduces and halo exchange are peer stores + local polling. */
int psb_dist_p2p_alloc(psb_dist_t D, void* h_handle64, int64_t layout[6]);
int psb_dist_p2p_open(psb_dist_t D, const void* h_handles, int32_t n_push,
                      const int32_t* h_push_rank, const int64_t* h_send_off,
                      const int64_t* h_send_cnt, const int64_t* h_remote_off0,
                      const int64_t* h_remote_off1, const int64_t* h_remote_off_r,
                      const int32_t* h_remote_flag_index);
/* y_loc = A_loc [x_loc | halo]; d_x_ext has n_loc + n_halo entries, the halo part is
 * filled by the exchange (NCCL send/recv on a side stream, overlapped with interior rows). */
int psb_dist_spmv(psb_dist_t D, double* d_x_ext, double* d_y, void* stream);
int64_t psb_dist_pcg_workspace_bytes(int64_t n_loc, int64_t n_halo);
/* Un-preconditioned PCG on the row-partitioned system; same loop, result codes and history
 * as psb_pcg_solve, with p.Ap and r.r all-reduced (2 NCCL all-reduces per iteration). */
int psb_dist_pcg_solve(psb_dist_t D, const double* d_b_loc, double* d_x_loc, void* d_work,
                       int64_t work_bytes, int32_t maxiter, double tau, int32_t fail_on_maxiter,
                       double* d_hist, psb_solve_result* result, void* stream);


/* ---- row-partitioned GMRES and AMG V-cycle (the reference is single-process:
 * GMRESSolver.py:104-125 and VCycleManager.py:31-62 are the loops being sharded) ------------------
 * psb_dist_create also accepts RECTANGULAR row blocks (restriction / prolongation): the local
 * matrix is n_loc x (n_own + n_halo), n_own = owned entries of the operator's INPUT vector.
 *
 * psb_dist_amg_create: level l >= 1 (n_levels - 1 = finest): A[l] = this rank's row block of A_l,
 * d_dinv[l] = reciprocal diagonal of those rows; R[l-1] = its block of the restriction to level l-1;
 * P[l-1] (l-1 >= 1) = its block of the prolongator from level l-1; P0 = its rows of the prolongator
 * from the COARSEST level as a plain local CSR with GLOBAL column indices -- the coarsest vector is
 * replicated: the ranks gather the coarse right-hand side (h_starts0[nranks+1] = its row
 * partition) and each runs `coarse`, an exact solver of the whole coarsest system.  Damped-Jacobi
 * smoothing (Gauss-Seidel has a global dependency chain: single GPU only).  The handle works with
 * psb_prec_apply (slices in, slices out), psb_amg_solve and psb_dist_gmres_solve. */
int psb_dist_amg_create(psb_comm_t comm, int32_t n_levels, const psb_dist_t* A, const psb_dist_t* P,
                        const psb_dist_t* R, psb_csr_t P0, const double* const* d_dinv,
                        psb_prec_t coarse, const int64_t* h_starts0, double omega, int32_t nu_pre,
                        int32_t nu_post, int32_t n_iters, double tau, psb_prec_t* out);
int64_t psb_dist_gmres_workspace_bytes(int64_t n_loc, int64_t n_halo, int32_t maxiter);
/* psb_gmres_solve on this rank's row block D: every reduction (||b||, the batched Gram-Schmidt
 * dots, ||w||^2, the true residual) is all-reduced, the SpMV exchanges its halo. */
int psb_dist_gmres_solve(psb_dist_t D, psb_prec_t prec, const double* d_b_loc, double* d_x_loc,
                         void* d_work, int64_t work_bytes, int32_t maxiter, double tau,
                         int32_t fail_on_maxiter, int32_t orth, double* d_hist,
                         psb_solve_result* result, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PYSOLV_B200_H */
