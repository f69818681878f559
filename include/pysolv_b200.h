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
typedef struct psb_comm*  psb_comm_t;   /* NCCL communicator + halo plan        */

/* ------------------------------------------------------------------ misc -- */
int         psb_version(void);
const char* psb_last_error(void);
/* kernels launched by this library since load (bench.py's gpu_launches). */
long long   psb_launch_count(void);

/* ------------------------------------------------------------------- CSR -- */
/* SpMV kernel kinds chosen from the row-length histogram at create time. */
#define PSB_SPMV_STREAM  1  /* CTA streams a contiguous nnz chunk through smem;
                               one thread per row sums in STORED order (bit-equal
                               to scipy csr_matvec)                              */
#define PSB_SPMV_VECTOR  2  /* sub-warp per row, shuffle reduction              */

/* Wraps caller-owned device arrays (no copy).  Synchronises `stream` once to
 * read back the row-length statistics that pick the kernel.
 * Replaces: the scipy csr_matrix operand of mvmult(),
 * PySolvers/Linear/IterativeLinearSolver.py:94-106. */
int psb_csr_create(int64_t n_rows, int64_t n_cols, int64_t nnz,
                   const int32_t* d_rowptr, const int32_t* d_colind,
                   const double* d_vals, void* stream, psb_csr_t* out);
int psb_csr_destroy(psb_csr_t A);
/* info[0]=kernel kind, [1]=max row length, [2]=max nnz per 256-row tile,
 * [3]=rows per tile, [4]=vector width (VECTOR kind), [5]=grid size. */
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

#ifdef __cplusplus
}
#endif
#endif /* PYSOLV_B200_H */
