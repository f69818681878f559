// AMG V-cycle SOLVE phase on the device (the hierarchy is built on the host by the
// reference's smoothed-aggregation setup and uploaded once).
//
// Replaces, kernel for kernel:
//   AMGPreconditioner.apply      PySolvers/Linear/AMGPreconditioner.py:46-51
//   AMGVCycleSolver.solve        PySolvers/Linear/VCycleSolver.py:52-95  (x0 = b, strict '<' test)
//   VCycleManager.runLevel       PySolvers/Linear/VCycleManager.py:31-62
//   JacobiSmoother.apply         PySolvers/Linear/ClassicSmoothers.py:10-16   -> fused sweep
//                                x_new = x + omega D^-1 (f - A x), one SpMV-shaped pass
//   GaussSeidelSmoother.apply    ClassicSmoothers.py:28-36  (x += triu(A)^-1 (f - A x))
//                                -> residual SpMV + sync-free SpTRSV + axpy
//   coarsest level spsolve       VCycleManager.py:34-37 -> LU factored ONCE on the host (splu),
//                                applied with two SpTRSVs (the reference re-factorises per cycle)
// Residual / restriction / prolongation are SpMV epilogues (spmv.cu): r = f - A x, f_c = R r,
// x += P e.  The per-cycle residual norm, the '<' test and the "stop cycling" flag stay on
// the device; the kernels of later cycles see the flag and return.
#include "prec.cuh"
#include "spmv.cuh"
#include "sptrsv.cuh"

#include <algorithm>
#include <new>
#include <vector>

namespace psb {

__global__ void __launch_bounds__(kBlock)
amg_copy_kernel(double* __restrict__ dst, const double* __restrict__ src, int64_t n, const int* d_skip) {
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    dst[i] = src[i];
}

__global__ void __launch_bounds__(kBlock)
amg_zero_kernel(double* __restrict__ dst, int64_t n, const int* d_skip) {
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    dst[i] = 0.0;
}

// x += dx   (ClassicSmoothers.py:34)
__global__ void __launch_bounds__(kBlock)
amg_add_kernel(double* __restrict__ x, const double* __restrict__ dx, int64_t n, const int* d_skip) {
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    x[i] = x[i] + dx[i];
}

// start of a solve: skip = outer flag, x = b, ||b||
__global__ void __launch_bounds__(kBlock)
amg_begin_kernel(AmgState* st, const int* outer_skip, int64_t n, const double* __restrict__ b,
                 double* __restrict__ x, double tau, int maxiter, ReduceBuf rb) {
  __shared__ double scratch[kWarps];
  const int outer = outer_skip ? ld_cg(outer_skip) : 0;
  double acc = 0.0;
  if (!outer) {
    for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
      const double v = b[i];
      x[i] = v;                                             // x0 = b   (VCycleSolver.py:69)
      acc += v * v;
    }
  }
  double t = block_sum(acc, scratch);
  if (threadIdx.x == 0) rb.partials[blockIdx.x] = t;
  if (last_block(rb.ticket)) {
    double bb = sum_partials(rb.partials, gridDim.x, scratch);
    if (threadIdx.x == 0) {
      st->tau = tau; st->maxiter = maxiter; st->cycles = 0; st->norm_r = 0.0;
      st->norm_b = sqrt(bb);
      st->status = PSB_MAXITER;
      st->skip = outer;
      if (!outer && bb == 0.0) { st->skip = 1; st->status = PSB_TRIVIAL; }   // x = b = 0
    }
  }
}

// row-partitioned solve: the same two steps with the all-reduce of the local sums in between
__global__ void __launch_bounds__(kBlock)
amg_begin_local_kernel(AmgState* st, const int* outer_skip, int64_t n, const double* __restrict__ b,
                       double* __restrict__ x, ReduceBuf rb) {
  __shared__ double scratch[kWarps];
  const int outer = outer_skip ? ld_cg(outer_skip) : 0;
  double acc = 0.0;
  if (!outer) {
    for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
      const double v = b[i];
      x[i] = v;                                             // x0 = b   (VCycleSolver.py:69)
      acc += v * v;
    }
  }
  double t = block_sum(acc, scratch);
  if (threadIdx.x == 0) rb.partials[blockIdx.x] = t;
  if (last_block(rb.ticket)) {
    double bb = sum_partials(rb.partials, gridDim.x, scratch);
    if (threadIdx.x == 0) { st->bb = bb; st->skip = outer; }
  }
}

__global__ void amg_begin_finish_kernel(AmgState* st, double tau, int maxiter) {
  if (threadIdx.x != 0) return;
  const double bb = st->bb;                                 // summed over the ranks
  st->tau = tau; st->maxiter = maxiter; st->cycles = 0; st->norm_r = 0.0;
  st->norm_b = sqrt(bb);
  st->status = PSB_MAXITER;
  if (!st->skip && bb == 0.0) { st->skip = 1; st->status = PSB_TRIVIAL; }
}

__global__ void amg_cycle_finish_kernel(AmgState* st, double* hist) {   // VCycleSolver.py:87-91
  if (threadIdx.x != 0 || st->skip != 0) return;
  const double nr = sqrt(st->rr);
  const int k = st->cycles;
  st->norm_r = nr;
  if (hist != nullptr) hist[k] = nr;
  st->cycles = k + 1;
  if (nr < st->tau * st->norm_b) { st->status = PSB_CONVERGED; st->skip = 1; }
  else if (k + 1 >= st->maxiter) { st->skip = 1; }
}

// An operator of the hierarchy: the whole matrix (one GPU) or this rank's row block with its halo
// plan (row-partitioned: the input vector must then be an EXTENDED buffer, owned entries
// followed by room for the halo, which the exchange fills in place).
struct AmgOp {
  psb_csr* local = nullptr;
  psb_dist* dist = nullptr;
  bool present() const { return local != nullptr || dist != nullptr; }
  int launch(Epi epi, double* x, double* y, const EpiArgs& ea, const int* skip, cudaStream_t s) const {
    return dist ? dist_spmv_epi(dist, epi, x, y, ea, skip, s) : spmv_launch(local, epi, x, y, ea, skip, s);
  }
  int64_t rows() const { return dist ? dist_n_loc(dist) : local->n_rows; }
  int64_t x_len() const { return dist ? dist_n_own(dist) + dist_n_halo(dist) : local->n_cols; }
};

struct AmgLevel {
  AmgOp A;
  AmgOp P;                     // this level -> next finer level
  AmgOp R;                     // next finer level -> this level
  const double* dinv = nullptr;
  psb_trsv* gsU = nullptr;
  int64_t n = 0;
  double *x = nullptr, *x2 = nullptr, *f = nullptr, *r = nullptr;   // owned work vectors
};

struct AmgPrec : psb_prec {
  std::vector<AmgLevel> lev;           // 0 = coarsest
  psb_prec* coarse = nullptr;          // exact LU of A_0 (not owned)
  int smoother = PSB_SMOOTH_JACOBI;
  double omega = 1.0;
  int nu_pre = 2, nu_post = 2, n_iters = 5;
  double tau = 1.0e-8;
  AmgState* st = nullptr;              // owned
  ReduceBuf rb;                        // owned
  double* hist_scratch = nullptr;      // owned
  // row-partitioned (psb_dist_amg_create): the coarsest system is solved REPLICATED -- its
  // right-hand side is gathered from the ranks, every rank solves, and the prolongator block
  // reads the full coarse vector through global column numbers (no exchange)
  psb_comm* comm = nullptr;
  std::vector<int64_t> starts0;        // coarsest-level row partition [nranks + 1]

  ~AmgPrec() override {
    for (auto& l : lev) { cudaFree(l.x); cudaFree(l.x2); cudaFree(l.f); cudaFree(l.r); }
    cudaFree(st); cudaFree(rb.partials); cudaFree(rb.ticket); cudaFree(hist_scratch);
  }
  const char* kind() const override { return "amg"; }
  int check_error() override {
    int bad = coarse ? coarse->check_error() : 0;
    for (auto& l : lev) {
      if (l.gsU) {
        bad |= trsv_take_error(l.gsU);
      }
    }
    return bad;
  }

  // nu sweeps at level l on (f, x): returns the buffer that holds the result
  int smooth(int l, const double* f, double*& x, double*& spare, int nu, cudaStream_t s) {
    AmgLevel& L = lev[l];
    const int* skip = &st->skip;
    for (int i = 0; i < nu; ++i) {
      if (smoother == PSB_SMOOTH_JACOBI) {
        EpiArgs ea; ea.f = f; ea.dinv = L.dinv; ea.omega = omega;
        int rc = L.A.launch(EPI_JACOBI, x, spare, ea, skip, s);
        if (rc != PSB_OK) return rc;
        std::swap(x, spare);
      } else {
        EpiArgs ea; ea.f = f;
        int rc = L.A.launch(EPI_RESID, x, L.r, ea, skip, s);                // r = f - A x
        if (rc != PSB_OK) return rc;
        rc = trsv_solve(L.gsU, L.r, spare, nullptr, nullptr, nullptr, skip, s);   // dx = U^-1 r
        if (rc != PSB_OK) return rc;
        amg_add_kernel<<<stream_grid(L.n, rb.max_grid), kBlock, 0, s>>>(x, spare, L.n, skip);
        PSB_LAUNCH_CHECK();
      }
    }
    return PSB_OK;
  }

  // V-cycle from level l: x (in/out) may end up in either of the two buffers; `x`/`spare`
  // are updated so that x names the result
  // `pre_done`: sweeps of the pre-smoothing that the caller has already applied to x
  int run_level(int l, const double* f, double*& x, double*& spare, cudaStream_t s, int pre_done = 0) {
    const int* skip = &st->skip;
    AmgLevel& L = lev[l];
    if (l == 0) {                                                             // VCycleManager.py:34-37
      if (comm != nullptr) {          // f, x are the FULL coarse vectors: gather the slices, solve everywhere
        int rc = dist_allgather_slices(comm, const_cast<double*>(f), starts0.data(), s);
        if (rc != PSB_OK) return rc;
      }
      return coarse->apply(f, x, skip, s);
    }
    int rc = smooth(l, f, x, spare, nu_pre - pre_done, s);                    // :42
    if (rc != PSB_OK) return rc;
    EpiArgs ea; ea.f = f;
    rc = L.A.launch(EPI_RESID, x, L.r, ea, skip, s);                          // :45
    if (rc != PSB_OK) return rc;
    AmgLevel& C = lev[l - 1];
    // row-partitioned, coarsest level: C.f is the full vector, this rank's rows start at starts0[rank]
    double* cf_rows = (comm != nullptr && l - 1 == 0) ? C.f + starts0[dist_rank(comm)] : C.f;
    rc = C.R.launch(EPI_STORE, L.r, cf_rows, EpiArgs(), skip, s);             // :48
    if (rc != PSB_OK) return rc;
    double* cx = C.x;
    double* cs = C.x2;
    if (l - 1 > 0) {
      amg_zero_kernel<<<stream_grid(C.n, rb.max_grid), kBlock, 0, s>>>(cx, C.n, skip);   // :51
      PSB_LAUNCH_CHECK();
    }
    rc = run_level(l - 1, C.f, cx, cs, s);                                    // :52
    if (rc != PSB_OK) return rc;
    rc = C.P.launch(EPI_ADD, cx, x, EpiArgs(), skip, s);                      // :55
    if (rc != PSB_OK) return rc;
    return smooth(l, f, x, spare, nu_post, s);                                // :60
  }

  // maxiter V-cycles on (b -> x_out); hist nullable.  `final_residual`: also evaluate ||b - A x||
  // after the LAST cycle (the solver reports it; as a preconditioner nothing depends on it --
  // AMGPreconditioner.py:46-51 returns x whether or not the last test passes, failOnMaxiter=False).
  // Row-partitioned (comm != nullptr): b, x_out are this rank's slices; the iterate lives in the
  // level's own EXTENDED buffers (room for the halo) and is copied to x_out at the end.
  int solve(const double* b, double* x_out, int maxiter, double tol, double* hist,
            const int* outer_skip, bool final_residual, cudaStream_t s) {
    const int top = (int)lev.size() - 1;
    AmgLevel& F = lev[top];
    const bool dist = comm != nullptr;
    const int grid = stream_grid(F.n, rb.max_grid);
    double* home = dist ? F.x : x_out;          // where the iterate lives between cycles
    if (dist) {
      amg_begin_local_kernel<<<grid, kBlock, 0, s>>>(st, outer_skip, F.n, b, home, rb);
      PSB_LAUNCH_CHECK();
      int rc = dist_allreduce(comm, &st->bb, 1, s);
      if (rc != PSB_OK) return rc;
      amg_begin_finish_kernel<<<1, 32, 0, s>>>(st, tol, maxiter);
      PSB_LAUNCH_CHECK();
    } else {
      amg_begin_kernel<<<grid, kBlock, 0, s>>>(st, outer_skip, F.n, b, home, tol, maxiter, rb);
      PSB_LAUNCH_CHECK();
    }
    const int* skip = &st->skip;
    // The residual that ends cycle k and the first Jacobi sweep of cycle k + 1 need the same
    // r = b - A x: the residual kernel forms that sweep as well (into the second buffer) and the next
    // cycle starts one sweep in -- a pass over the fine matrix less per cycle, same arithmetic.
    const bool fuse_sweep = top > 0 && smoother == PSB_SMOOTH_JACOBI && nu_pre >= 1;
    bool presmoothed = false;
    for (int k = 0; k < maxiter; ++k) {
      double* x = presmoothed ? F.x2 : home;
      double* spare = presmoothed ? home : F.x2;
      int rc;
      if (top == 0) {
        // single level: the "cycle" is the direct solve
        rc = coarse->apply(b, F.x2, skip, s);
        if (rc != PSB_OK) return rc;
        x = F.x2; spare = home;
      } else {
        rc = run_level(top, b, x, spare, s, presmoothed ? 1 : 0);
        if (rc != PSB_OK) return rc;
      }
      presmoothed = false;
      if (x != home) {                        // odd number of ping-pong sweeps: bring x home
        amg_copy_kernel<<<grid, kBlock, 0, s>>>(home, x, F.n, skip);
        PSB_LAUNCH_CHECK();
      }
      if (k + 1 == maxiter && !final_residual) break;
      // r = b - A x with ||r||, the history entry and the strict '<' test done by the kernel's
      // last CTA (VCycleSolver.py:84-91): no separate pass over r
      EpiArgs ea; ea.f = b;
      if (fuse_sweep && k + 1 < maxiter) {
        ea.jac_out = F.x2; ea.dinv = F.dinv; ea.omega = omega;
        presmoothed = true;
      }
      if (dist) {
        ea.dot = &st->rr;                     // this rank's part; finished after the all-reduce
        rc = F.A.launch(EPI_RESID_NORM, home, F.r, ea, skip, s);
        if (rc != PSB_OK) return rc;
        rc = dist_allreduce(comm, &st->rr, 1, s);
        if (rc != PSB_OK) return rc;
        amg_cycle_finish_kernel<<<1, 32, 0, s>>>(st, hist);
        PSB_LAUNCH_CHECK();
      } else {
        ea.amg_state = st; ea.amg_hist = hist;
        rc = F.A.launch(EPI_RESID_NORM, home, F.r, ea, skip, s);
        if (rc != PSB_OK) return rc;
      }
    }
    if (dist) {                               // the result, whether or not the cycles were skipped
      amg_copy_kernel<<<grid, kBlock, 0, s>>>(x_out, home, F.n, outer_skip);
      PSB_LAUNCH_CHECK();
    }
    return PSB_OK;
  }

  int apply(const double* r, double* z, const int* d_skip, cudaStream_t s) override {
    return solve(r, z, n_iters, tau, nullptr, d_skip, false, s);
  }
};

}  // namespace psb

using namespace psb;

static int amg_finish_create(AmgPrec* M, psb_prec_t* out) {
  cudaError_t e = cudaSuccess;
  const int n_levels = (int)M->lev.size();
  const bool dist = M->comm != nullptr;
  for (int l = 0; l < n_levels && e == cudaSuccess; ++l) {
    AmgLevel& L = M->lev[l];
    // buffer lengths: x / x2 are read by A_l and by P_l (towards level l+1), r by R_{l-1};
    // row-partitioned these are extended buffers, and the coarsest x / f hold the FULL vectors
    int64_t len_x = L.n, len_r = L.n, len_f = L.n;
    if (dist) {
      if (l == 0) { len_x = len_f = M->starts0.back(); }
      else {
        len_x = std::max(len_x, L.A.x_len());
        if (L.P.present()) len_x = std::max(len_x, L.P.x_len());
        len_r = std::max(len_r, M->lev[l - 1].R.x_len());
      }
    }
    auto bytes = [](int64_t n) { return (size_t)std::max<int64_t>(n, 1) * sizeof(double); };
    e = cudaMalloc((void**)&L.x, bytes(len_x));
    if (e == cudaSuccess) e = cudaMemset(L.x, 0, bytes(len_x));
    if (e == cudaSuccess) e = cudaMalloc((void**)&L.x2, bytes(len_x));
    if (e == cudaSuccess) e = cudaMemset(L.x2, 0, bytes(len_x));
    if (e == cudaSuccess) e = cudaMalloc((void**)&L.f, bytes(len_f));
    if (e == cudaSuccess) e = cudaMalloc((void**)&L.r, bytes(len_r));
    if (e == cudaSuccess) e = cudaMemset(L.r, 0, bytes(len_r));
  }
  M->n = M->lev[n_levels - 1].n;
  M->rb.max_grid = sm_count() * 16;
  if (e == cudaSuccess) e = cudaMalloc((void**)&M->st, sizeof(AmgState));
  if (e == cudaSuccess) e = cudaMemset(M->st, 0, sizeof(AmgState));
  if (e == cudaSuccess) e = cudaMalloc((void**)&M->rb.partials, sizeof(double) * M->rb.max_grid);
  if (e == cudaSuccess) e = cudaMalloc((void**)&M->rb.ticket, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(M->rb.ticket, 0, sizeof(unsigned int));
  if (e != cudaSuccess) { delete M; set_error("psb_amg_create: %s", cudaGetErrorString(e)); return PSB_ERR_CUDA; }
  *out = M;
  return PSB_OK;
}

extern "C" int psb_amg_create(int32_t n_levels, const psb_csr_t* A, const psb_csr_t* P, const psb_csr_t* R,
                              const double* const* d_dinv, const psb_trsv_t* gsU, psb_prec_t coarse,
                              int32_t smoother, double omega, int32_t nu_pre, int32_t nu_post,
                              int32_t n_iters, double tau, psb_prec_t* out) {
  PSB_REQUIRE(n_levels >= 1 && A && out && coarse, PSB_ERR_ARG, "psb_amg_create: bad argument");
  PSB_REQUIRE(smoother == PSB_SMOOTH_JACOBI || smoother == PSB_SMOOTH_GS, PSB_ERR_ARG,
              "psb_amg_create: unknown smoother");
  PSB_REQUIRE(n_levels == 1 || (P && R), PSB_ERR_ARG, "psb_amg_create: transfer operators missing");
  PSB_REQUIRE(nu_pre >= 0 && nu_post >= 0 && n_iters >= 1, PSB_ERR_ARG, "psb_amg_create: bad sweep counts");
  AmgPrec* M = new (std::nothrow) AmgPrec();
  PSB_REQUIRE(M != nullptr, PSB_ERR_ARG, "psb_amg_create: out of host memory");
  M->smoother = smoother; M->omega = omega; M->nu_pre = nu_pre; M->nu_post = nu_post;
  M->n_iters = n_iters; M->tau = tau; M->coarse = coarse;
  M->lev.resize(n_levels);
  for (int l = 0; l < n_levels; ++l) {
    AmgLevel& L = M->lev[l];
    L.A.local = A[l];
    if (!A[l] || A[l]->n_rows != A[l]->n_cols) { delete M; set_error("psb_amg_create: level matrix %d invalid", l); return PSB_ERR_ARG; }
    L.n = A[l]->n_rows;
    if (l < n_levels - 1) {
      L.P.local = P[l]; L.R.local = R[l];
      if (!P[l] || !R[l] || P[l]->n_cols != L.n || R[l]->n_rows != L.n || P[l]->n_rows != A[l + 1]->n_rows ||
          R[l]->n_cols != A[l + 1]->n_rows) {
        delete M; set_error("psb_amg_create: transfer operator shapes do not match at level %d", l); return PSB_ERR_ARG;
      }
    }
    if (l > 0) {
      if (smoother == PSB_SMOOTH_JACOBI) {
        if (!d_dinv || !d_dinv[l]) { delete M; set_error("psb_amg_create: dinv missing at level %d", l); return PSB_ERR_ARG; }
        L.dinv = d_dinv[l];
      } else {
        if (!gsU || !gsU[l] || gsU[l]->n != L.n || gsU[l]->lower) {
          delete M; set_error("psb_amg_create: triu(A) factor missing at level %d", l); return PSB_ERR_ARG;
        }
        L.gsU = gsU[l];
      }
    }
  }
  if (coarse->n != M->lev[0].n) { delete M; set_error("psb_amg_create: coarse solver size mismatch"); return PSB_ERR_ARG; }
  return amg_finish_create(M, out);
}

// Row-partitioned V-cycle ("AMG with Jacobi smoothing shards like SpMV, coarse
// solve replicated").  Level l >= 1: A[l] is this rank's row block of A_l; R[l-1] its block of the
// restriction (rows: level l-1, input: level-l vector); P[l-1] its block of the prolongator (rows:
// level l, input: level l-1 vector) -- for l-1 == 0 that is a plain local CSR P0 whose columns
// are GLOBAL coarse indices, because the coarsest vector is replicated: every rank gathers the
// coarse right-hand side (h_starts0 = its row partition) and runs the same exact solve `coarse`
// (whole coarsest system).  Jacobi smoothing only: Gauss-Seidel has a global dependency chain.
extern "C" int psb_dist_amg_create(psb_comm_t comm, int32_t n_levels, const psb_dist_t* A, const psb_dist_t* P,
                                   const psb_dist_t* R, psb_csr_t P0, const double* const* d_dinv,
                                   psb_prec_t coarse, const int64_t* h_starts0, double omega, int32_t nu_pre,
                                   int32_t nu_post, int32_t n_iters, double tau, psb_prec_t* out) {
  PSB_REQUIRE(comm && n_levels >= 2 && A && R && P0 && d_dinv && coarse && h_starts0 && out, PSB_ERR_ARG,
              "psb_dist_amg_create: bad argument (needs at least two levels)");
  PSB_REQUIRE(n_levels == 2 || P, PSB_ERR_ARG, "psb_dist_amg_create: prolongator blocks missing");
  PSB_REQUIRE(nu_pre >= 0 && nu_post >= 0 && n_iters >= 1, PSB_ERR_ARG, "psb_dist_amg_create: bad sweep counts");
  AmgPrec* M = new (std::nothrow) AmgPrec();
  PSB_REQUIRE(M != nullptr, PSB_ERR_ARG, "psb_dist_amg_create: out of host memory");
  M->smoother = PSB_SMOOTH_JACOBI; M->omega = omega; M->nu_pre = nu_pre; M->nu_post = nu_post;
  M->n_iters = n_iters; M->tau = tau; M->coarse = coarse; M->comm = comm;
  const int nr = dist_nranks(comm), me = dist_rank(comm);
  M->starts0.assign(h_starts0, h_starts0 + nr + 1);
  M->lev.resize(n_levels);
  M->lev[0].n = M->starts0[me + 1] - M->starts0[me];
  for (int l = 1; l < n_levels; ++l) {
    AmgLevel& L = M->lev[l];
    if (!A[l] || !R[l - 1] || !d_dinv[l] || dist_n_own(A[l]) != dist_n_loc(A[l])) {
      delete M; set_error("psb_dist_amg_create: operator missing or not square at level %d", l); return PSB_ERR_ARG;
    }
    L.A.dist = A[l];
    L.n = dist_n_loc(A[l]);
    L.dinv = d_dinv[l];
    AmgLevel& C = M->lev[l - 1];
    C.R.dist = R[l - 1];
    if (l - 1 == 0) C.P.local = P0; else C.P.dist = P[l - 1];
    const int64_t c_rows = dist_n_loc(R[l - 1]);
    const bool ok = c_rows == C.n && dist_n_own(R[l - 1]) == L.n &&
                    (l - 1 == 0 ? (P0->n_rows == L.n && P0->n_cols == M->starts0[nr])
                                : (P[l - 1] && dist_n_loc(P[l - 1]) == L.n && dist_n_own(P[l - 1]) == C.n));
    if (!ok) { delete M; set_error("psb_dist_amg_create: transfer operator shapes do not match at level %d", l - 1); return PSB_ERR_ARG; }
  }
  if (coarse->n != M->starts0[nr]) { delete M; set_error("psb_dist_amg_create: the coarse solver must cover the whole coarsest system"); return PSB_ERR_ARG; }
  return amg_finish_create(M, out);
}

// x += dx on the device (ClassicSmoothers.py:34, the Gauss-Seidel plug-in smoother's update)
extern "C" int psb_vec_add(int64_t n, const double* d_dx, double* d_x, void* stream) {
  PSB_REQUIRE(n >= 0 && (n == 0 || (d_dx && d_x)), PSB_ERR_ARG, "psb_vec_add: bad argument");
  if (n == 0) return PSB_OK;
  amg_add_kernel<<<stream_grid(n, sm_count() * 16), kBlock, 0, (cudaStream_t)stream>>>(d_x, d_dx, n, nullptr);
  PSB_LAUNCH_CHECK();
  return PSB_OK;
}

// AMGVCycleSolver.solve: up to maxiter V-cycles from x0 = b; d_hist (maxiter doubles) receives
// ||b - A x|| after every cycle.  Synchronises.
extern "C" int psb_amg_solve(psb_prec_t amg, const double* d_b, double* d_x, int32_t maxiter, double tau,
                             double* d_hist, psb_solve_result* result, void* stream) {
  PSB_REQUIRE(amg && d_b && d_x && result && maxiter >= 1, PSB_ERR_ARG, "psb_amg_solve: bad argument");
  PSB_REQUIRE(strcmp(amg->kind(), "amg") == 0, PSB_ERR_ARG, "psb_amg_solve: not an AMG handle");
  PSB_REQUIRE(d_b != d_x, PSB_ERR_ARG, "psb_amg_solve: x must not alias b");
  AmgPrec* M = static_cast<AmgPrec*>(amg);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = M->solve(d_b, d_x, maxiter, tau, d_hist, nullptr, true, st);
  if (rc != PSB_OK) return rc;
  PSB_CUDA(cudaStreamSynchronize(st));
  AmgState hs;
  PSB_CUDA(cudaMemcpy(&hs, M->st, sizeof(hs), cudaMemcpyDeviceToHost));
  result->status = hs.status;
  result->k = hs.cycles - 1;
  result->n_hist = hs.cycles;
  result->lucky = 0;
  result->norm_r = hs.norm_r; result->norm_b = hs.norm_b; result->norm_r_rec = hs.norm_r;
  if (M->check_error() != 0) {
    set_error("psb_amg_solve: a triangular solve of the hierarchy reported a device-side failure");
    return PSB_ERR_CUDA;
  }
  return PSB_OK;
}
