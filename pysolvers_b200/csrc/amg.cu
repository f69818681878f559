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

struct AmgLevel {
  psb_csr* A = nullptr;
  psb_csr* P = nullptr;        // this level -> next finer level
  psb_csr* R = nullptr;        // next finer level -> this level
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

  ~AmgPrec() override {
    for (auto& l : lev) { cudaFree(l.x); cudaFree(l.x2); cudaFree(l.f); cudaFree(l.r); }
    cudaFree(st); cudaFree(rb.partials); cudaFree(rb.ticket); cudaFree(hist_scratch);
  }
  const char* kind() const override { return "amg"; }
  int check_error() override {
    int bad = coarse ? coarse->check_error() : 0;
    for (auto& l : lev) {
      if (l.gsU) {
        int f = 0;
        cudaMemcpy(&f, l.gsU->d_error, sizeof(int), cudaMemcpyDeviceToHost);
        bad |= f;
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
        int rc = spmv_launch(L.A, EPI_JACOBI, x, spare, ea, skip, s);
        if (rc != PSB_OK) return rc;
        std::swap(x, spare);
      } else {
        EpiArgs ea; ea.f = f;
        int rc = spmv_launch(L.A, EPI_RESID, x, L.r, ea, skip, s);          // r = f - A x
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
  int run_level(int l, const double* f, double*& x, double*& spare, cudaStream_t s) {
    const int* skip = &st->skip;
    AmgLevel& L = lev[l];
    if (l == 0) return coarse->apply(f, x, skip, s);                          // VCycleManager.py:34-37
    int rc = smooth(l, f, x, spare, nu_pre, s);                               // :42
    if (rc != PSB_OK) return rc;
    EpiArgs ea; ea.f = f;
    rc = spmv_launch(L.A, EPI_RESID, x, L.r, ea, skip, s);                    // :45
    if (rc != PSB_OK) return rc;
    AmgLevel& C = lev[l - 1];
    rc = spmv_launch(C.R, EPI_STORE, L.r, C.f, EpiArgs(), skip, s);           // :48
    if (rc != PSB_OK) return rc;
    double* cx = C.x;
    double* cs = C.x2;
    if (l - 1 > 0) {
      amg_zero_kernel<<<stream_grid(C.n, rb.max_grid), kBlock, 0, s>>>(cx, C.n, skip);   // :51
      PSB_LAUNCH_CHECK();
    }
    rc = run_level(l - 1, C.f, cx, cs, s);                                    // :52
    if (rc != PSB_OK) return rc;
    rc = spmv_launch(C.P, EPI_ADD, cx, x, EpiArgs(), skip, s);                // :55
    if (rc != PSB_OK) return rc;
    return smooth(l, f, x, spare, nu_post, s);                                // :60
  }

  // maxiter V-cycles on (b -> x_out); hist nullable.  `final_residual`: also evaluate ||b - A x||
  // after the LAST cycle (the solver reports it; as a preconditioner nothing depends on it --
  // AMGPreconditioner.py:46-51 returns x whether or not the last test passes, failOnMaxiter=False).
  int solve(const double* b, double* x_out, int maxiter, double tol, double* hist,
            const int* outer_skip, bool final_residual, cudaStream_t s) {
    const int top = (int)lev.size() - 1;
    AmgLevel& F = lev[top];
    const int grid = stream_grid(F.n, rb.max_grid);
    amg_begin_kernel<<<grid, kBlock, 0, s>>>(st, outer_skip, F.n, b, x_out, tol, maxiter, rb);
    PSB_LAUNCH_CHECK();
    const int* skip = &st->skip;
    for (int k = 0; k < maxiter; ++k) {
      double* x = x_out;
      double* spare = F.x2;
      int rc;
      if (top == 0) {
        // single level: the "cycle" is the direct solve
        rc = coarse->apply(b, F.x2, skip, s);
        if (rc != PSB_OK) return rc;
        x = F.x2; spare = x_out;
      } else {
        rc = run_level(top, b, x, spare, s);
        if (rc != PSB_OK) return rc;
      }
      if (x != x_out) {                       // odd number of ping-pong sweeps: bring x home
        amg_copy_kernel<<<grid, kBlock, 0, s>>>(x_out, x, F.n, skip);
        PSB_LAUNCH_CHECK();
      }
      if (k + 1 == maxiter && !final_residual) break;
      // r = b - A x with ||r||, the history entry and the strict '<' test done by the kernel's
      // last CTA (VCycleSolver.py:84-91): no separate pass over r
      EpiArgs ea; ea.f = b; ea.amg_state = st; ea.amg_hist = hist;
      rc = spmv_launch(F.A, EPI_RESID_NORM, x_out, F.r, ea, skip, s);
      if (rc != PSB_OK) return rc;
    }
    return PSB_OK;
  }

  int apply(const double* r, double* z, const int* d_skip, cudaStream_t s) override {
    return solve(r, z, n_iters, tau, nullptr, d_skip, false, s);
  }
};

}  // namespace psb

using namespace psb;

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
  cudaError_t e = cudaSuccess;
  for (int l = 0; l < n_levels && e == cudaSuccess; ++l) {
    AmgLevel& L = M->lev[l];
    L.A = A[l];
    if (!L.A || L.A->n_rows != L.A->n_cols) { delete M; set_error("psb_amg_create: level matrix %d invalid", l); return PSB_ERR_ARG; }
    L.n = L.A->n_rows;
    if (l < n_levels - 1) {
      L.P = P[l]; L.R = R[l];
      if (!L.P || !L.R || L.P->n_cols != L.n || L.R->n_rows != L.n || L.P->n_rows != A[l + 1]->n_rows ||
          L.R->n_cols != A[l + 1]->n_rows) {
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
    const size_t bytes = (size_t)std::max<int64_t>(L.n, 1) * sizeof(double);
    e = cudaMalloc((void**)&L.x, bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&L.x2, bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&L.f, bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&L.r, bytes);
  }
  M->n = M->lev[n_levels - 1].n;
  if (coarse->n != M->lev[0].n) { delete M; set_error("psb_amg_create: coarse solver size mismatch"); return PSB_ERR_ARG; }
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
