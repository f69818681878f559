// Exact sparse LU solve with a DENSE, explicitly inverted trailing block:
//     x = Pc U^-1 L^-1 Pr v,      L = [L11 0; L21 L22],   U = [U11 U12; 0 U22].
//
// Replaces the coarsest-level solve of the AMG V-cycle, spsolve(A_c, f) in
// PySolvers/Linear/VCycleManager.py:34-37 (a full SuperLU factorisation + gstrs per cycle in the
// reference; here the factors are computed once on the host and uploaded).
//
// Why: with a fill-reducing ordering the last few thousand rows of L and U are the top separators
// of the elimination tree: an almost dense triangle in which every row is a dependency level of
// its own (Bratu 512^2: 1 830 levels, of which the trailing 2 048 rows span 1 598).  A sparse
// triangular solve is bound by levels x hand-over latency (profiles/round1d_trsv.md), so that
// tail is the whole cost.  The trailing diagonal blocks are small and very well conditioned
// (cond ~ 20 - 60, measured), hence they are inverted once (setup) and applied as dense triangular
// matrix-vector products, which are plain HBM streaming:
//     y1 = L11^-1 w1               sparse triangular solve on the leading block (few levels)
//     y2 = inv(L22) (w2 - L21 y1)  SpMV + triangular GEMV
//     x2 = inv(U22) y2             triangular GEMV
//     x1 = U11^-1 (y1 - U12 x2)    SpMV + sparse triangular solve
// Agreement with SuperLU.solve: 5e-16 relative (tests/test_gpu_amg.py).
#include "prec.cuh"
#include "spmv.cuh"
#include "sptrsv.cuh"

#include <algorithm>
#include <new>
#include <vector>

namespace psb {

namespace {

constexpr int kGemvWarps = 8;

// y = M x with M dense row-major n x n, lower (j <= i) or upper (j >= i) triangular: only the
// triangle is read.  One warp per row, rows dealt so that long and short rows alternate; each
// lane sums its strided share in index order, then a fixed shuffle tree: deterministic.
template <bool kLower>
__global__ void __launch_bounds__(kGemvWarps * 32)
tri_gemv_kernel(const double* __restrict__ M, int n, const double* __restrict__ x,
                double* __restrict__ y, const int* d_skip) {
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * kGemvWarps + (threadIdx.x >> 5);
  const int nw = gridDim.x * kGemvWarps;
  for (int t = w; t < n; t += nw) {
    // pair a long row with a short one: t even -> from the long end, t odd -> from the short end
    const int i = (t & 1) ? (kLower ? (t >> 1) : n - 1 - (t >> 1)) : (kLower ? n - 1 - (t >> 1) : (t >> 1));
    const int j0 = kLower ? 0 : i, j1 = kLower ? i + 1 : n;
    const double* row = M + (int64_t)i * n;
    double acc = 0.0;
    // aligned body with 128-bit loads (rows start 16-byte aligned when n is even)
    int j = j0 + lane * 2;
    if ((n & 1) == 0) {
      const int ja = (j0 + 1) & ~1;                 // first even index >= j0
      if (lane == 0 && ja > j0) acc = row[j0] * x[j0];
      for (j = ja + lane * 2; j + 1 < j1; j += 64) {
        const double2 m2 = ld_stream2(row + j);
        acc += m2.x * x[j];
        acc += m2.y * x[j + 1];
      }
      if (j < j1) acc += row[j] * x[j];
    } else {
      for (j = j0 + lane; j < j1; j += 32) acc += row[j] * x[j];
    }
    acc = warp_sum(acc);
    if (lane == 0) y[i] = acc;
  }
}

__global__ void __launch_bounds__(kBlock)
gather_kernel(const double* __restrict__ v, const int32_t* __restrict__ map, int64_t off, int64_t cnt,
              double* __restrict__ out, const int* d_skip) {
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * kBlock)
    out[i] = v[map[off + i]];
}

__global__ void __launch_bounds__(kBlock)
scatter_kernel(const double* __restrict__ v, const int32_t* __restrict__ map, int64_t off, int64_t cnt,
               double* __restrict__ out, const int* d_skip) {
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * kBlock)
    out[map[off + i]] = v[i];
}

struct SplitLuPrec : psb_prec {
  int64_t n1 = 0, n2 = 0;
  psb_trsv* L11 = nullptr;      // not owned (null when n1 == 0)
  psb_trsv* U11 = nullptr;
  const psb_csr* L21 = nullptr; // n2 x n1, not owned
  const psb_csr* U12 = nullptr; // n1 x n2
  const double* invL22 = nullptr;   // n2 x n2 row-major, not owned
  const double* invU22 = nullptr;
  int32_t* iperm_r = nullptr;   // owned: row r of the L solve takes v[iperm_r[r]]
  int32_t* iperm_c = nullptr;   // owned: result[iperm_c[r]] = z[r]
  double* buf = nullptr;        // owned: y1 (n1) | s1 (n1) | w2 (n2) | t2 (n2) | y2 (n2)
  ~SplitLuPrec() override { cudaFree(iperm_r); cudaFree(iperm_c); cudaFree(buf); }
  const char* kind() const override { return "splitlu"; }
  int check_error() override {
    int a = 0, b = 0;
    if (L11) cudaMemcpy(&a, L11->d_error, sizeof(int), cudaMemcpyDeviceToHost);
    if (U11) cudaMemcpy(&b, U11->d_error, sizeof(int), cudaMemcpyDeviceToHost);
    return a | b;
  }
  int apply(const double* r, double* z, const int* d_skip, cudaStream_t st) override {
    double* y1 = buf;
    double* s1 = buf + n1;
    double* w2 = buf + 2 * n1;
    double* t2 = w2 + n2;
    double* y2 = t2 + n2;
    const int g2 = (int)std::max<int64_t>(1, std::min<int64_t>((n2 + kBlock - 1) / kBlock, (int64_t)sm_count() * 4));
    const int gg = (int)std::max<int64_t>(1, std::min<int64_t>((n2 + kGemvWarps - 1) / kGemvWarps, (int64_t)sm_count() * 8));
    int rc = PSB_OK;
    if (n1 > 0) {
      rc = trsv_solve(L11, r, y1, iperm_r, nullptr, nullptr, d_skip, st);          // y1 = L11^-1 (Pr v)_1
      if (rc != PSB_OK) return rc;
    }
    gather_kernel<<<g2, kBlock, 0, st>>>(r, iperm_r, n1, n2, w2, d_skip);            // w2 = (Pr v)_2
    PSB_LAUNCH_CHECK();
    const double* rhs2 = w2;
    if (n1 > 0) {
      EpiArgs ea; ea.f = w2;
      rc = spmv_launch(L21, EPI_RESID, y1, t2, ea, d_skip, st);                      // t2 = w2 - L21 y1
      if (rc != PSB_OK) return rc;
      rhs2 = t2;
    }
    tri_gemv_kernel<true><<<gg, kGemvWarps * 32, 0, st>>>(invL22, (int)n2, rhs2, y2, d_skip);   // y2 = L22^-1 .
    PSB_LAUNCH_CHECK();
    double* x2 = w2;                                                                 // w2 is free again
    tri_gemv_kernel<false><<<gg, kGemvWarps * 32, 0, st>>>(invU22, (int)n2, y2, x2, d_skip);    // x2 = U22^-1 y2
    PSB_LAUNCH_CHECK();
    scatter_kernel<<<g2, kBlock, 0, st>>>(x2, iperm_c, n1, n2, z, d_skip);           // result[iperm_c[n1 + i]] = x2[i]
    PSB_LAUNCH_CHECK();
    if (n1 > 0) {
      EpiArgs ea; ea.f = y1;
      rc = spmv_launch(U12, EPI_RESID, x2, s1, ea, d_skip, st);                      // s1 = y1 - U12 x2
      if (rc != PSB_OK) return rc;
      rc = trsv_solve(U11, s1, y1, nullptr, z, iperm_c, d_skip, st);                 // x1 = U11^-1 s1, scattered
      if (rc != PSB_OK) return rc;
    }
    return PSB_OK;
  }
};

}  // namespace

}  // namespace psb

using namespace psb;

extern "C" int psb_splitlu_create(int64_t n, int64_t n1, psb_trsv_t L11, psb_trsv_t U11, psb_csr_t L21,
                                  psb_csr_t U12, const double* d_invL22, const double* d_invU22,
                                  const int32_t* h_perm_r, const int32_t* h_perm_c, void* stream,
                                  psb_prec_t* out) {
  PSB_REQUIRE(out && h_perm_r && h_perm_c && d_invL22 && d_invU22, PSB_ERR_ARG, "psb_splitlu_create: NULL argument");
  PSB_REQUIRE(n >= 1 && n1 >= 0 && n1 < n, PSB_ERR_ARG, "psb_splitlu_create: need 0 <= n1 < n");
  const int64_t n2 = n - n1;
  PSB_REQUIRE(n2 < (int64_t)46000, PSB_ERR_UNSUPP, "psb_splitlu_create: trailing block too large for int32 indexing");
  if (n1 > 0) {
    PSB_REQUIRE(L11 && U11 && L21 && U12, PSB_ERR_ARG, "psb_splitlu_create: leading blocks missing");
    PSB_REQUIRE(L11->n == n1 && U11->n == n1 && L11->lower && !U11->lower, PSB_ERR_ARG,
                "psb_splitlu_create: L11 / U11 must be lower / upper factors of order n1");
    PSB_REQUIRE(L21->n_rows == n2 && L21->n_cols == n1 && U12->n_rows == n1 && U12->n_cols == n2, PSB_ERR_ARG,
                "psb_splitlu_create: L21 must be n2 x n1 and U12 n1 x n2");
  }
  SplitLuPrec* P = new (std::nothrow) SplitLuPrec();
  PSB_REQUIRE(P != nullptr, PSB_ERR_ARG, "psb_splitlu_create: out of host memory");
  P->n = n; P->n1 = n1; P->n2 = n2;
  P->L11 = n1 > 0 ? L11 : nullptr; P->U11 = n1 > 0 ? U11 : nullptr;
  P->L21 = n1 > 0 ? L21 : nullptr; P->U12 = n1 > 0 ? U12 : nullptr;
  P->invL22 = d_invL22; P->invU22 = d_invU22;
  std::vector<int32_t> ipr((size_t)n), ipc((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    if (h_perm_r[i] < 0 || h_perm_r[i] >= n || h_perm_c[i] < 0 || h_perm_c[i] >= n) {
      delete P; set_error("psb_splitlu_create: permutation entry out of range"); return PSB_ERR_ARG;
    }
    ipr[h_perm_r[i]] = (int32_t)i;     // (Pr v)[perm_r[i]] = v[i]
    ipc[h_perm_c[i]] = (int32_t)i;     // (Pc z)[i] = z[perm_c[i]]
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMalloc((void**)&P->iperm_r, (size_t)n * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc((void**)&P->iperm_c, (size_t)n * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc((void**)&P->buf, (size_t)(2 * n1 + 3 * n2) * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpyAsync(P->iperm_r, ipr.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(P->iperm_c, ipc.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { delete P; set_error("psb_splitlu_create: %s", cudaGetErrorString(e)); return PSB_ERR_CUDA; }
  *out = P;
  return PSB_OK;
}
