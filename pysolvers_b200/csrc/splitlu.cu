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
// The dense blocks are chosen by dependency LEVEL, not by position: for L the rows of its last
// levels (level >= l*), for U the rows of its first levels, each moved to the end by a symmetric
// permutation that keeps the factor triangular (a row of a late level never feeds an earlier one).
// At Bratu 1024^2 (175 104 coarse rows, 3 494 levels) a dense block of 8 192 rows leaves 218 levels
// for the sparse solves when chosen by level, 1 079 when it is simply the trailing rows.
// Agreement with SuperLU.solve: 5e-16 relative (tests/test_gpu_amg.py).
#include "prec.cuh"
#include "spmv.cuh"
#include "sptrsv.cuh"

#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

namespace psb {

namespace {

constexpr int kGemvWarps = 8;

// y = M x with M dense row-major n x n, lower (j <= i) or upper (j >= i) triangular: only the
// triangle is read.  One CTA per row (a warp per row left the 64 KB rows of an 8 192-row block on one
// warp with one 512-byte load in flight: 70 us, 3.8 TB/s): thread t takes the entries 2 t, 2 t + 1
// (mod 512), four 16-byte loads in flight, sums them in index order; the 256 partial sums are
// combined by a fixed shuffle tree and in warp order: deterministic.
template <bool kLower>
__global__ void __launch_bounds__(kGemvWarps * 32)
tri_gemv_kernel(const double* __restrict__ M, int n, const double* __restrict__ x,
                double* __restrict__ y, const int* d_skip) {
  __shared__ double s_part[kGemvWarps];
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;
  constexpr int kT = kGemvWarps * 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int round = 0; round * (int)gridDim.x < n; ++round) {
    // work item t: even -> from the long end of the triangle, odd -> from the short end; the
    // rotation by `round` makes a CTA alternate between the two whatever the parity of the grid
    const int t = round * (int)gridDim.x + (int)((blockIdx.x + (unsigned)round) % gridDim.x);
    if (t >= n) continue;
    const int p = (t & 1) ? n - 1 - (t >> 1) : (t >> 1);   // position counted from the long end
    const int i = kLower ? n - 1 - p : p;
    const int j0 = kLower ? 0 : i, j1 = kLower ? i + 1 : n;
    const double* row = M + (int64_t)i * n;
    double acc = 0.0;
    int j;
    if ((n & 1) == 0) {
      const int ja = (j0 + 1) & ~1;                 // first even index >= j0 (rows start 16-byte aligned)
      if (tid == 0 && ja > j0) acc = row[j0] * x[j0];
      for (j = ja + tid * 2; j + 1 + 6 * kT < j1; j += 8 * kT) {
        const double2 m0 = ld_stream2(row + j), m1 = ld_stream2(row + j + 2 * kT);
        const double2 m2 = ld_stream2(row + j + 4 * kT), m3 = ld_stream2(row + j + 6 * kT);
        acc += m0.x * x[j];              acc += m0.y * x[j + 1];
        acc += m1.x * x[j + 2 * kT];     acc += m1.y * x[j + 2 * kT + 1];
        acc += m2.x * x[j + 4 * kT];     acc += m2.y * x[j + 4 * kT + 1];
        acc += m3.x * x[j + 6 * kT];     acc += m3.y * x[j + 6 * kT + 1];
      }
      for (; j + 1 < j1; j += 2 * kT) {
        const double2 m2 = ld_stream2(row + j);
        acc += m2.x * x[j];
        acc += m2.y * x[j + 1];
      }
      if (j < j1) acc += row[j] * x[j];
    } else {
      for (j = j0 + tid; j < j1; j += kT) acc += row[j] * x[j];
    }
    acc = warp_sum(acc);
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < kGemvWarps; ++w) tot += s_part[w];
      y[i] = tot;
    }
    __syncthreads();
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

// Block-diagonal stage of the supernodal collapse (Linear/supernodes.py): y = B x where B is the
// identity outside the collapsed supernodes and a small dense triangle inside; row r reads
// x[row0[r] + c], c in [c_lo[r], c_hi[r]), with the weights vals[off[r] + c - c_lo[r]] in that order.
struct BlockDiag {
  int64_t n = 0;
  int32_t* row0 = nullptr;
  int32_t* c_lo = nullptr;
  int32_t* c_hi = nullptr;
  int64_t* off = nullptr;
  double* vals = nullptr;
  int32_t* long_rows = nullptr;    // rows of >= kLongRow entries (warp-per-row kernel)
  int64_t n_long = 0;
  void release() {
    cudaFree(row0); cudaFree(c_lo); cudaFree(c_hi); cudaFree(off); cudaFree(vals); cudaFree(long_rows);
    row0 = c_lo = c_hi = long_rows = nullptr; off = nullptr; vals = nullptr; n = 0; n_long = 0;
  }
};

// y = blockdiag(B) x: row r of a collapsed supernode starting at line row0[r] reads x[row0 + c] for
// c in [c_lo, c_hi) with the weights vals[off + c - c_lo]; rows outside a supernode copy.
// Two row classes: a thread per row for the short rows (sequential sum), and a WARP per row for the
// rows of >= kLongRow entries (the large supernodes near the top of the elimination tree, hundreds
// of lines: with a thread per row a few hundred uncoalesced entries held the stage for 120 us;
// listed at set-up time, coalesced 256-byte reads, four in flight per lane, fixed-order reduction).
constexpr int kLongRow = 32;

// One launch: the first `g_long` CTAs take the listed long rows (a warp each), the others the short
// rows (a thread each) -- the two parts are independent and overlap.
__global__ void __launch_bounds__(kBlock)
blockdiag_kernel(int64_t n, int g_long, int64_t n_long, const int32_t* __restrict__ long_rows,
                 const int32_t* __restrict__ row0, const int32_t* __restrict__ c_lo,
                 const int32_t* __restrict__ c_hi, const int64_t* __restrict__ off,
                 const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
                 const int* d_skip) {
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;
  if ((int)blockIdx.x < g_long) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)kBlock + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)g_long * kBlock) >> 5;
    for (int64_t i = warp; i < n_long; i += n_warps) {
      const int r = long_rows[i];
      const int lo = c_lo[r], hi = c_hi[r];
      const double* w = vals + off[r] - lo;
      const double* xv = x + row0[r];
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      for (int c = lo + lane; c < hi; c += 128) {
        const int c1 = c + 32, c2 = c + 64, c3 = c + 96;
        const double w0 = w[c], x0 = xv[c];
        const double w1 = c1 < hi ? w[c1] : 0.0, x1 = c1 < hi ? xv[c1] : 0.0;
        const double w2 = c2 < hi ? w[c2] : 0.0, x2 = c2 < hi ? xv[c2] : 0.0;
        const double w3 = c3 < hi ? w[c3] : 0.0, x3 = c3 < hi ? xv[c3] : 0.0;
        a0 += w0 * x0; a1 += w1 * x1; a2 += w2 * x2; a3 += w3 * x3;
      }
      double a = (a0 + a1) + (a2 + a3);
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
      if (lane == 0) y[r] = a;
    }
    return;
  }
  const int64_t g_short = (int64_t)gridDim.x - g_long;
  for (int64_t r = (blockIdx.x - g_long) * (int64_t)kBlock + threadIdx.x; r < n; r += g_short * kBlock) {
    const int lo = c_lo[r], hi = c_hi[r];
    if (hi - lo >= kLongRow) continue;                   // a listed long row
    if (hi <= lo) { y[r] = x[r]; continue; }
    const double* xv = x + row0[r];
    const double* w = vals + off[r];
    double acc = 0.0;
    for (int c = lo; c < hi; ++c) acc += w[c - lo] * xv[c];
    y[r] = acc;
  }
}

static int blockdiag_apply(const BlockDiag& B, int64_t n, const double* x, double* y, const int* d_skip,
                           cudaStream_t st) {
  const int g_short = (int)std::max<int64_t>(1, std::min<int64_t>((n + kBlock - 1) / kBlock, (int64_t)sm_count() * 6));
  const int g_long = B.n_long > 0
      ? (int)std::max<int64_t>(1, std::min<int64_t>((B.n_long * 32 + kBlock - 1) / kBlock, (int64_t)sm_count() * 2)) : 0;
  blockdiag_kernel<<<g_long + g_short, kBlock, 0, st>>>(n, g_long, B.n_long, B.long_rows, B.row0, B.c_lo, B.c_hi,
                                                       B.off, B.vals, x, y, d_skip);
  PSB_LAUNCH_CHECK();
  return PSB_OK;
}

struct SplitLuPrec : psb_prec {
  // L and U are split independently (each in its own symmetric permutation, chosen by the host:
  // the dense block of L holds its LAST dependency levels, the one of U its FIRST ones):
  //   L-stage works on positions p: row p takes v[map_in[p]];  n1L sparse + n2L dense rows
  //   U-stage element p takes ycat[map_mid[p]];                 n1U sparse + n2U dense rows
  //   result[map_out[p]] = x[p]
  int64_t n1L = 0, n2L = 0, n1U = 0, n2U = 0;
  psb_trsv* L11 = nullptr;      // not owned (null when n1L == 0)
  psb_trsv* U11 = nullptr;      // not owned (null when n1U == 0)
  const psb_csr* L21 = nullptr; // n2L x n1L, not owned
  const psb_csr* U12 = nullptr; // n1U x n2U
  const double* invL22 = nullptr;   // n2L x n2L row-major, not owned
  const double* invU22 = nullptr;   // n2U x n2U
  int32_t* map_in = nullptr;    // owned
  int32_t* map_mid = nullptr;   // owned (null: identity, same split for L and U)
  int32_t* map_out = nullptr;   // owned
  double* buf = nullptr;        // owned: ycat (n) | yU (n) | w2 (max n2) | t2 (max n2) | s1 (n1U) | tmp (max n1)
  BlockDiag bdL, bdU;           // owned: block-diagonal stages of collapsed supernodes (n == 0: none)
  ~SplitLuPrec() override {
    cudaFree(map_in); cudaFree(map_mid); cudaFree(map_out); cudaFree(buf);
    bdL.release(); bdU.release();
  }
  const char* kind() const override { return "splitlu"; }
  int check_error() override {
    int a = 0, b = 0;
    a = trsv_take_error(L11);
    b = trsv_take_error(U11);
    return a | b;
  }
  int apply(const double* r, double* z, const int* d_skip, cudaStream_t st) override {
    const int64_t n2max = std::max(n2L, n2U);
    double* ycat = buf;                 // [y1 (n1L) | y2 (n2L)] in L order
    double* yU = buf + n;               // the same vector in U order (aliases ycat when map_mid is null)
    double* w2 = buf + 2 * n;
    double* t2 = w2 + n2max;
    double* s1 = t2 + n2max;
    double* tmp = s1 + n1U;
    auto grid_of = [&](int64_t cnt) {
      return (int)std::max<int64_t>(1, std::min<int64_t>((cnt + kBlock - 1) / kBlock, (int64_t)sm_count() * 4));
    };
    auto gemv_grid = [&](int64_t rows) {
      return (int)std::max<int64_t>(1, std::min<int64_t>(rows, (int64_t)sm_count() * 8));      // a CTA per row
    };
    int rc = PSB_OK;
    // ---- L stage: y1 = L11^-1 w1 ; y2 = inv(L22) (w2 - L21 y1) ------------------------------
    double* y1 = ycat;
    double* y2 = ycat + n1L;
    if (n1L > 0) {
      // with collapsed supernodes: z1 = L~11^-1 w1, then y1 = blockdiag(D^-1) z1
      rc = trsv_solve(L11, r, bdL.n ? tmp : y1, map_in, nullptr, nullptr, d_skip, st);
      if (rc != PSB_OK) return rc;
      if (bdL.n) {
        rc = blockdiag_apply(bdL, n1L, tmp, y1, d_skip, st);
        if (rc != PSB_OK) return rc;
      }
    }
    gather_kernel<<<grid_of(n2L), kBlock, 0, st>>>(r, map_in, n1L, n2L, w2, d_skip);
    PSB_LAUNCH_CHECK();
    const double* rhs2 = w2;
    if (n1L > 0) {
      EpiArgs ea; ea.f = w2;
      rc = spmv_launch(L21, EPI_RESID, y1, t2, ea, d_skip, st);
      if (rc != PSB_OK) return rc;
      rhs2 = t2;
    }
    tri_gemv_kernel<true><<<gemv_grid(n2L), kGemvWarps * 32, 0, st>>>(invL22, (int)n2L, rhs2, y2, d_skip);
    PSB_LAUNCH_CHECK();
    // ---- the same vector in the order of the U split ------------------------------------------
    const double* yu = ycat;
    if (map_mid != nullptr) {
      gather_kernel<<<grid_of(n), kBlock, 0, st>>>(ycat, map_mid, 0, n, yU, d_skip);
      PSB_LAUNCH_CHECK();
      yu = yU;
    }
    // ---- U stage: x2 = inv(U22) y2 ; x1 = U11^-1 (y1 - U12 x2) --------------------------------
    double* x2 = w2;
    tri_gemv_kernel<false><<<gemv_grid(n2U), kGemvWarps * 32, 0, st>>>(invU22, (int)n2U, yu + n1U, x2, d_skip);
    PSB_LAUNCH_CHECK();
    scatter_kernel<<<grid_of(n2U), kBlock, 0, st>>>(x2, map_out, n1U, n2U, z, d_skip);
    PSB_LAUNCH_CHECK();
    if (n1U > 0) {
      EpiArgs ea; ea.f = yu;
      rc = spmv_launch(U12, EPI_RESID, x2, s1, ea, d_skip, st);          // s1 = y1 - U12 x2
      if (rc != PSB_OK) return rc;
      const double* rhs1 = s1;
      if (bdU.n) {                                                         // s1' = blockdiag(D^-1) s1, then U~11
        rc = blockdiag_apply(bdU, n1U, s1, tmp, d_skip, st);
        if (rc != PSB_OK) return rc;
        rhs1 = tmp;
      }
      rc = trsv_solve(U11, rhs1, ycat, nullptr, z, map_out, d_skip, st);   // x1, scattered into z; ycat is free again
      if (rc != PSB_OK) return rc;
    }
    return PSB_OK;
  }
};

}  // namespace

}  // namespace psb

using namespace psb;

static int upload_map(int32_t** d, const int32_t* h, int64_t n, cudaStream_t st) {
  PSB_CUDA(cudaMalloc((void**)d, (size_t)n * sizeof(int32_t)));
  PSB_CUDA(cudaMemcpyAsync(*d, h, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  return PSB_OK;
}

extern "C" int psb_splitlu2_create(int64_t n, int64_t n1L, int64_t n1U, psb_trsv_t L11, psb_trsv_t U11,
                                   psb_csr_t L21, psb_csr_t U12, const double* d_invL22,
                                   const double* d_invU22, const int32_t* h_map_in,
                                   const int32_t* h_map_mid, const int32_t* h_map_out, void* stream,
                                   psb_prec_t* out) {
  PSB_REQUIRE(out && h_map_in && h_map_out && d_invL22 && d_invU22, PSB_ERR_ARG, "psb_splitlu2_create: NULL argument");
  PSB_REQUIRE(n >= 1 && n1L >= 0 && n1L < n && n1U >= 0 && n1U < n, PSB_ERR_ARG,
              "psb_splitlu2_create: need 0 <= n1 < n for both factors");
  const int64_t n2L = n - n1L, n2U = n - n1U;
  PSB_REQUIRE(n2L < (int64_t)46000 && n2U < (int64_t)46000, PSB_ERR_UNSUPP,
              "psb_splitlu2_create: dense block too large for int32 indexing");
  if (n1L > 0) {
    PSB_REQUIRE(L11 && L21, PSB_ERR_ARG, "psb_splitlu2_create: leading blocks of L missing");
    PSB_REQUIRE(L11->n == n1L && L11->lower && L21->n_rows == n2L && L21->n_cols == n1L, PSB_ERR_ARG,
                "psb_splitlu2_create: L11 must be a lower factor of order n1L and L21 n2L x n1L");
  }
  if (n1U > 0) {
    PSB_REQUIRE(U11 && U12, PSB_ERR_ARG, "psb_splitlu2_create: leading blocks of U missing");
    PSB_REQUIRE(U11->n == n1U && !U11->lower && U12->n_rows == n1U && U12->n_cols == n2U, PSB_ERR_ARG,
                "psb_splitlu2_create: U11 must be an upper factor of order n1U and U12 n1U x n2U");
  }
  for (int64_t i = 0; i < n; ++i) {
    const bool bad = h_map_in[i] < 0 || h_map_in[i] >= n || h_map_out[i] < 0 || h_map_out[i] >= n ||
                     (h_map_mid && (h_map_mid[i] < 0 || h_map_mid[i] >= n));
    PSB_REQUIRE(!bad, PSB_ERR_ARG, "psb_splitlu2_create: map entry out of range");
  }
  SplitLuPrec* P = new (std::nothrow) SplitLuPrec();
  PSB_REQUIRE(P != nullptr, PSB_ERR_ARG, "psb_splitlu2_create: out of host memory");
  P->n = n; P->n1L = n1L; P->n2L = n2L; P->n1U = n1U; P->n2U = n2U;
  P->L11 = n1L > 0 ? L11 : nullptr; P->L21 = n1L > 0 ? L21 : nullptr;
  P->U11 = n1U > 0 ? U11 : nullptr; P->U12 = n1U > 0 ? U12 : nullptr;
  P->invL22 = d_invL22; P->invU22 = d_invU22;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = upload_map(&P->map_in, h_map_in, n, st);
  if (rc == PSB_OK) rc = upload_map(&P->map_out, h_map_out, n, st);
  if (rc == PSB_OK && h_map_mid) rc = upload_map(&P->map_mid, h_map_mid, n, st);
  if (rc == PSB_OK) {
    const int64_t n2max = std::max(n2L, n2U);
    cudaError_t e = cudaMalloc((void**)&P->buf, (size_t)(2 * n + 2 * n2max + n1U + std::max(n1L, n1U) + 1) * sizeof(double));
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error("psb_splitlu2_create: %s", cudaGetErrorString(e)); rc = PSB_ERR_CUDA; }
  }
  if (rc != PSB_OK) { delete P; return rc; }
  *out = P;
  return PSB_OK;
}

extern "C" int psb_splitlu_set_blockdiag(psb_prec_t P_, int upper, int64_t n_rows, const int32_t* h_row0,
                                         const int32_t* h_c_lo, const int32_t* h_c_hi, const int64_t* h_off,
                                         const double* h_vals, int64_t n_vals, void* stream) {
  PSB_REQUIRE(P_ && h_row0 && h_c_lo && h_c_hi && h_off && (n_vals == 0 || h_vals), PSB_ERR_ARG,
              "psb_splitlu_set_blockdiag: NULL argument");
  PSB_REQUIRE(strcmp(P_->kind(), "splitlu") == 0, PSB_ERR_ARG, "psb_splitlu_set_blockdiag: not a split LU");
  SplitLuPrec* P = static_cast<SplitLuPrec*>(P_);
  PSB_REQUIRE(n_rows == (upper ? P->n1U : P->n1L) && n_rows > 0, PSB_ERR_ARG,
              "psb_splitlu_set_blockdiag: one entry per row of the sparse leading block");
  for (int64_t r = 0; r < n_rows; ++r) {
    const bool bad = h_c_lo[r] < 0 || h_c_hi[r] < h_c_lo[r] || h_row0[r] < 0 || h_row0[r] + h_c_hi[r] > n_rows ||
                     (h_c_hi[r] > h_c_lo[r] && (h_off[r] < 0 || h_off[r] + (h_c_hi[r] - h_c_lo[r]) > n_vals));
    PSB_REQUIRE(!bad, PSB_ERR_ARG, "psb_splitlu_set_blockdiag: row descriptor out of range");
  }
  BlockDiag& B = upper ? P->bdU : P->bdL;
  B.release();
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMalloc((void**)&B.row0, (size_t)n_rows * 4);
  if (e == cudaSuccess) e = cudaMalloc((void**)&B.c_lo, (size_t)n_rows * 4);
  if (e == cudaSuccess) e = cudaMalloc((void**)&B.c_hi, (size_t)n_rows * 4);
  if (e == cudaSuccess) e = cudaMalloc((void**)&B.off, (size_t)n_rows * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&B.vals, (size_t)std::max<int64_t>(n_vals, 1) * 8);
  if (e == cudaSuccess) e = cudaMemcpyAsync(B.row0, h_row0, (size_t)n_rows * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(B.c_lo, h_c_lo, (size_t)n_rows * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(B.c_hi, h_c_hi, (size_t)n_rows * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(B.off, h_off, (size_t)n_rows * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n_vals) e = cudaMemcpyAsync(B.vals, h_vals, (size_t)n_vals * 8, cudaMemcpyHostToDevice, st);
  std::vector<int32_t> longs;
  for (int64_t r = 0; r < n_rows; ++r) if (h_c_hi[r] - h_c_lo[r] >= kLongRow) longs.push_back((int32_t)r);
  if (e == cudaSuccess && !longs.empty()) {
    e = cudaMalloc((void**)&B.long_rows, longs.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(B.long_rows, longs.data(), longs.size() * 4, cudaMemcpyHostToDevice, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { B.release(); set_error("psb_splitlu_set_blockdiag: %s", cudaGetErrorString(e)); return PSB_ERR_CUDA; }
  B.n = n_rows;
  B.n_long = (int64_t)longs.size();
  return PSB_OK;
}

// the same split position for both factors, permutations as SuperLU reports them
extern "C" int psb_splitlu_create(int64_t n, int64_t n1, psb_trsv_t L11, psb_trsv_t U11, psb_csr_t L21,
                                  psb_csr_t U12, const double* d_invL22, const double* d_invU22,
                                  const int32_t* h_perm_r, const int32_t* h_perm_c, void* stream,
                                  psb_prec_t* out) {
  PSB_REQUIRE(h_perm_r && h_perm_c && n >= 1, PSB_ERR_ARG, "psb_splitlu_create: NULL argument");
  std::vector<int32_t> ipr((size_t)n), ipc((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    PSB_REQUIRE(h_perm_r[i] >= 0 && h_perm_r[i] < n && h_perm_c[i] >= 0 && h_perm_c[i] < n, PSB_ERR_ARG,
                "psb_splitlu_create: permutation entry out of range");
    ipr[h_perm_r[i]] = (int32_t)i;     // (Pr v)[perm_r[i]] = v[i]
    ipc[h_perm_c[i]] = (int32_t)i;     // (Pc z)[i] = z[perm_c[i]]
  }
  return psb_splitlu2_create(n, n1, n1, L11, U11, L21, U12, d_invL22, d_invU22, ipr.data(), nullptr, ipc.data(),
                             stream, out);
}

// dependency levels of a triangular CSR matrix in host memory (level = 1 + max level of the rows a
// row depends on): the host uses them to choose the dense blocks
extern "C" int psb_tri_levels(int64_t n, const int32_t* h_rowptr, const int32_t* h_colind, int lower,
                              int32_t* h_level) {
  PSB_REQUIRE(n >= 0 && h_rowptr && h_level && (n == 0 || h_colind), PSB_ERR_ARG, "psb_tri_levels: NULL argument");
  auto visit = [&](int64_t i) {
    int32_t lv = 0;
    for (int32_t p = h_rowptr[i]; p < h_rowptr[i + 1]; ++p) {
      const int32_t j = h_colind[p];
      if (lower ? (j < i) : (j > i)) lv = std::max(lv, h_level[j] + 1);
    }
    h_level[i] = lv;
  };
  if (lower) for (int64_t i = 0; i < n; ++i) visit(i);
  else       for (int64_t i = n - 1; i >= 0; --i) visit(i);
  return PSB_OK;
}

// Heights of the rows of an UPPER triangular CSR matrix in its elimination tree: h(i) = 1 + max h(k)
// over the rows k < i that need row i (U[k, i] != 0), i.e. the dependency level of row i in U^T --
// computed by pushing along the rows, without transposing (host code).
extern "C" int psb_tri_heights_upper(int64_t n, const int32_t* h_rowptr, const int32_t* h_colind, int32_t* h_height) {
  PSB_REQUIRE(n >= 0 && h_rowptr && h_height && (n == 0 || h_colind), PSB_ERR_ARG, "psb_tri_heights_upper: NULL argument");
  for (int64_t i = 0; i < n; ++i) h_height[i] = 0;
  for (int64_t k = 0; k < n; ++k) {
    const int32_t hk = h_height[k] + 1;
    for (int32_t p = h_rowptr[k]; p < h_rowptr[k + 1]; ++p) {
      const int32_t i = h_colind[p];
      if (i > k && h_height[i] < hk) h_height[i] = hk;
    }
  }
  return PSB_OK;
}
