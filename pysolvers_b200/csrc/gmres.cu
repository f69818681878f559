// Right-preconditioned, un-restarted GMRES with the whole Arnoldi / Givens loop on the device.
//
// Replaces the Python loop of PySolvers/Linear/GMRESSolver.py:75-174 (+ Givens.py:7-34):
//
//   z = M^-1 q_k ; w = A z                      preconditioner + spmv.cu
//   orthogonalise w against q_0..q_k            CGS2 (default): two rounds of
//                                               batched dots (8 basis vectors per pass over w)
//                                               + batched axpy; or MGS (reference order,
//                                               GMRESSolver.py:110-112) as k+2 fused
//                                               "axpy-then-dot" passes
//   h_{k+1,k} = ||w|| ; Givens ; residual       gmres_givens_kernel (one warp): lucky-breakdown
//                                               test :121-123, previous rotations :133-135, new
//                                               rotation with the plain-sqrt formula of
//                                               Givens.py:8-10, rotate g, |g_{k+1}|, convergence
//   q_{k+1} = w / h_{k+1,k}                     scale kernel (division, as :125)
//
// and, after the loop: back-substitution R y = g, x = M^-1 (Q y), true residual b - A x and
// its norm (:159-166).  The basis Q is stored vector-contiguous (the reference's n x (m+1)
// row-major array makes every basis vector a strided column).  The Hessenberg column, the
// rotations, g, the iteration counter and the flags live in device memory; the host polls a
// pinned copy one iteration behind.
#include <cstdlib>
#include "prec.cuh"
#include "spmv.cuh"

#include <algorithm>

namespace psb {

constexpr int kCh = 8;     // basis vectors handled per pass of the batched kernels

struct GmresState {
  double norm_b, beta, tau;
  double ww;          // ||w||^2 after orthogonalisation
  double hk1;         // h_{k+1,k}
  double norm_r_rec;  // |g_{k+1}|
  double norm_true;
  double tmp_dot;
  int k, maxiter, done, status, k_final, lucky, n_hist, pad;
};

struct GmresSmall {      // small dense arrays, all in device memory
  double* hcol;   // [maxiter + 2] current Hessenberg column (CGS2 round-1 values)
  double* hcol2;  // [maxiter + 2] CGS2 round-2 corrections
  double* R;      // [maxiter * (maxiter + 1)] rotated columns, column k at R + k*(maxiter+1)
  double* cs;     // [maxiter]
  double* sn;     // [maxiter]
  double* g;      // [maxiter + 1]
  double* y;      // [maxiter]
};

__device__ __forceinline__ double2 ld2rw(const double* p) {
  double2 v;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// out = a / s  (s read from device memory); used for q_0 = b/beta and q_{k+1} = w/h
__global__ void __launch_bounds__(kBlock)
gmres_scale_kernel(const GmresState* st, int64_t n, const double* __restrict__ a,
                   const double* __restrict__ s_ptr, double* __restrict__ out, int check_done) {
  if (check_done && ld_cg(&st->done) != 0) return;
  const double s = ld_cg(s_ptr);
  const int64_t n2 = n >> 1;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kBlock) {
    double2 v = ld_stream2(a + 2 * i);
    v.x = v.x / s; v.y = v.y / s;
    st_stream2(out + 2 * i, v);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[n - 1] = a[n - 1] / s;
}

// norm_b = ||b||; trivial check; beta; g[0]
__global__ void __launch_bounds__(kBlock)
gmres_init_kernel(GmresState* st, GmresSmall sm, int64_t n, const double* __restrict__ b, ReduceBuf rb) {
  __shared__ double scratch[kWarps];
  double acc = 0.0;
  const int64_t n2 = n >> 1;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kBlock) {
    double2 v = ld_stream2(b + 2 * i);
    acc += v.x * v.x;
    acc += v.y * v.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) acc += b[n - 1] * b[n - 1];
  double t = block_sum(acc, scratch);
  if (threadIdx.x == 0) rb.partials[blockIdx.x] = t;
  if (last_block(rb.ticket)) {
    double bb = sum_partials(rb.partials, gridDim.x, scratch);
    if (threadIdx.x == 0) st->tmp_dot = bb;           // this rank's part when row-partitioned
  }
}

// after b.b is complete (all-reduced when row-partitioned)
__global__ void gmres_init_finish_kernel(GmresState* st, GmresSmall sm) {
  if (threadIdx.x != 0) return;
  const double nb = sqrt(st->tmp_dot);
  st->norm_b = nb; st->beta = nb;
  sm.g[0] = nb;                                       // g = beta * e1   (GMRESSolver.py:95-97)
  if (nb == 0.0) { st->done = 1; st->status = PSB_TRIVIAL; st->k_final = 0; }
}

__global__ void gmres_norm_finish_kernel(GmresState* st) {
  if (threadIdx.x == 0) st->norm_true = sqrt(st->tmp_dot);
}

__global__ void __launch_bounds__(kBlock)
gmres_copy_kernel(const GmresState* st, int64_t n, const double* __restrict__ src, double* __restrict__ dst,
                  int check_done) {
  if (check_done && ld_cg(&st->done) != 0) return;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    dst[i] = src[i];
}

// MGS step j of iteration k (reference order, GMRESSolver.py:110-115):
//   j > 0 : w -= h_{j-1} q_{j-1}
//   j <= k: h_j = q_j . w          j == k+1: ww = w . w
__global__ void __launch_bounds__(kBlock)
gmres_mgs_kernel(GmresState* st, GmresSmall sm, int64_t n, const double* __restrict__ Q, int64_t ldq,
                 double* __restrict__ w, int j, int k, ReduceBuf rb) {
  __shared__ double scratch[kWarps];
  if (ld_cg(&st->done) != 0) return;
  const bool sub = j > 0;
  const bool last = j == k + 1;
  const double hprev = sub ? ld_cg(sm.hcol + (j - 1)) : 0.0;
  const double* qp = Q + (int64_t)(j - 1) * ldq;
  const double* qj = Q + (int64_t)j * ldq;
  double acc = 0.0;
  const int64_t n2 = n >> 1;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kBlock) {
    double2 wv = ld2rw(w + 2 * i);
    if (sub) {
      double2 q = ld_stream2(qp + 2 * i);
      wv.x = wv.x - hprev * q.x; wv.y = wv.y - hprev * q.y;
      st_stream2(w + 2 * i, wv);
    }
    if (last) { acc += wv.x * wv.x; acc += wv.y * wv.y; }
    else { double2 q = ld_stream2(qj + 2 * i); acc += q.x * wv.x; acc += q.y * wv.y; }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t e = n - 1;
    double wv = w[e];
    if (sub) { wv = wv - hprev * qp[e]; w[e] = wv; }
    acc += last ? wv * wv : qj[e] * wv;
  }
  double t = block_sum(acc, scratch);
  if (threadIdx.x == 0) rb.partials[blockIdx.x] = t;
  if (last_block(rb.ticket)) {
    double s = sum_partials(rb.partials, gridDim.x, scratch);
    if (threadIdx.x == 0) { if (last) st->ww = s; else sm.hcol[j] = s; }
  }
}

// batched dots: out[j0 + c] = q_{j0+c} . w for c < cnt (<= kCh); one pass over w
__global__ void __launch_bounds__(kBlock)
gmres_multidot_kernel(const GmresState* st, int64_t n, const double* __restrict__ Q, int64_t ldq,
                      const double* __restrict__ w, int j0, int cnt, double* out, ReduceBuf rb) {
  __shared__ double scratch[kWarps];
  if (ld_cg(&st->done) != 0) return;
  double acc[kCh];
#pragma unroll
  for (int c = 0; c < kCh; ++c) acc[c] = 0.0;
  const double* q0 = Q + (int64_t)j0 * ldq;
  const int64_t n2 = n >> 1;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kBlock) {
    const double2 wv = ld_stream2(w + 2 * i);
#pragma unroll
    for (int c = 0; c < kCh; ++c) {
      if (c < cnt) {
        const double2 q = ld_stream2(q0 + (int64_t)c * ldq + 2 * i);
        acc[c] += q.x * wv.x;
        acc[c] += q.y * wv.y;
      }
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
    for (int c = 0; c < kCh; ++c)
      if (c < cnt) acc[c] += q0[(int64_t)c * ldq + n - 1] * w[n - 1];
  }
#pragma unroll
  for (int c = 0; c < kCh; ++c) {
    if (c < cnt) {
      double t = block_sum(acc[c], scratch);
      if (threadIdx.x == 0) rb.partials[(int64_t)c * gridDim.x + blockIdx.x] = t;
    }
  }
  if (last_block(rb.ticket)) {
    for (int c = 0; c < cnt; ++c) {
      double s = sum_partials(rb.partials + (int64_t)c * gridDim.x, gridDim.x, scratch);
      if (threadIdx.x == 0) out[j0 + c] = s;
    }
  }
}

// batched axpy: w = (init ? 0 : w) + sign * sum_c coef[j0 + c] q_{j0+c}; optional w.w
__global__ void __launch_bounds__(kBlock, 4)
gmres_multiaxpy_kernel(GmresState* st, int64_t n, const double* __restrict__ Q, int64_t ldq,
                       double* __restrict__ w, const double* coef, int j0, int cnt, double sign,
                       int init_zero, int want_norm, int check_done, ReduceBuf rb) {
  __shared__ double scratch[kWarps];
  if (check_done && ld_cg(&st->done) != 0) return;
  double h[kCh];
#pragma unroll
  for (int c = 0; c < kCh; ++c) h[c] = (c < cnt) ? sign * ld_cg(coef + j0 + c) : 0.0;
  const double* q0 = Q + (int64_t)j0 * ldq;
  double acc = 0.0;
  const int64_t n2 = n >> 1;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kBlock) {
    double2 wv = init_zero ? make_double2(0.0, 0.0) : ld2rw(w + 2 * i);
#pragma unroll
    for (int c = 0; c < kCh; ++c) {
      if (c < cnt) {
        const double2 q = ld_stream2(q0 + (int64_t)c * ldq + 2 * i);
        wv.x = wv.x + h[c] * q.x;
        wv.y = wv.y + h[c] * q.y;
      }
    }
    st_stream2(w + 2 * i, wv);
    if (want_norm) { acc += wv.x * wv.x; acc += wv.y * wv.y; }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t e = n - 1;
    double wv = init_zero ? 0.0 : w[e];
#pragma unroll
    for (int c = 0; c < kCh; ++c)
      if (c < cnt) wv = wv + h[c] * q0[(int64_t)c * ldq + e];
    w[e] = wv;
    if (want_norm) acc += wv * wv;
  }
  if (want_norm) {
    double t = block_sum(acc, scratch);
    if (threadIdx.x == 0) rb.partials[blockIdx.x] = t;
    if (last_block(rb.ticket)) {
      double s = sum_partials(rb.partials, gridDim.x, scratch);
      if (threadIdx.x == 0) st->ww = s;
    }
  }
}

// CGS2, last chunk of round 1 fused with the round-2 dots OF THE SAME CHUNK: w = w - sum_c h1_c q_c
// is final after this chunk, so q_c . w (the round-2 coefficients of these <= 8 vectors) is formed
// from the q values already in registers -- one pass over w and the chunk's basis vectors less
// per iteration (all of round 2's dot pass while the Krylov dimension is <= 8).
__global__ void __launch_bounds__(kBlock, 3)
gmres_axpy_dot_kernel(GmresState* st, int64_t n, const double* __restrict__ Q, int64_t ldq,
                      double* __restrict__ w, const double* coef, int j0, int cnt, double* dot_out,
                      ReduceBuf rb) {
  __shared__ double scratch[kWarps];
  if (ld_cg(&st->done) != 0) return;
  double h[kCh], acc[kCh];
#pragma unroll
  for (int c = 0; c < kCh; ++c) { h[c] = (c < cnt) ? -ld_cg(coef + j0 + c) : 0.0; acc[c] = 0.0; }
  const double* q0 = Q + (int64_t)j0 * ldq;
  const int64_t n2 = n >> 1;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kBlock) {
    double2 wv = ld2rw(w + 2 * i);
    double2 q[kCh];
#pragma unroll
    for (int c = 0; c < kCh; ++c)
      if (c < cnt) q[c] = ld_stream2(q0 + (int64_t)c * ldq + 2 * i);
#pragma unroll
    for (int c = 0; c < kCh; ++c) {
      if (c < cnt) {
        wv.x = wv.x + h[c] * q[c].x;
        wv.y = wv.y + h[c] * q[c].y;
      }
    }
    st_stream2(w + 2 * i, wv);
#pragma unroll
    for (int c = 0; c < kCh; ++c) {
      if (c < cnt) {
        acc[c] += q[c].x * wv.x;
        acc[c] += q[c].y * wv.y;
      }
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t e = n - 1;
    double wv = w[e];
#pragma unroll
    for (int c = 0; c < kCh; ++c)
      if (c < cnt) wv = wv + h[c] * q0[(int64_t)c * ldq + e];
    w[e] = wv;
#pragma unroll
    for (int c = 0; c < kCh; ++c)
      if (c < cnt) acc[c] += q0[(int64_t)c * ldq + e] * wv;
  }
#pragma unroll
  for (int c = 0; c < kCh; ++c) {
    if (c < cnt) {
      double t = block_sum(acc[c], scratch);
      if (threadIdx.x == 0) rb.partials[(int64_t)c * gridDim.x + blockIdx.x] = t;
    }
  }
  if (last_block(rb.ticket)) {
    for (int c = 0; c < cnt; ++c) {
      double s = sum_partials(rb.partials + (int64_t)c * gridDim.x, gridDim.x, scratch);
      if (threadIdx.x == 0) dot_out[j0 + c] = s;
    }
  }
}

// CGS2 with up to kFusedMax basis vectors: round 1's axpy over ALL vectors fused with ALL of round 2's
// dots.  w1[i] = w[i] - sum_j h1_j q_j[i] depends on element i only, so a CTA finishes a tile of w1
// (256 x 2 elements, in registers) from the q tiles and then walks the same q tiles again -- now in
// L1 / L2 -- for q_j . w1: the basis is read from HBM once for both steps instead of twice
// (3 instead of 4 passes over the basis per iteration).  Accumulators: one shared-memory slot per
// (vector, thread) -- 32 doubles per thread in registers spilled at 128 registers.
constexpr int kFusedMax = 32;

__global__ void __launch_bounds__(kBlock, 2)
gmres_cgs2_fused_kernel(GmresState* st, int64_t n, const double* Q, int64_t ldq, double* __restrict__ w,
                        const double* coef, int cnt, double* dot_out, ReduceBuf rb) {
  __shared__ double scratch[kWarps];
  __shared__ double sh[kFusedMax];
  if (ld_cg(&st->done) != 0) return;
  extern __shared__ double sacc[];                       // [cnt][kBlock]
  if (threadIdx.x < kFusedMax) sh[threadIdx.x] = (int)threadIdx.x < cnt ? -ld_cg(coef + threadIdx.x) : 0.0;
  for (int j = 0; j < cnt; ++j) sacc[j * kBlock + threadIdx.x] = 0.0;
  __syncthreads();
  double* const acc = sacc + threadIdx.x;                // acc[j * kBlock]
  const int64_t n2 = n >> 1;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kBlock) {
    double2 wv = ld2rw(w + 2 * i);
#pragma unroll
    for (int j0 = 0; j0 < kFusedMax; j0 += 8) {
      if (j0 < cnt) {                                   // uniform
        double2 q[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (j0 + c < cnt) {                           // L1-allocating: read again below
            const double2* p = reinterpret_cast<const double2*>(Q + (int64_t)(j0 + c) * ldq + 2 * i);
            q[c] = *p;
          }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (j0 + c < cnt) { wv.x = wv.x + sh[j0 + c] * q[c].x; wv.y = wv.y + sh[j0 + c] * q[c].y; }
      }
    }
    st_stream2(w + 2 * i, wv);
#pragma unroll
    for (int j0 = 0; j0 < kFusedMax; j0 += 8) {
      if (j0 < cnt) {
        double2 q[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (j0 + c < cnt) q[c] = *reinterpret_cast<const double2*>(Q + (int64_t)(j0 + c) * ldq + 2 * i);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (j0 + c < cnt) {
            double a = acc[(j0 + c) * kBlock];
            a += q[c].x * wv.x; a += q[c].y * wv.y;
            acc[(j0 + c) * kBlock] = a;
          }
      }
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t e = n - 1;
    double wv = w[e];
    for (int j = 0; j < cnt; ++j) wv = wv + sh[j] * Q[(int64_t)j * ldq + e];
    w[e] = wv;
    for (int j = 0; j < cnt; ++j) acc[j * kBlock] += Q[(int64_t)j * ldq + e] * wv;
  }
  for (int j = 0; j < cnt; ++j) {
    const double t = block_sum(acc[j * kBlock], scratch);
    if (threadIdx.x == 0) rb.partials[(int64_t)j * gridDim.x + blockIdx.x] = t;
  }
  if (last_block(rb.ticket)) {
    for (int j = 0; j < cnt; ++j) {
      const double t = sum_partials(rb.partials + (int64_t)j * gridDim.x, gridDim.x, scratch);
      if (threadIdx.x == 0) dot_out[j] = t;
    }
  }
}

// One warp: finish column k of the Hessenberg matrix, rotate, test convergence.
__global__ void gmres_givens_kernel(GmresState* st, GmresSmall sm, double* __restrict__ hist, int cgs2) {
  if (st->done != 0) return;
  if (threadIdx.x != 0) return;
  const int k = st->k;
  const int ld = st->maxiter + 1;
  double* h = sm.hcol;
  if (cgs2) for (int j = 0; j <= k; ++j) h[j] = h[j] + sm.hcol2[j];
  const double hk1 = sqrt(st->ww);                               // GMRESSolver.py:115
  h[k + 1] = hk1;
  st->hk1 = hk1;
  double cn = 0.0;
  for (int j = 0; j <= k; ++j) cn += h[j] * h[j];
  cn = sqrt(cn);                                                 // :121
  const int lucky = fabs(hk1) <= 1.0e-16 * cn;                   // :122
  st->lucky = lucky;
  for (int j = 0; j < k; ++j) {                                  // :133-135
    const double c = sm.cs[j], s = sm.sn[j];
    const double a = h[j], b = h[j + 1];
    h[j] = c * a + s * b;
    h[j + 1] = -s * a + c * b;
  }
  const double hyp = sqrt(h[k + 1] * h[k + 1] + h[k] * h[k]);    // Givens.py:8
  const double s = h[k + 1] / hyp, c = h[k] / hyp;
  sm.cs[k] = c; sm.sn[k] = s;
  {
    const double a = h[k], b = h[k + 1];
    h[k] = c * a + s * b;
    h[k + 1] = -s * a + c * b;
    const double ga = sm.g[k], gb = 0.0;                         // g[k+1] is still zero
    sm.g[k] = c * ga + s * gb;
    sm.g[k + 1] = -s * ga + c * gb;
  }
  for (int j = 0; j <= k + 1 && j < ld; ++j) sm.R[(int64_t)k * ld + j] = h[j];
  const double nr = fabs(sm.g[k + 1]);                           // :152
  st->norm_r_rec = nr;
  hist[k] = nr;
  st->n_hist = k + 1;
  if (lucky || nr <= st->tau * st->norm_b) {                     // :158
    st->status = PSB_CONVERGED; st->k_final = k; st->done = 1;
  } else {
    st->k = k + 1;
    if (k + 1 >= st->maxiter) { st->status = PSB_MAXITER; st->k_final = k; st->done = 1; }
  }
}

// y = R^-1 g for the leading (kf+1) x (kf+1) block (GMRESSolver.py:159)
__global__ void gmres_backsolve_kernel(const GmresState* st, GmresSmall sm) {
  if (threadIdx.x != 0) return;
  const int m = st->k_final + 1;
  const int ld = st->maxiter + 1;
  for (int i = m - 1; i >= 0; --i) {
    double acc = sm.g[i];
    for (int j = i + 1; j < m; ++j) acc = acc - sm.R[(int64_t)j * ld + i] * sm.y[j];
    sm.y[i] = acc / sm.R[(int64_t)i * ld + i];
  }
}

__global__ void __launch_bounds__(kBlock)
gmres_norm_kernel(GmresState* st, int64_t n, const double* __restrict__ r, ReduceBuf rb) {
  __shared__ double scratch[kWarps];
  double acc = 0.0;
  const int64_t n2 = n >> 1;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kBlock) {
    double2 v = ld_stream2(r + 2 * i);
    acc += v.x * v.x;
    acc += v.y * v.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) acc += r[n - 1] * r[n - 1];
  double t = block_sum(acc, scratch);
  if (threadIdx.x == 0) rb.partials[blockIdx.x] = t;
  if (last_block(rb.ticket)) {
    double s = sum_partials(rb.partials, gridDim.x, scratch);
    if (threadIdx.x == 0) st->tmp_dot = s;
  }
}

constexpr int64_t kBasisPad = 288;   // elements; see basis_ld
static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// Leading dimension of the Krylov basis.  The batched CGS2 kernels stream 8 basis vectors and w at
// the same offset; with n a power of two the nine streams are exactly 2^k bytes apart and collide in
// the memory system, so consecutive vectors are staggered by a pad (multiple of 32 elements).
static int64_t basis_ld(int64_t n) {
  static int64_t pad = -1;
  if (pad < 0) { const char* e = getenv("PSB_GMRES_PAD"); pad = e ? align_up(atoll(e), 32) : kBasisPad; }
  return align_up(n, 32) + pad;
}

struct GmresWork {
  GmresState* st;
  ReduceBuf rb;
  GmresSmall sm;
  double *Q, *w, *z, *t;
  int64_t ldq;
};

static int64_t small_bytes(int64_t m) {
  return align_up(((m + 2) * 2 + m * (m + 1) + 2 * m + (m + 1) + m) * (int64_t)sizeof(double), 256);
}
static int64_t reduce_bytes() { return align_up((int64_t)sm_count() * 16 * 32 * sizeof(double), 256); }   // 32 = kFusedMax partial rows

static GmresWork carve(void* d_work, int64_t n, int64_t m) {
  char* base = (char*)d_work;
  GmresWork w;
  w.st = (GmresState*)base;
  w.rb.ticket = (unsigned int*)(base + 1024);
  w.rb.partials = (double*)(base + 4096);
  w.rb.max_grid = sm_count() * 16;
  double* s = (double*)(base + 4096 + reduce_bytes());
  w.sm.hcol = s;            s += m + 2;
  w.sm.hcol2 = s;           s += m + 2;
  w.sm.R = s;               s += m * (m + 1);
  w.sm.cs = s;              s += m;
  w.sm.sn = s;              s += m;
  w.sm.g = s;               s += m + 1;
  w.sm.y = s;
  w.ldq = basis_ld(n);
  char* v = base + 4096 + reduce_bytes() + small_bytes(m);
  const int64_t vec = w.ldq * (int64_t)sizeof(double);
  w.Q = (double*)v;                         v += (m + 1) * vec;
  w.w = (double*)v;                         v += vec;
  w.z = (double*)v;                         v += vec;
  w.t = (double*)v;
  return w;
}

struct GmresPoll {
  GmresState* pinned = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int init() {
    if (pinned) return PSB_OK;
    PSB_CUDA(cudaHostAlloc((void**)&pinned, 2 * sizeof(GmresState), cudaHostAllocDefault));
    PSB_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    PSB_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    return PSB_OK;
  }
};
static thread_local GmresPoll t_gpoll;

}  // namespace psb

using namespace psb;

extern "C" int64_t psb_gmres_workspace_bytes(int64_t n, int32_t maxiter) {
  if (n < 0 || maxiter < 1) return PSB_ERR_ARG;
  const int64_t m = maxiter;
  return 4096 + reduce_bytes() + small_bytes(m) + (m + 4) * basis_ld(n) * (int64_t)sizeof(double);
}

// One GMRES solve; A whole on this GPU (D == nullptr) or this rank's row block D of a
// row-partitioned system: then every vector is the rank's slice, the SpMV exchanges its halo
// (the input is staged in an extended buffer `xe`), and every reduction -- ||b||, the batched
// Gram-Schmidt dots, ||w||^2, the true residual -- is all-reduced between the kernel that forms the
// local sums and the one that consumes them.  All ranks take identical decisions.
static int gmres_solve_impl(psb_csr* A, psb_dist* D, psb_prec_t prec, const double* d_b, double* d_x,
                            void* d_work, int64_t work_bytes, int32_t maxiter, double tau, int32_t orth,
                            double* d_hist, psb_solve_result* result, cudaStream_t st) {
  const int64_t n = D ? dist_n_loc(D) : A->n_rows, m = maxiter;
  const bool has_prec = prec != nullptr;
  psb_comm* comm = D ? dist_comm(D) : nullptr;
  int rc = t_gpoll.init();
  if (rc != PSB_OK) return rc;

  GmresWork w = carve(d_work, n, m);
  double* xe = nullptr;                           // extended SpMV input (row-partitioned only)
  if (D) xe = (double*)((char*)d_work + psb_gmres_workspace_bytes(n, maxiter));
  (void)work_bytes;
  const int64_t small_total = 4096 + reduce_bytes() + small_bytes(m);
  PSB_CUDA(cudaMemsetAsync(d_work, 0, small_total, st));
  GmresState h0;
  memset(&h0, 0, sizeof(h0));
  h0.tau = tau; h0.maxiter = maxiter;
  PSB_CUDA(cudaMemcpyAsync(w.st, &h0, sizeof(h0), cudaMemcpyHostToDevice, st));
  PSB_CUDA(cudaStreamSynchronize(st));

  auto allreduce = [&](double* buf, int count) -> int { return comm ? dist_allreduce(comm, buf, count, st) : PSB_OK; };
  const int grid = stream_grid(n, w.rb.max_grid);
  // y = A x (x: n owned entries, not extended) with an SpMV epilogue
  auto matvec = [&](Epi epi, const double* x, double* y, const EpiArgs& ea, const int* skip) -> int {
    if (!D) return spmv_launch(A, epi, x, y, ea, skip, st);
    if (x != xe) {
      gmres_copy_kernel<<<grid, kBlock, 0, st>>>(w.st, n, x, xe, skip != nullptr ? 1 : 0);
      PSB_LAUNCH_CHECK();
    }
    return dist_spmv_epi(D, epi, xe, y, ea, skip, st);
  };
  // The basis kernels run as exactly ONE wave of resident CTAs (grid-stride inside): with the
  // generic 8 CTAs per SM and 3 - 5 resident ones the second wave left 40 % of the slots empty
  // (ncu: 53 % of the DRAM peak, profiles/round1h_gmres.md).
  auto one_wave = [&](const void* kernel, int* cache) -> int {
    if (*cache == 0) {
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
      *cache = per_sm * sm_count();
    }
    const int64_t need = ((n >> 1) + kBlock - 1) / kBlock;
    return (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(*cache, need), w.rb.max_grid));
  };
  static thread_local int wave_dot = 0, wave_axpy = 0, wave_mgs = 0, wave_fused = 0;
  const int grid_fused = one_wave((const void*)gmres_axpy_dot_kernel, &wave_fused);
  // dynamic shared memory of the fully fused CGS2 kernel: kFusedMax x kBlock accumulators
  static thread_local int wave_cgs2 = 0;
  constexpr size_t kFusedSmem = (size_t)kFusedMax * kBlock * sizeof(double);
  if (wave_cgs2 == 0) {
    PSB_CUDA(cudaFuncSetAttribute(gmres_cgs2_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gmres_cgs2_fused_kernel, kBlock, kFusedSmem) != cudaSuccess || per_sm < 1) per_sm = 1;
    wave_cgs2 = per_sm * sm_count();
  }
  const int grid_cgs2 = one_wave((const void*)gmres_cgs2_fused_kernel, &wave_cgs2);
  static const bool fuse_all = getenv("PSB_GMRES_NOFUSE_ALL") == nullptr;
  static const bool fuse_cgs2 = getenv("PSB_GMRES_NOFUSE") == nullptr;
  const int grid_dot = one_wave((const void*)gmres_multidot_kernel, &wave_dot);
  const int grid_axpy = one_wave((const void*)gmres_multiaxpy_kernel, &wave_axpy);
  const int grid_mgs = one_wave((const void*)gmres_mgs_kernel, &wave_mgs);
  gmres_init_kernel<<<grid, kBlock, 0, st>>>(w.st, w.sm, n, d_b, w.rb);
  PSB_LAUNCH_CHECK();
  rc = allreduce(&w.st->tmp_dot, 1);
  if (rc != PSB_OK) return rc;
  gmres_init_finish_kernel<<<1, 32, 0, st>>>(w.st, w.sm);
  PSB_LAUNCH_CHECK();
  // q_0 = b / beta   (GMRESSolver.py:90-91); harmless when b == 0 (result unused)
  gmres_scale_kernel<<<grid, kBlock, 0, st>>>(w.st, n, d_b, &w.st->beta, w.Q, 1);
  PSB_LAUNCH_CHECK();

  int slot = 0;
  bool pending[2] = {false, false};
  bool finished = false;
  for (int k = 0; k < maxiter && !finished; ++k) {
    const double* qk = w.Q + (int64_t)k * w.ldq;
    const double* zin = qk;
    if (has_prec) {                                              // z = M^-1 q_k   (:107)
      double* zout = D ? xe : w.z;                               // straight into the extended buffer
      rc = prec->apply(qk, zout, &w.st->done, st);
      if (rc != PSB_OK) return rc;
      zin = zout;
    }
    rc = matvec(EPI_STORE, zin, w.w, EpiArgs(), &w.st->done);    // w = A z
    if (rc != PSB_OK) return rc;
    if (orth == PSB_ORTH_MGS) {
      for (int j = 0; j <= k + 1; ++j) {
        gmres_mgs_kernel<<<grid_mgs, kBlock, 0, st>>>(w.st, w.sm, n, w.Q, w.ldq, w.w, j, k, w.rb);
        PSB_LAUNCH_CHECK();
        rc = allreduce(j == k + 1 ? &w.st->ww : w.sm.hcol + j, 1);
        if (rc != PSB_OK) return rc;
      }
    } else {
      const int j_last = (k / kCh) * kCh;              // first vector of the last chunk
      const bool all_fused = fuse_cgs2 && fuse_all && k + 1 <= kFusedMax;
      if (all_fused) {
        // round-1 dots | all-reduce | [round-1 axpy over all vectors + all round-2 dots] | all-reduce |
        // round-2 axpy (+ ||w||^2): the basis is streamed from HBM three times, not four
        for (int j0 = 0; j0 <= k; j0 += kCh) {
          const int cnt = std::min(kCh, k + 1 - j0);
          gmres_multidot_kernel<<<grid_dot, kBlock, 0, st>>>(w.st, n, w.Q, w.ldq, w.w, j0, cnt, w.sm.hcol, w.rb);
          PSB_LAUNCH_CHECK();
        }
        rc = allreduce(w.sm.hcol, k + 1);
        if (rc != PSB_OK) return rc;
        gmres_cgs2_fused_kernel<<<grid_cgs2, kBlock, kFusedSmem, st>>>(w.st, n, w.Q, w.ldq, w.w, w.sm.hcol, k + 1, w.sm.hcol2, w.rb);
        PSB_LAUNCH_CHECK();
        rc = allreduce(w.sm.hcol2, k + 1);
        if (rc != PSB_OK) return rc;
        for (int j0 = 0; j0 <= k; j0 += kCh) {
          const int cnt = std::min(kCh, k + 1 - j0);
          gmres_multiaxpy_kernel<<<grid_axpy, kBlock, 0, st>>>(w.st, n, w.Q, w.ldq, w.w, w.sm.hcol2, j0, cnt, -1.0,
                                                          0, j0 + kCh > k ? 1 : 0, 1, w.rb);
          PSB_LAUNCH_CHECK();
        }
      }
      for (int round = 0; round < 2 && !all_fused; ++round) {
        double* hout = round == 0 ? w.sm.hcol : w.sm.hcol2;
        for (int j0 = 0; j0 <= k; j0 += kCh) {
          if (round == 1 && fuse_cgs2 && j0 == j_last) continue;   // formed by the fused kernel below
          const int cnt = std::min(kCh, k + 1 - j0);
          gmres_multidot_kernel<<<grid_dot, kBlock, 0, st>>>(w.st, n, w.Q, w.ldq, w.w, j0, cnt, hout, w.rb);
          PSB_LAUNCH_CHECK();
        }
        rc = allreduce(hout, k + 1);
        if (rc != PSB_OK) return rc;
        for (int j0 = 0; j0 <= k; j0 += kCh) {
          const int cnt = std::min(kCh, k + 1 - j0);
          if (round == 0 && fuse_cgs2 && j0 == j_last) {
            // last axpy chunk of round 1 + the round-2 dots of the same chunk in one pass
            gmres_axpy_dot_kernel<<<grid_fused, kBlock, 0, st>>>(w.st, n, w.Q, w.ldq, w.w, hout, j0, cnt,
                                                                w.sm.hcol2, w.rb);
            PSB_LAUNCH_CHECK();
            continue;
          }
          const int want_norm = (round == 1 && j0 + kCh > k) ? 1 : 0;
          gmres_multiaxpy_kernel<<<grid_axpy, kBlock, 0, st>>>(w.st, n, w.Q, w.ldq, w.w, hout, j0, cnt, -1.0,
                                                          0, want_norm, 1, w.rb);
          PSB_LAUNCH_CHECK();
        }
      }
      rc = allreduce(&w.st->ww, 1);
      if (rc != PSB_OK) return rc;
    }
    gmres_givens_kernel<<<1, 32, 0, st>>>(w.st, w.sm, d_hist, orth == PSB_ORTH_CGS2 ? 1 : 0);
    PSB_LAUNCH_CHECK();
    // q_{k+1} = w / h_{k+1,k}  (skipped once done: converged, lucky breakdown or maxiter)
    gmres_scale_kernel<<<grid, kBlock, 0, st>>>(w.st, n, w.w, &w.st->hk1, w.Q + (int64_t)(k + 1) * w.ldq, 1);
    PSB_LAUNCH_CHECK();

    PSB_CUDA(cudaMemcpyAsync(&t_gpoll.pinned[slot], w.st, sizeof(GmresState), cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaEventRecord(t_gpoll.ev[slot], st));
    pending[slot] = true;
    const int prev = slot ^ 1;
    if (pending[prev]) {            // the same logical point on every rank -> the same decision
      PSB_CUDA(cudaEventSynchronize(t_gpoll.ev[prev]));
      pending[prev] = false;
      if (t_gpoll.pinned[prev].done) finished = true;
    }
    slot ^= 1;
  }
  PSB_CUDA(cudaStreamSynchronize(st));
  GmresState hs;
  PSB_CUDA(cudaMemcpy(&hs, w.st, sizeof(hs), cudaMemcpyDeviceToHost));
  if (!hs.done) {
    set_error("psb_gmres_solve: device loop ended without a terminal state (k=%d)", hs.k);
    return PSB_ERR_CUDA;
  }
  result->status = hs.status; result->k = hs.k_final; result->n_hist = hs.n_hist;
  result->lucky = hs.lucky; result->norm_b = hs.norm_b; result->norm_r_rec = hs.norm_r_rec;
  result->norm_r = hs.norm_r_rec;
  if (hs.status == PSB_TRIVIAL) return PSB_OK;

  // ---- x = M^-1 (Q y), true residual (GMRESSolver.py:159-166) ------------------------------
  const int kf = hs.k_final;
  gmres_backsolve_kernel<<<1, 32, 0, st>>>(w.st, w.sm);
  PSB_LAUNCH_CHECK();
  double* tvec = has_prec ? w.t : d_x;
  for (int j0 = 0; j0 <= kf; j0 += kCh) {
    const int cnt = std::min(kCh, kf + 1 - j0);
    gmres_multiaxpy_kernel<<<grid_axpy, kBlock, 0, st>>>(w.st, n, w.Q, w.ldq, tvec, w.sm.y, j0, cnt, 1.0,
                                                    j0 == 0 ? 1 : 0, 0, 0, w.rb);
    PSB_LAUNCH_CHECK();
  }
  if (has_prec) {
    rc = prec->apply(w.t, d_x, nullptr, st);
    if (rc != PSB_OK) return rc;
  }
  EpiArgs ea; ea.f = d_b;
  rc = matvec(EPI_RESID, d_x, w.w, ea, nullptr);                          // r = b - A x
  if (rc != PSB_OK) return rc;
  gmres_norm_kernel<<<grid, kBlock, 0, st>>>(w.st, n, w.w, w.rb);
  PSB_LAUNCH_CHECK();
  rc = allreduce(&w.st->tmp_dot, 1);
  if (rc != PSB_OK) return rc;
  gmres_norm_finish_kernel<<<1, 32, 0, st>>>(w.st);
  PSB_LAUNCH_CHECK();
  PSB_CUDA(cudaStreamSynchronize(st));
  PSB_CUDA(cudaMemcpy(&hs, w.st, sizeof(hs), cudaMemcpyDeviceToHost));
  result->norm_r = hs.norm_true;
  if (hs.status == PSB_CONVERGED && !(hs.norm_true <= tau * hs.norm_b))
    result->status = PSB_GMRES_FALSE_CONV;
  if (has_prec && prec->check_error() != 0) {
    set_error("psb_gmres_solve: the %s preconditioner reported a device-side failure", prec->kind());
    return PSB_ERR_CUDA;
  }
  return PSB_OK;
}

extern "C" int psb_gmres_solve(psb_csr_t A, psb_prec_t prec, const double* d_b, double* d_x,
                               void* d_work, int64_t work_bytes, int32_t maxiter, double tau,
                               int32_t fail_on_maxiter, int32_t orth, double* d_hist,
                               psb_solve_result* result, void* stream) {
  (void)fail_on_maxiter;
  PSB_REQUIRE(A && d_b && d_x && d_work && d_hist && result, PSB_ERR_ARG, "psb_gmres_solve: NULL argument");
  PSB_REQUIRE(A->n_rows == A->n_cols, PSB_ERR_ARG, "psb_gmres_solve: matrix must be square");
  PSB_REQUIRE(maxiter >= 1, PSB_ERR_ARG, "psb_gmres_solve: maxiter must be >= 1");
  PSB_REQUIRE(orth == PSB_ORTH_CGS2 || orth == PSB_ORTH_MGS, PSB_ERR_ARG, "psb_gmres_solve: unknown orth mode");
  PSB_REQUIRE(!prec || prec->n == A->n_rows, PSB_ERR_ARG, "psb_gmres_solve: preconditioner size mismatch");
  PSB_REQUIRE(work_bytes >= psb_gmres_workspace_bytes(A->n_rows, maxiter), PSB_ERR_ARG, "psb_gmres_solve: workspace too small");
  PSB_REQUIRE(aligned16(d_b) && aligned16(d_x) && ((uintptr_t)d_work & 255u) == 0, PSB_ERR_ARG,
              "psb_gmres_solve: b, x must be 16-byte and work 256-byte aligned");
  return gmres_solve_impl(A, nullptr, prec, d_b, d_x, d_work, work_bytes, maxiter, tau, orth, d_hist, result,
                          (cudaStream_t)stream);
}

extern "C" int64_t psb_dist_gmres_workspace_bytes(int64_t n_loc, int64_t n_halo, int32_t maxiter) {
  if (n_loc < 0 || n_halo < 0 || maxiter < 1) return PSB_ERR_ARG;
  return psb_gmres_workspace_bytes(n_loc, maxiter) + align_up((n_loc + n_halo + 32) * (int64_t)sizeof(double), 256);
}

// GMRES on a row-partitioned system: same loop, result codes and history as
// psb_gmres_solve; `prec` (nullable) must act on this rank's slices (e.g. psb_dist_amg_create).
extern "C" int psb_dist_gmres_solve(psb_dist_t D, psb_prec_t prec, const double* d_b_loc, double* d_x_loc,
                                    void* d_work, int64_t work_bytes, int32_t maxiter, double tau,
                                    int32_t fail_on_maxiter, int32_t orth, double* d_hist,
                                    psb_solve_result* result, void* stream) {
  (void)fail_on_maxiter;
  PSB_REQUIRE(D && d_b_loc && d_x_loc && d_work && d_hist && result, PSB_ERR_ARG, "psb_dist_gmres_solve: NULL argument");
  PSB_REQUIRE(dist_n_own(D) == dist_n_loc(D), PSB_ERR_ARG, "psb_dist_gmres_solve: operator must be square");
  PSB_REQUIRE(maxiter >= 1, PSB_ERR_ARG, "psb_dist_gmres_solve: maxiter must be >= 1");
  PSB_REQUIRE(orth == PSB_ORTH_CGS2 || orth == PSB_ORTH_MGS, PSB_ERR_ARG, "psb_dist_gmres_solve: unknown orth mode");
  PSB_REQUIRE(!prec || prec->n == dist_n_loc(D), PSB_ERR_ARG, "psb_dist_gmres_solve: preconditioner size mismatch");
  PSB_REQUIRE(work_bytes >= psb_dist_gmres_workspace_bytes(dist_n_loc(D), dist_n_halo(D), maxiter), PSB_ERR_ARG,
              "psb_dist_gmres_solve: workspace too small");
  PSB_REQUIRE(aligned16(d_b_loc) && aligned16(d_x_loc) && ((uintptr_t)d_work & 255u) == 0, PSB_ERR_ARG,
              "psb_dist_gmres_solve: b, x must be 16-byte and work 256-byte aligned");
  return gmres_solve_impl(nullptr, D, prec, d_b_loc, d_x_loc, d_work, work_bytes, maxiter, tau, orth, d_hist,
                          result, (cudaStream_t)stream);
}
