// Preconditioned conjugate gradients, whole loop resident on the device.
//
// Replaces the Python loop of PySolvers/Linear/PCGSolver.py:97-142.  One
// iteration is the minimum number of HBM passes:
//
//   K1  Ap = A p ,  p.Ap                       spmv.cu (EPI_DOT)   12 nnz + 4 n + 16 n bytes
//   K2  x += a p ; r -= a Ap ; r.r             pcg_update_kernel   48 n bytes
//   [   z = M^-1 r ; z.r                        preconditioner + pcg_zr_kernel ]
//   K3  p = z + b p                            pcg_direction_kernel 24 n bytes
//
// alpha, beta, ||r||, the iteration counter k, the convergence / breakdown flags
// and the residual history all live in device memory (PcgState); the host only
// enqueues chunks of iterations and polls a pinned copy of the state one chunk
// behind, so the GPU never waits for the host.  Once `done` is set every later
// kernel returns immediately, which makes the returned x exactly the iterate at
// the first converged k, as in the reference.
//
// Reductions are deterministic (fixed tree per CTA, per-CTA partials added in
// index order by the last CTA).  No FMA contraction (-fmad=false): x + alpha*p
// rounds the product first, like numpy.
#include "pcg_mega.cuh"
#include "prec.cuh"
#include "spmv.cuh"

#include <algorithm>
#include <cstdlib>

namespace psb {

struct PcgState {
  double udr[2];     // dot(u, r), ping-pong on k parity
  double pAp;
  double rr;
  double norm_b;
  double norm_r;
  double tau;
  int k;             // current iteration (0-based); number of completed iterations
  int maxiter;
  int done;          // != 0 -> every kernel is a no-op
  int status;
  int k_final;
  int fail_on_maxiter;
  int has_prec;
  int n_hist;
};

int stream_grid(int64_t n, int max_grid) {
  int64_t per_cta = (int64_t)kBlock * 4;                 // >= 4 elements per thread
  int64_t g = (n + per_cta - 1) / per_cta;
  g = std::min<int64_t>(g, (int64_t)sm_count() * 8);
  g = std::min<int64_t>(g, max_grid);
  return (int)std::max<int64_t>(g, 1);
}

__device__ __forceinline__ double2 ld2(const double* p) {           // read-write arrays
  double2 v;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
               : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// ---- init: r = b, x = 0, (identity: p = b), b.b ---------------------------------
__global__ void __launch_bounds__(kBlock)
pcg_init_kernel(PcgState* st, int64_t n, const double* __restrict__ b, double* __restrict__ x,
                double* __restrict__ r, double* __restrict__ p, ReduceBuf rb) {
  __shared__ double scratch[kWarps];
  double acc = 0.0;
  const int64_t n2 = n >> 1;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n2;
       i += (int64_t)gridDim.x * kBlock) {
    double2 v = ld_stream2(b + 2 * i);
    st_stream2(r + 2 * i, v);
    st_stream2(x + 2 * i, make_double2(0.0, 0.0));
    if (p != nullptr) st_stream2(p + 2 * i, v);
    acc += v.x * v.x;
    acc += v.y * v.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    double v = b[n - 1];
    r[n - 1] = v; x[n - 1] = 0.0;
    if (p != nullptr) p[n - 1] = v;
    acc += v * v;
  }
  double t = block_sum(acc, scratch);
  if (threadIdx.x == 0) rb.partials[blockIdx.x] = t;
  if (last_block(rb.ticket)) {
    double bb = sum_partials(rb.partials, gridDim.x, scratch);
    if (threadIdx.x == 0) {
      double nb = sqrt(bb);
      st->norm_b = nb;
      if (nb == 0.0) {                      // PCGSolver.py:87-88
        st->done = 1; st->status = PSB_TRIVIAL; st->k_final = 0; st->norm_r = 0.0;
      } else if (!st->has_prec) {
        st->udr[0] = bb;                    // u aliases r: dot(u, r) = dot(b, b) > 0
      }
    }
  }
}

// ---- generic dot with a PCG-specific finish: dot(z, r) --------------------------
// mode 0: before the loop  -> udr[0] = z.r, breakdown if 0  (PCGSolver.py:102-105)
// mode 1: inside the loop  -> udr[(k+1)&1] = z.r ; k += 1 ; maxiter check
__global__ void __launch_bounds__(kBlock)
pcg_zr_kernel(PcgState* st, int64_t n, const double* __restrict__ z, const double* __restrict__ r,
              ReduceBuf rb, int mode) {
  __shared__ double scratch[kWarps];
  if (ld_cg(&st->done) != 0) return;
  double acc = 0.0;
  const int64_t n2 = n >> 1;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n2;
       i += (int64_t)gridDim.x * kBlock) {
    double2 a = ld_stream2(z + 2 * i), c = ld_stream2(r + 2 * i);
    acc += a.x * c.x;
    acc += a.y * c.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) acc += z[n - 1] * r[n - 1];
  double t = block_sum(acc, scratch);
  if (threadIdx.x == 0) rb.partials[blockIdx.x] = t;
  if (last_block(rb.ticket)) {
    double zr = sum_partials(rb.partials, gridDim.x, scratch);
    if (threadIdx.x == 0) {
      if (mode == 0) {
        st->udr[0] = zr;
        if (zr == 0.0) { st->done = 1; st->status = PSB_BREAKDOWN_UR; st->k_final = 0; }
      } else {
        const int k = st->k;
        st->udr[(k + 1) & 1] = zr;
        st->k = k + 1;
        if (k + 1 >= st->maxiter) { st->done = 1; st->status = PSB_MAXITER; st->k_final = k; }
      }
    }
  }
}

// ---- K2: x += alpha p ; r -= alpha Ap ; ||r||, convergence test -----------------
__global__ void __launch_bounds__(kBlock)
pcg_update_kernel(PcgState* st, int64_t n, double* __restrict__ x, const double* __restrict__ p,
                  double* __restrict__ r, const double* __restrict__ Ap,
                  double* __restrict__ hist, ReduceBuf rb) {
  __shared__ double scratch[kWarps];
  if (ld_cg(&st->done) != 0) return;
  const int k = ld_cg(&st->k);
  const double pAp = ld_cg(&st->pAp);
  if (pAp == 0.0) {                           // PCGSolver.py:114-115
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      st->status = PSB_BREAKDOWN_PAP; st->k_final = k; st->done = 1;
    }
    return;
  }
  const double alpha = ld_cg(&st->udr[k & 1]) / pAp;      // :118
  double acc = 0.0;
  const int64_t n2 = n >> 1;
  const int64_t stride = (int64_t)gridDim.x * kBlock;
  int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x;
  // two independent 128-bit chunks per thread per trip: 8 loads in flight
  for (; i + stride < n2; i += 2 * stride) {
    const int64_t j = i + stride;
    double2 x0 = ld2(x + 2 * i), p0 = ld_stream2(p + 2 * i), r0 = ld2(r + 2 * i), a0 = ld_stream2(Ap + 2 * i);
    double2 x1 = ld2(x + 2 * j), p1 = ld_stream2(p + 2 * j), r1 = ld2(r + 2 * j), a1 = ld_stream2(Ap + 2 * j);
    x0.x = x0.x + alpha * p0.x; x0.y = x0.y + alpha * p0.y;
    r0.x = r0.x - alpha * a0.x; r0.y = r0.y - alpha * a0.y;
    x1.x = x1.x + alpha * p1.x; x1.y = x1.y + alpha * p1.y;
    r1.x = r1.x - alpha * a1.x; r1.y = r1.y - alpha * a1.y;
    st_stream2(x + 2 * i, x0); st_stream2(r + 2 * i, r0);
    st_stream2(x + 2 * j, x1); st_stream2(r + 2 * j, r1);
    acc += r0.x * r0.x; acc += r0.y * r0.y;
    acc += r1.x * r1.x; acc += r1.y * r1.y;
  }
  for (; i < n2; i += stride) {
    double2 x0 = ld2(x + 2 * i), p0 = ld_stream2(p + 2 * i), r0 = ld2(r + 2 * i), a0 = ld_stream2(Ap + 2 * i);
    x0.x = x0.x + alpha * p0.x; x0.y = x0.y + alpha * p0.y;
    r0.x = r0.x - alpha * a0.x; r0.y = r0.y - alpha * a0.y;
    st_stream2(x + 2 * i, x0); st_stream2(r + 2 * i, r0);
    acc += r0.x * r0.x; acc += r0.y * r0.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t e = n - 1;
    double xv = x[e] + alpha * p[e];
    double rv = r[e] - alpha * Ap[e];
    x[e] = xv; r[e] = rv;
    acc += rv * rv;
  }
  double t = block_sum(acc, scratch);
  if (threadIdx.x == 0) rb.partials[blockIdx.x] = t;
  if (last_block(rb.ticket)) {
    double rr = sum_partials(rb.partials, gridDim.x, scratch);
    if (threadIdx.x == 0) {
      const double nr = sqrt(rr);                                  // :125
      st->rr = rr; st->norm_r = nr;
      hist[k] = nr;                                                // :126
      st->n_hist = k + 1;
      const bool conv = (nr <= st->tau * st->norm_b) ||
                        (!st->fail_on_maxiter && k == st->maxiter - 1);   // :129-130
      if (conv) {
        st->status = PSB_CONVERGED; st->k_final = k; st->done = 1;
      } else if (!st->has_prec) {
        st->udr[(k + 1) & 1] = rr;          // u aliases r: dot(u, r) = dot(r, r)
        st->k = k + 1;
        if (k + 1 >= st->maxiter) { st->status = PSB_MAXITER; st->k_final = k; st->done = 1; }
      }
    }
  }
}

// ---- K3: p = z + beta p ------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
pcg_direction_kernel(const PcgState* st, int64_t n, const double* __restrict__ z,
                     double* __restrict__ p) {
  if (ld_cg(&st->done) != 0) return;
  const int k = ld_cg(&st->k);                             // already advanced
  const double beta = ld_cg(&st->udr[k & 1]) / ld_cg(&st->udr[(k - 1) & 1]);   // :135
  const int64_t n2 = n >> 1;
  const int64_t stride = (int64_t)gridDim.x * kBlock;
  int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x;
  for (; i + stride < n2; i += 2 * stride) {
    const int64_t j = i + stride;
    double2 z0 = ld_stream2(z + 2 * i), p0 = ld2(p + 2 * i);
    double2 z1 = ld_stream2(z + 2 * j), p1 = ld2(p + 2 * j);
    p0.x = z0.x + beta * p0.x; p0.y = z0.y + beta * p0.y;
    p1.x = z1.x + beta * p1.x; p1.y = z1.y + beta * p1.y;
    st_stream2(p + 2 * i, p0);
    st_stream2(p + 2 * j, p1);
  }
  for (; i < n2; i += stride) {
    double2 z0 = ld_stream2(z + 2 * i), p0 = ld2(p + 2 * i);
    p0.x = z0.x + beta * p0.x; p0.y = z0.y + beta * p0.y;
    st_stream2(p + 2 * i, p0);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) p[n - 1] = z[n - 1] + beta * p[n - 1];
}

// ---- plain deterministic dot (psb_dot) --------------------------------------------
__global__ void __launch_bounds__(kBlock)
dot_kernel(int64_t n, const double* __restrict__ a, const double* __restrict__ b, double* out,
           ReduceBuf rb) {
  __shared__ double scratch[kWarps];
  double acc = 0.0;
  const int64_t n2 = n >> 1;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n2;
       i += (int64_t)gridDim.x * kBlock) {
    double2 u = ld_stream2(a + 2 * i), v = ld_stream2(b + 2 * i);
    acc += u.x * v.x;
    acc += u.y * v.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) acc += a[n - 1] * b[n - 1];
  double t = block_sum(acc, scratch);
  if (threadIdx.x == 0) rb.partials[blockIdx.x] = t;
  if (last_block(rb.ticket)) {
    double s = sum_partials(rb.partials, gridDim.x, scratch);
    if (threadIdx.x == 0) *out = s;
  }
}

// layout of the caller-provided workspace
struct PcgWork {
  PcgState* st;
  ReduceBuf rb;
  double *r, *p, *Ap, *z, *p2;
};

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
static constexpr int64_t kHeaderBytes = 4096;

static int64_t header_bytes() {
  return kHeaderBytes + align_up((int64_t)sm_count() * 16 * sizeof(double), 256);
}

static PcgWork carve(void* d_work, int64_t n, bool has_prec) {
  char* base = (char*)d_work;
  PcgWork w;
  w.st = (PcgState*)base;
  w.rb.ticket = (unsigned int*)(base + 1024);
  w.rb.partials = (double*)(base + kHeaderBytes);
  w.rb.max_grid = sm_count() * 16;
  const int64_t vec = align_up(n * (int64_t)sizeof(double), 256);
  char* v = base + header_bytes();
  w.r = (double*)v;
  w.p = (double*)(v + vec);
  w.Ap = (double*)(v + 2 * vec);
  w.p2 = (double*)(v + 3 * vec);
  w.z = has_prec ? (double*)(v + 4 * vec) : w.r;
  return w;
}

// pinned mirror of the device state, polled one chunk behind the GPU
struct HostPoll {
  PcgState* pinned = nullptr;     // 2 slots
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int init() {
    if (pinned) return PSB_OK;
    PSB_CUDA(cudaHostAlloc((void**)&pinned, 2 * sizeof(PcgState), cudaHostAllocDefault));
    PSB_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    PSB_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    return PSB_OK;
  }
};
static thread_local HostPoll t_poll;

}  // namespace psb

using namespace psb;

extern "C" int64_t psb_pcg_workspace_bytes(int64_t n, int has_prec) {
  if (n < 0) return PSB_ERR_ARG;
  return header_bytes() + (4 + (has_prec ? 1 : 0)) * align_up(n * (int64_t)sizeof(double), 256);
}

extern "C" int psb_dot(int64_t n, const double* d_x, const double* d_y, double* d_out, void* stream) {
  PSB_REQUIRE(n >= 0 && d_out, PSB_ERR_ARG, "psb_dot: bad argument");
  PSB_REQUIRE(n == 0 || (d_x && d_y), PSB_ERR_ARG, "psb_dot: NULL vector");
  PSB_REQUIRE(aligned16(d_x) && aligned16(d_y), PSB_ERR_ARG, "psb_dot: vectors must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  static thread_local ReduceBuf rb;
  if (rb.partials == nullptr) {
    rb.max_grid = sm_count() * 16;
    PSB_CUDA(cudaMalloc(&rb.partials, sizeof(double) * rb.max_grid));
    PSB_CUDA(cudaMalloc(&rb.ticket, sizeof(unsigned int)));
    PSB_CUDA(cudaMemset(rb.ticket, 0, sizeof(unsigned int)));
  }
  dot_kernel<<<stream_grid(n, rb.max_grid), kBlock, 0, st>>>(n, d_x, d_y, d_out, rb);
  PSB_LAUNCH_CHECK();
  return PSB_OK;
}

extern "C" int psb_pcg_solve(psb_csr_t A, psb_prec_t prec, const double* d_b, double* d_x,
                             void* d_work, int64_t work_bytes, int32_t maxiter, double tau,
                             int32_t fail_on_maxiter, double* d_hist, psb_solve_result* result,
                             void* stream) {
  PSB_REQUIRE(A && d_b && d_x && d_work && d_hist && result, PSB_ERR_ARG, "psb_pcg_solve: NULL argument");
  PSB_REQUIRE(A->n_rows == A->n_cols, PSB_ERR_ARG, "psb_pcg_solve: matrix must be square");
  PSB_REQUIRE(maxiter >= 1, PSB_ERR_ARG, "psb_pcg_solve: maxiter must be >= 1");
  const int64_t n = A->n_rows;
  const bool has_prec = prec != nullptr;
  PSB_REQUIRE(!has_prec || prec->n == n, PSB_ERR_ARG, "psb_pcg_solve: preconditioner size mismatch");
  PSB_REQUIRE(work_bytes >= psb_pcg_workspace_bytes(n, has_prec), PSB_ERR_ARG,
              "psb_pcg_solve: workspace too small");
  PSB_REQUIRE(aligned16(d_b) && aligned16(d_x) && ((uintptr_t)d_work & 255u) == 0, PSB_ERR_ARG,
              "psb_pcg_solve: b, x must be 16-byte and work 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = t_poll.init();
  if (rc != PSB_OK) return rc;

  PcgWork w = carve(d_work, n, has_prec);
  PSB_CUDA(cudaMemsetAsync(d_work, 0, kHeaderBytes, st));

  // ---- identity preconditioner + STREAM matrix: the whole solve is ONE persistent cooperative
  // kernel (pcg_mega.cu); the reductions double as grid barriers, no launch per phase ----------
  {
    const char* env = getenv("PSB_PCG_MEGA");
    const bool mega = !has_prec && A->kind == PSB_SPMV_STREAM && A->rpt == 1 && n > 0 &&
                      !(env && env[0] == '0');
    if (mega) {
      char* base = (char*)d_work;
      MegaParams P;
      memset(&P, 0, sizeof(P));
      P.n = n; P.n_halo = 0;
      P.b = d_b; P.x = d_x; P.r = w.r; P.Ap = w.Ap; P.pbuf[0] = w.p; P.pbuf[1] = w.p2;
      P.hist = d_hist;
      P.st = (MegaState*)base;
      P.ticket = (unsigned int*)(base + 1024);
      P.partials = w.rb.partials;
      unsigned long long* slots = (unsigned long long*)(base + 2048);
      unsigned long long** d_ptrs = (unsigned long long**)(base + 1536);
      unsigned long long* h_ptrs[kRing];
      for (int e = 0; e < kRing; ++e) h_ptrs[e] = slots + (size_t)e * kSlotWords;   // one rank: a line per ring entry
      PSB_CUDA(cudaMemcpyAsync(d_ptrs, h_ptrs, sizeof(h_ptrs), cudaMemcpyHostToDevice, st));
      PSB_CUDA(cudaStreamSynchronize(st));                  // h_ptrs is on the stack
      P.my_slots = slots; P.slot_ptrs = d_ptrs; P.nranks = 1; P.epoch0 = 1; P.ring_words = kSlotWords;
      P.n_push = 0; P.n_wait = 0; P.halo_epoch0 = 1;
      P.int_r0 = P.int_r1 = 0;
      P.maxiter = maxiter; P.tau = tau; P.fail_on_maxiter = fail_on_maxiter;
      P.error = (int*)(base + 1024 + 128);
      rc = pcg_mega_launch(P, A, st);
      if (rc == PSB_OK) {
      PSB_CUDA(cudaStreamSynchronize(st));
      MegaState ms;
      PSB_CUDA(cudaMemcpy(&ms, P.st, sizeof(ms), cudaMemcpyDeviceToHost));
      if (!ms.done) { set_error("psb_pcg_solve: persistent kernel ended without a terminal state"); return PSB_ERR_CUDA; }
      result->status = ms.status; result->k = ms.k_final; result->n_hist = ms.n_hist; result->lucky = 0;
      result->norm_r = ms.norm_r; result->norm_b = ms.norm_b; result->norm_r_rec = ms.norm_r;
      return PSB_OK;
      }
      // the cooperative launch was refused (e.g. the device is shared and the grid cannot be
      // co-resident): clear the error and run the kernel-per-phase driver below -- still on the GPU
      (void)cudaGetLastError();
      PSB_CUDA(cudaMemsetAsync(d_work, 0, kHeaderBytes, st));
    }
  }
  PcgState h0;
  memset(&h0, 0, sizeof(h0));
  h0.tau = tau; h0.maxiter = maxiter; h0.fail_on_maxiter = fail_on_maxiter; h0.has_prec = has_prec;
  // pageable source is fine: the copy is tiny and ordered on the stream
  PSB_CUDA(cudaMemcpyAsync(w.st, &h0, sizeof(h0), cudaMemcpyHostToDevice, st));
  PSB_CUDA(cudaStreamSynchronize(st));     // h0 is on the stack

  const int grid = stream_grid(n, w.rb.max_grid);
  pcg_init_kernel<<<grid, kBlock, 0, st>>>(w.st, n, d_b, d_x, w.r, has_prec ? nullptr : w.p, w.rb);
  PSB_LAUNCH_CHECK();
  if (has_prec) {
    rc = prec->apply(w.r, w.p, &w.st->done, st);           // p = M^-1 r   (PCGSolver.py:98)
    if (rc != PSB_OK) return rc;
    pcg_zr_kernel<<<grid, kBlock, 0, st>>>(w.st, n, w.p, w.r, w.rb, 0);
    PSB_LAUNCH_CHECK();
  }

  // chunked enqueue; the host looks at chunk c-1 while chunk c runs
  const int chunk = 32;
  int enq = 0, slot = 0;
  bool pending[2] = {false, false};
  bool finished = false;
  EpiArgs ea; ea.dot = &w.st->pAp;
  // p ping-pong for the fused direction update (STREAM kernel); otherwise p is updated in place
  const bool fuse = (A->kind == PSB_SPMV_STREAM) && (getenv("PSB_PCG_NOFUSE") == nullptr);
  double* pbuf[2] = {w.p, fuse ? w.p2 : w.p};
  while (!finished) {
    const int todo = std::min(chunk, maxiter - enq);
    for (int i = 0; i < todo; ++i) {
      const int it = enq + i;                       // == device k while the solve is running
      double* pcur = pbuf[it & 1];
      if (fuse && it > 0) {
        // K3 folded into K1: p_k = z + beta p_{k-1} is formed on the fly by the SpMV (gathers
        // and own row, same rounding as the separate kernel), stored once, never re-read.
        EpiArgs eb = ea;
        eb.pold = pbuf[(it - 1) & 1]; eb.pnew = pcur;
        eb.beta_num = &w.st->udr[it & 1]; eb.beta_den = &w.st->udr[(it - 1) & 1];
        rc = spmv_launch(A, EPI_DOT_PUP, w.z, w.Ap, eb, &w.st->done, st);
      } else {
        rc = spmv_launch(A, EPI_DOT, pcur, w.Ap, ea, &w.st->done, st);
      }
      if (rc != PSB_OK) return rc;
      pcg_update_kernel<<<grid, kBlock, 0, st>>>(w.st, n, d_x, pcur, w.r, w.Ap, d_hist, w.rb);
      PSB_LAUNCH_CHECK();
      if (has_prec) {
        rc = prec->apply(w.r, w.z, &w.st->done, st);
        if (rc != PSB_OK) return rc;
        pcg_zr_kernel<<<grid, kBlock, 0, st>>>(w.st, n, w.z, w.r, w.rb, 1);
        PSB_LAUNCH_CHECK();
      }
      if (!fuse) {
        pcg_direction_kernel<<<grid, kBlock, 0, st>>>(w.st, n, w.z, w.p);
        PSB_LAUNCH_CHECK();
      }
    }
    enq += todo;
    PSB_CUDA(cudaMemcpyAsync(&t_poll.pinned[slot], w.st, sizeof(PcgState), cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaEventRecord(t_poll.ev[slot], st));
    pending[slot] = true;
    const int prev = slot ^ 1;
    if (pending[prev]) {
      PSB_CUDA(cudaEventSynchronize(t_poll.ev[prev]));
      pending[prev] = false;
      if (t_poll.pinned[prev].done) finished = true;
    }
    if (enq >= maxiter) finished = true;
    slot ^= 1;
  }
  PSB_CUDA(cudaStreamSynchronize(st));
  PcgState hs;
  PSB_CUDA(cudaMemcpy(&hs, w.st, sizeof(hs), cudaMemcpyDeviceToHost));
  if (!hs.done) {
    set_error("psb_pcg_solve: device loop ended without a terminal state (k=%d)", hs.k);
    return PSB_ERR_CUDA;
  }
  if (has_prec && prec->check_error() != 0) {
    set_error("psb_pcg_solve: the %s preconditioner reported a device-side failure "
              "(a triangular-solve dependency never became ready)", prec->kind());
    return PSB_ERR_CUDA;
  }
  result->status = hs.status;
  result->k = hs.k_final;
  result->n_hist = hs.n_hist;
  result->lucky = 0;
  result->norm_r = hs.norm_r;
  result->norm_b = hs.norm_b;
  result->norm_r_rec = hs.norm_r;
  return PSB_OK;
}
