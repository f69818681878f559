// Un-preconditioned PCG as ONE persistent cooperative kernel per GPU.
//
// The loop of PySolvers/Linear/PCGSolver.py:97-142 has two global reductions per iteration
// (p.Ap and r.r); each of them is a grid-wide -- and, row-partitioned, machine-wide --
// synchronisation point.  Instead of ending a kernel at each of them, every CTA stays resident
// and the all-reduce itself is the barrier:
//
//   phase A   p = r + beta p_old formed on the fly (gathers + own rows), Ap = A p, CTA partial of
//             p.Ap  (bulk-async staged STREAM SpMV, spmv_bulk.cuh); new boundary rows of p are
//             also stored into the neighbours' halo (NVLink peer stores)
//   reduce    CTA partials -> last CTA (ticket) -> fixed-order sum -> epoch-tagged store of the
//             rank's value into its slot in EVERY rank's memory; all CTAs of all ranks poll their
//             LOCAL slots and add them in rank order (same bits everywhere)
//   phase B   x += alpha p ; r -= alpha Ap ; CTA partial of r.r ; boundary entries of r are
//             stored into the neighbours' halo, then their halo flag is raised
//   reduce    as above; convergence test (PCGSolver.py:125-131) taken identically by every CTA
//
// alpha, beta, r.r live in registers (every CTA derives the same values from the same slots);
// kernel-launch boundaries, their drain/fill bubbles and the NCCL launches are gone.  The
// acquire side of each reduce is a gpu-scope fence, which also invalidates L1 so that the
// cached x gathers of the next phase see what other CTAs / GPUs wrote.
#include "pcg_mega.cuh"
#include "prec.cuh"
#include "spmv_bulk.cuh"

#include <algorithm>
#include <cstdlib>

namespace psb {

// Grid-wide (and rank-wide) sum of `v`.  Every thread of every CTA calls it; all return the same
// bits.  `raise`: the calling phase stored halo data into the neighbours' memory -- publish it
// (system fence by every thread, flags by the last CTA).
__device__ __forceinline__ double mega_allreduce(const MegaParams& P, double v, unsigned int epoch,
                                                 double* scratch, bool raise,
                                                 unsigned long long halo_epoch) {
  __shared__ double s_sum;
  const double t = block_sum(v, scratch);
  if (threadIdx.x == 0) P.partials[blockIdx.x] = t;
  if (raise && P.n_push > 0) __threadfence_system();
  if (last_block(P.ticket)) {
    const double s = sum_partials(P.partials, gridDim.x, scratch);
    if (threadIdx.x == 0)
      for (int q = 0; q < P.nranks; ++q)
        peer_push(P.slot_ptrs[(size_t)(epoch % kRing) * P.nranks + q], s, epoch);
    if (raise && threadIdx.x < P.n_push) {
      __threadfence_system();
      asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(P.push_flag[threadIdx.x]), "l"(halo_epoch) : "memory");
    }
  }
  if (threadIdx.x < 32) {                       // one lane per rank polls, then a fixed-order sum
    double mine = 0.0;
    if ((int)threadIdx.x < P.nranks) {
      if (!peer_wait(P.my_slots + ((size_t)(epoch % kRing) * kMaxRanks + threadIdx.x) * 2, epoch, &mine))
        *P.error = 1;
    }
    double s = 0.0;
    for (int q = 0; q < P.nranks; ++q) s += __shfl_sync(0xffffffffu, mine, q);
    if (threadIdx.x == 0) s_sum = s;
  }
  __syncthreads();
  __threadfence();                              // acquire; invalidates L1 (weak loads below see fresh data)
  return s_sum;
}

__device__ __forceinline__ double2 mega_ld2(const double* p) {     // coherent, no L1 allocation
  double2 v;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

__device__ __forceinline__ void mega_push_r(const MegaParams& P, long long i, double v) {
#pragma unroll
  for (int k = 0; k < kMaxPush; ++k)
    if (k < P.n_push && i >= P.push_off[k] && i < P.push_off[k] + P.push_cnt[k])
      P.push_r[k][i - P.push_off[k]] = v;
}

template <int MINB, bool C16>
__global__ void __launch_bounds__(kBlock, MINB)
pcg_mega_kernel(const MegaParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double scratch[kWarps];
  __shared__ __align__(8) uint64_t full[2];
  BulkPipe pipe;
  bulk_pipe_init(pipe, smem_raw, full, P.cap_v, P.cap_c);

  const long long n = P.n;
  const long long gtid = blockIdx.x * (long long)kBlock + threadIdx.x;
  const long long gstride = (long long)gridDim.x * kBlock;
  const bool leader = blockIdx.x == 0 && threadIdx.x == 0;
  unsigned int e = P.epoch0;
  const unsigned long long h0 = P.halo_epoch0;

  // ---- init: r = b, x = 0, p_{-1} = 0 (so that p_0 = r + 0 * p_{-1}), b.b  (PCGSolver.py:97-102)
  double acc = 0.0;
  for (long long i = gtid; i < n; i += gstride) {
    const double v = P.b[i];
    P.r[i] = v; P.x[i] = 0.0; P.pbuf[1][i] = 0.0;
    mega_push_r(P, i, v);
    acc += v * v;
  }
  for (long long i = gtid; i < P.n_halo; i += gstride) P.pbuf[1][n + i] = 0.0;
  const double bb = mega_allreduce(P, acc, e++, scratch, true, h0);
  const double norm_b = sqrt(bb);
  int status = PSB_MAXITER, k_final = 0, n_hist = 0;
  double norm_r = 0.0;
  if (bb == 0.0) {                                          // :87-88
    status = PSB_TRIVIAL;
  } else {
    double rr_old = bb;                                     // dot(u, r) with u = r
    double beta = 0.0;
    for (int it = 0;; ++it) {
      if (*((volatile int*)P.error) != 0) { status = PSB_MAXITER; k_final = it; break; }
      // ---------------- phase A: p = r + beta p_old ; Ap = A p ; p.Ap -------------------------
      EpiArgs ea;
      ea.pold = P.pbuf[(it + 1) & 1];
      ea.pnew = P.pbuf[it & 1];
      ea.rot_t0 = P.rot_t0; ea.rot_t1 = P.rot_t1;
      ea.error_flag = P.error;
      if (P.n_wait > 0) { ea.wait_flags = P.my_flags; ea.wait_n = P.n_wait; ea.wait_value = h0 + (unsigned long long)it; }
      ea.pp_n = P.n_push;
#pragma unroll
      for (int k = 0; k < kMaxPush; ++k) {
        ea.pp_off[k] = P.push_off[k]; ea.pp_cnt[k] = P.push_cnt[k]; ea.pp_remote[k] = P.push_p[it & 1][k];
      }
      acc = 0.0;
      bulk_pass<EPI_DOT_PUP, 1, C16>(P.A, P.r, P.Ap, ea, beta, pipe, acc);
      const double pAp = mega_allreduce(P, acc, e++, scratch, false, 0ull);
      if (pAp == 0.0) { status = PSB_BREAKDOWN_PAP; k_final = it; break; }     // :114-115
      const double alpha = rr_old / pAp;                                         // :118
      // ---------------- phase B: x += alpha p ; r -= alpha Ap ; r.r ------------------------------
      const double* p = P.pbuf[it & 1];
      acc = 0.0;
      const long long n2 = n >> 1;
      // two independent 128-bit chunks per thread per trip, streaming (no L1 allocation)
      long long i = gtid;
      for (; i + gstride < n2; i += 2 * gstride) {
        const long long j = i + gstride;
        double2 x0 = mega_ld2(P.x + 2 * i), p0 = mega_ld2(p + 2 * i), r0 = mega_ld2(P.r + 2 * i), a0 = mega_ld2(P.Ap + 2 * i);
        double2 x1 = mega_ld2(P.x + 2 * j), p1 = mega_ld2(p + 2 * j), r1 = mega_ld2(P.r + 2 * j), a1 = mega_ld2(P.Ap + 2 * j);
        x0.x = x0.x + alpha * p0.x; x0.y = x0.y + alpha * p0.y;
        r0.x = r0.x - alpha * a0.x; r0.y = r0.y - alpha * a0.y;
        x1.x = x1.x + alpha * p1.x; x1.y = x1.y + alpha * p1.y;
        r1.x = r1.x - alpha * a1.x; r1.y = r1.y - alpha * a1.y;
        st_stream2(P.x + 2 * i, x0); st_stream2(P.r + 2 * i, r0);
        st_stream2(P.x + 2 * j, x1); st_stream2(P.r + 2 * j, r1);
        if (P.n_push > 0) {
          mega_push_r(P, 2 * i, r0.x); mega_push_r(P, 2 * i + 1, r0.y);
          mega_push_r(P, 2 * j, r1.x); mega_push_r(P, 2 * j + 1, r1.y);
        }
        acc += r0.x * r0.x; acc += r0.y * r0.y;
        acc += r1.x * r1.x; acc += r1.y * r1.y;
      }
      for (; i < n2; i += gstride) {
        double2 x0 = mega_ld2(P.x + 2 * i), p0 = mega_ld2(p + 2 * i), r0 = mega_ld2(P.r + 2 * i), a0 = mega_ld2(P.Ap + 2 * i);
        x0.x = x0.x + alpha * p0.x; x0.y = x0.y + alpha * p0.y;
        r0.x = r0.x - alpha * a0.x; r0.y = r0.y - alpha * a0.y;
        st_stream2(P.x + 2 * i, x0); st_stream2(P.r + 2 * i, r0);
        if (P.n_push > 0) { mega_push_r(P, 2 * i, r0.x); mega_push_r(P, 2 * i + 1, r0.y); }
        acc += r0.x * r0.x; acc += r0.y * r0.y;
      }
      if ((n & 1) && leader) {
        const long long i = n - 1;
        const double xv = P.x[i] + alpha * p[i];
        const double rv = P.r[i] - alpha * P.Ap[i];
        P.x[i] = xv; P.r[i] = rv;
        mega_push_r(P, i, rv);
        acc += rv * rv;
      }
      const double rr = mega_allreduce(P, acc, e++, scratch, true, h0 + (unsigned long long)it + 1ull);
      norm_r = sqrt(rr);                                                         // :125
      if (leader) P.hist[it] = norm_r;                                           // :126
      n_hist = it + 1;
      if ((norm_r <= P.tau * norm_b) || (!P.fail_on_maxiter && it == P.maxiter - 1)) {   // :129-131
        status = PSB_CONVERGED; k_final = it; break;
      }
      if (it + 1 >= P.maxiter) { status = PSB_MAXITER; k_final = it; break; }
      beta = rr / rr_old;                                                        // :135
      rr_old = rr;
    }
  }
  if (leader) {
    P.st->norm_b = norm_b; P.st->norm_r = norm_r;
    P.st->status = status; P.st->k_final = k_final; P.st->n_hist = n_hist; P.st->done = 1;
    P.st->epochs_used = e - P.epoch0;
    P.st->halo_epochs_used = (unsigned int)n_hist + 1u;
  }
}

void pcg_mega_caps(const psb_csr* A, int* cap_v, int* cap_c, size_t* smem) {
  const int mt = A->max_tile_nnz[0];
  *cap_v = (mt + 2 + 1) & ~1;
  *cap_c = (mt + 6 + 3) & ~3;
  *smem = 2 * ((size_t)*cap_v * 8 + (size_t)*cap_c * 4 + (size_t)(kBlock + 4) * 4);
}

template <int MINB, bool C16>
static int mega_launch_t(const MegaParams& P, cudaStream_t stream) {
  int cv, cc;
  size_t smem;
  pcg_mega_caps(&P.A, &cv, &cc, &smem);
  static thread_local int per_sm = 0;
  static thread_local size_t cached_smem = 0;
  if (per_sm == 0 || cached_smem != smem) {
    if (smem > 48 * 1024)
      PSB_CUDA(cudaFuncSetAttribute(pcg_mega_kernel<MINB, C16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_mega_kernel<MINB, C16>, kBlock, smem));
    if (per_sm < 1) { set_error("pcg_mega_launch: kernel does not fit on an SM"); return PSB_ERR_UNSUPP; }
    cached_smem = smem;
  }
  const long long tiles = (P.A.n_rows + kBlock - 1) / kBlock;
  const long long want = std::max<long long>(tiles, (P.n + kBlock * 8 - 1) / (kBlock * 8));
  long long grid = std::min<long long>((long long)per_sm * sm_count(), std::max<long long>(want, 1));
  grid = std::min<long long>(grid, (long long)sm_count() * 16);       // partial buffers hold this many
  MegaParams Q = P;
  void* args[] = {(void*)&Q};
  PSB_CUDA(cudaLaunchCooperativeKernel((const void*)pcg_mega_kernel<MINB, C16>, dim3((unsigned)grid), dim3(kBlock),
                                       args, smem, stream));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PSB_OK;
}

int pcg_mega_launch(const MegaParams& P, cudaStream_t stream) {
  // resident CTAs per SM the kernel is compiled for: 4 (64 registers), 5 (48) or 6 (40)
  static int minb = 0;
  if (minb == 0) {
    const char* env = getenv("PSB_MEGA_MINB");
    minb = env ? atoi(env) : 5;       // measured best on B200: 5 CTAs/SM, 48 registers, no spills
    if (minb < 4 || minb > 6) minb = 5;
  }
  if (P.A.colind16 != nullptr) {
    if (minb == 6) return mega_launch_t<6, true>(P, stream);
    if (minb == 4) return mega_launch_t<4, true>(P, stream);
    return mega_launch_t<5, true>(P, stream);
  }
  if (minb == 6) return mega_launch_t<6, false>(P, stream);
  if (minb == 4) return mega_launch_t<4, false>(P, stream);
  return mega_launch_t<5, false>(P, stream);
}

}  // namespace psb
