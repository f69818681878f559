// Un-preconditioned PCG as ONE persistent cooperative kernel per GPU.
//
// The loop of PySolvers/Linear/PCGSolver.py:97-142 has two global reductions per iteration
// (p.Ap and r.r); each of them is a grid-wide -- and, row-partitioned, machine-wide --
// synchronisation point.  Instead of ending a kernel at each of them, every CTA stays resident
// and the all-reduce itself is the barrier:
//
//   phase A   p = r + beta p_old formed on the fly (gathers + own rows), Ap = A p, CTA partial of
//             p.Ap  (bulk-async staged STREAM SpMV, spmv_bulk.cuh); on its own rows the CTA also
//             applies the PREVIOUS iteration's solution update x += alpha_{k-1} p_{k-1} (p_{k-1}
//             is in a register there anyway); new boundary rows of p are also stored into the
//             neighbours' halo (NVLink peer stores)
//   reduce    CTA partials -> last CTA (ticket) -> fixed-order sum -> epoch-tagged store of the
//             rank's value into its slot in EVERY rank's memory; all CTAs of all ranks poll their
//             LOCAL slots and add them in rank order (same bits everywhere)
//   phase B   r -= alpha Ap ; CTA partial of r.r, on the SAME rows the CTA owned in phase A (its
//             Ap values were written by the very same threads); boundary entries of r are stored
//             into the neighbours' halo, then their halo flag is raised
//   reduce    as above; convergence test (PCGSolver.py:125-131) taken identically by every CTA
//
// Vector traffic per iteration: r, p_old, x read + p, Ap, x written (A) and r, Ap read + r written
// (B) = 72 n bytes (SURVEY.md section 8d counts 88 n for the three-kernel form).  The last
// x += alpha p is applied when the loop ends.
//
// Latency of the two barriers is what limits strong scaling (profiles/round2_mega_timeline.md), so
// everything that does not depend on the reduced scalar is put in flight BEFORE the CTA polls: the
// bulk copies of its first SpMV tile (the matrix never changes) and the loads of its first
// phase-B rows.  Rows per tile are chosen per launch so that every CTA gets the same number of
// tiles (no tail round).
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

__device__ __forceinline__ unsigned long long mega_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ double mega_sum_partials(const double* partials, int count, double* scratch);
constexpr unsigned long long kMegaFailBits = 0x7ff4dead00000001ull;   // signalling NaN: "a peer never answered"

// Grid-wide (and rank-wide) sum of `v`.  Every thread of every CTA calls it; all return the same
// bits.  `pushed`: this thread stored halo data into a neighbour's memory since the last raise
// (it orders those stores system-wide before its CTA takes the ticket).  `raise`: the last CTA
// publishes the halo by raising the neighbours' flags to `halo_epoch`.  `push_base` (lanes
// 0..nranks-1 of warp 0): this rank's ring-0 slot in rank <lane>'s memory.  Returns false in
// *ok when a peer never answered.
__device__ __forceinline__ double mega_allreduce(const MegaParams& P, double v, unsigned int epoch,
                                                 double* scratch, bool pushed, bool raise,
                                                 unsigned long long halo_epoch,
                                                 unsigned long long* push_base, bool* ok) {
  __shared__ double s_sum;
  __shared__ int s_ok;
  const double t = block_sum(v, scratch);
  if (threadIdx.x == 0) P.partials[blockIdx.x] = t;
  // Halo stores into the neighbours' memory are published by ONE system-scope fence of the last CTA
  // (below), not by a fence in every pushing thread: the pushing thread's stores happen before its
  // gpu-scope fence and ticket (last_block), the last CTA observes every ticket, and its
  // fence.sys is cumulative over everything it has observed (PTX memory model; the same pattern as
  // "bar.sync, then one thread fences and raises the flag" inside a CTA).  With the system fence in
  // the pushing threads the r.r reduce took 13 - 15 us at 8 GPUs against 6.5 us for p.Ap
  // (profiles/round2_mega_timeline.md); PSB_MEGA_FLAGS=4 restores it for A/B runs.
  if (pushed && (P.dbg_flags & 4)) __threadfence_system();
  const size_t ring = (size_t)(epoch % kRing) * (size_t)P.ring_words;
  const bool last = last_block(P.ticket);
  // Two-stage release for more than two ranks: only the LAST CTA of each rank polls the ranks' slots
  // (8 lanes instead of 8 lanes x every CTA: with all of them polling, a poll trip through L2 took
  // several microseconds and the reduce 9 - 14 us at 8 GPUs), adds them in rank order and stores the
  // total into a local broadcast slot; every other CTA polls that one slot.
  const bool two_stage = P.nranks > 2;
  const size_t bcast = ring + (size_t)(kMaxRanks - 1) * kSlotWords;
  if (last) {
    const double s = mega_sum_partials(P.partials, gridDim.x, scratch);
    if ((int)threadIdx.x < P.nranks) peer_push(push_base + ring, s, epoch);
    if (two_stage && threadIdx.x < 32) {
      double mine = 0.0;
      bool good = true;
      if ((int)threadIdx.x < P.nranks) good = peer_wait(P.my_slots + ring + threadIdx.x * kSlotWords, epoch, &mine);
      double tot = 0.0;
      for (int q = 0; q < P.nranks; ++q) tot += __shfl_sync(0xffffffffu, mine, q);
      const bool all_good = __all_sync(0xffffffffu, good);
      if (threadIdx.x == 0) {
        if (!all_good) *P.error = 1;
        // a time-out still releases the other CTAs, with a payload no arithmetic produces
        peer_push(const_cast<unsigned long long*>(P.my_slots) + bcast,
                  all_good ? tot : __longlong_as_double((long long)kMegaFailBits), epoch);
      }
    }
  }
  if (threadIdx.x < 32) {                       // one lane per rank polls, then a fixed-order sum
    double s = 0.0;
    bool all_good = true;
    if (two_stage) {
      double tot = 0.0;
      bool good = true;
      if (threadIdx.x == 0) good = peer_wait(P.my_slots + bcast, epoch, &tot);
      s = __shfl_sync(0xffffffffu, tot, 0);
      all_good = __all_sync(0xffffffffu, good) && (unsigned long long)__double_as_longlong(s) != kMegaFailBits;
    } else {
      double mine = 0.0;
      bool good = true;
      if ((int)threadIdx.x < P.nranks) good = peer_wait(P.my_slots + ring + threadIdx.x * kSlotWords, epoch, &mine);
      for (int q = 0; q < P.nranks; ++q) s += __shfl_sync(0xffffffffu, mine, q);
      all_good = __all_sync(0xffffffffu, good);
    }
    if (threadIdx.x == 0) { s_sum = s; s_ok = all_good ? 1 : 0; if (!all_good) *P.error = 1; }
  }
  // The halo flags are raised AFTER the poll: the neighbour needs them only when it reaches its
  // first boundary tile (its last tiles), and by now this thread's scalar push has landed, so the
  // system fence has nothing remote left to wait for (before the poll it cost 3 - 4 us of the last
  // CTA's -- i.e. the critical -- path, profiles/round2_mega_timeline.md).
  if (last && raise && (int)threadIdx.x < P.n_push) {
    __threadfence_system();
    asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(P.push_flag[threadIdx.x]), "l"(halo_epoch) : "memory");
  }
  __syncthreads();
  __threadfence();                              // acquire; invalidates L1 (weak loads below see fresh data)
  *ok = s_ok != 0;
  return s_sum;
}

// Per-CTA partials summed in a fixed order by the last CTA: up to 4 per thread, all four loads in
// flight at once (the generic sum_partials loop issues them one dependent trip at a time).
__device__ __forceinline__ double mega_sum_partials(const double* partials, int count, double* scratch) {
  double v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = threadIdx.x + j * kBlock;
    v[j] = i < count ? ld_cg(partials + i) : 0.0;
  }
  double a = ((v[0] + v[1]) + v[2]) + v[3];
  for (int i = threadIdx.x + 4 * kBlock; i < count; i += kBlock) a += ld_cg(partials + i);
  return block_sum(a, scratch);
}

__device__ __forceinline__ double mega_ld(const double* p) {       // coherent, no L1 allocation
  double v;
  asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void mega_st(double* p, double v) {
  asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}

__device__ __forceinline__ bool mega_push_r(const MegaParams& P, int i, double v) {
  bool did = false;
#pragma unroll
  for (int k = 0; k < kMaxPush; ++k)
    if (k < P.n_push && i >= P.push_off[k] && i < P.push_off[k] + P.push_cnt[k]) {
      P.push_r[k][i - P.push_off[k]] = v;
      did = true;
    }
  return did;
}

template <int MINB, bool C16, int G>
__global__ void __launch_bounds__(kBlock, MINB)
pcg_mega_kernel(const MegaParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double scratch[kWarps];
  __shared__ __align__(8) BulkShared bsh;
  constexpr int kMegaUnroll = MINB <= 4 ? 4 : 2;   // phase-B tiles whose loads are in flight together
  BulkPipe pipe;
  bulk_pipe_init(pipe, smem_raw, &bsh, P.cap_v, P.cap_c);
  pipe.tile_rows = P.tile_rows;

  const int n = (int)P.n;                                // rows fit 31 bits (psb_csr_create)
  const int tid = threadIdx.x;
  const int R = P.tile_rows;
  const int n_tiles = (n + R - 1) / R;
  const int rot_t0 = (int)P.rot_t0, rot_t1 = (int)P.rot_t1;
  const int n_int = rot_t1 - rot_t0;
  const int gstep = gridDim.x;
  // my tiles: logical lt = blockIdx.x + j * gridDim.x, interior tiles first (as in bulk_pass)
  auto tile_row = [&](int lt) -> int {
    const int t = lt < n_int ? rot_t0 + lt : (lt < rot_t1 ? lt - n_int : lt);
    return t * R + tid;                              // this thread's row of the tile
  };
  auto row_ok = [&](int lt, int row) -> bool { return lt >= 0 && lt < n_tiles && tid < R && row < n; };
  // this CTA's last logical tile (-1: none)
  const int lt_last = (int)blockIdx.x < n_tiles ? (int)blockIdx.x + ((n_tiles - 1 - (int)blockIdx.x) / gstep) * gstep : -1;
  const bool leader = blockIdx.x == 0 && tid == 0;
  unsigned int e = P.epoch0;
  const unsigned long long h0 = P.halo_epoch0;
  unsigned long long* push_base = nullptr;
  if (tid < P.nranks) push_base = P.slot_ptrs[tid];          // ring 0; ring k is k * ring_words further

  // constant part of the SpMV arguments; the bounds of the first two tiles are cached
  EpiArgs ea;
  ea.rot_t0 = P.rot_t0; ea.rot_t1 = P.rot_t1;
  ea.error_flag = P.error;
  ea.pp_n = P.n_push;
#pragma unroll
  for (int k = 0; k < kMaxPush; ++k) { ea.pp_off[k] = P.push_off[k]; ea.pp_cnt[k] = P.push_cnt[k]; }
  ea.pp_skip_lo = P.nopush_lo; ea.pp_skip_hi = P.nopush_hi;
  ea.xsol = P.x;
  bulk_cache_bounds<1, C16>(P.A, ea, pipe);

  unsigned long long* tl = nullptr;                          // profiling: 6 stamps per CTA per iteration
  auto stamp = [&](int it, int slot) {
    if (tl != nullptr && tid == 0 && it >= P.tl_first && it < P.tl_first + P.tl_count)
      tl[((size_t)(it - P.tl_first) * gridDim.x + blockIdx.x) * 6 + slot] = mega_now();
  };
  tl = P.timeline;

  // ---- init: r = b, x = 0, p_{-1} = 0 (so that p_0 = r + 0 * p_{-1}), b.b  (PCGSolver.py:97-102)
  double acc = 0.0;
  bool pushed = false, ok = true;
  for (int lt = blockIdx.x; lt < n_tiles; lt += gstep) {
    const int row = tile_row(lt);
    if (row_ok(lt, row)) {
      const double v = P.b[row];
      P.r[row] = v; P.x[row] = 0.0; P.pbuf[1][row] = 0.0;
      if (P.n_push > 0 && (row < P.nopush_lo || row >= P.nopush_hi)) pushed |= mega_push_r(P, row, v);
      acc += v * v;
    }
  }
  for (int i = blockIdx.x * kBlock + tid; i < (int)P.n_halo; i += gstep * kBlock) P.pbuf[1][n + i] = 0.0;
  if (!(P.dbg_flags & 1)) bulk_prime<1, C16>(P.A, ea, pipe);
  const double bb = mega_allreduce(P, acc, e++, scratch, pushed, true, h0, push_base, &ok);
  const double norm_b = sqrt(bb);
  int status = PSB_MAXITER, k_final = 0, n_hist = 0;
  double norm_r = 0.0;
  double alpha = 0.0;                                       // alpha of the previous iteration
  bool flush_x = false;
  if (!ok) {
    status = PSB_MAXITER;
  } else if (bb == 0.0) {                                   // :87-88
    status = PSB_TRIVIAL;
  } else {
    double rr_old = bb;                                     // dot(u, r) with u = r
    double beta = 0.0;
    for (int it = 0;; ++it) {
      // ---------------- phase A: p = r + beta p_old ; Ap = A p ; p.Ap ; x += alpha_prev p_old ------
      stamp(it, 0);
      ea.pold = P.pbuf[(it + 1) & 1];
      ea.pnew = P.pbuf[it & 1];
      if (P.n_wait > 0) { ea.wait_flags = P.my_flags; ea.wait_n = P.n_wait; ea.wait_value = h0 + (unsigned long long)it; }
#pragma unroll
      for (int k = 0; k < kMaxPush; ++k) ea.pp_remote[k] = P.push_p[it & 1][k];
      ea.alpha_prev = alpha;
      acc = 0.0;
      bulk_pass<EPI_DOT_PUP, 1, C16, G>(P.A, P.r, P.Ap, ea, beta, pipe, acc);
      stamp(it, 1);
      // first phase-B rows: r and this thread's own Ap are final -> loads in flight across the barrier.
      // Phase B walks the CTA's tiles LAST to first: the boundary tiles (the last logical ones) come
      // first, so their peer stores of r have long landed when the CTA fences before the ticket.
      double rv[kMegaUnroll], av[kMegaUnroll];
#pragma unroll
      for (int u = 0; u < kMegaUnroll; ++u) {
        const int lt = lt_last - u * gstep;
        const int row = tile_row(lt);
        rv[u] = 0.0; av[u] = 0.0;
        if (!(P.dbg_flags & 2) && row_ok(lt, row)) { rv[u] = mega_ld(P.r + row); av[u] = mega_ld(P.Ap + row); }
      }
      const double pAp = mega_allreduce(P, acc, e++, scratch, false, false, 0ull, push_base, &ok);
      stamp(it, 2);
      if (!ok) { status = PSB_MAXITER; k_final = it; break; }
      if (pAp == 0.0) { status = PSB_BREAKDOWN_PAP; k_final = it; break; }     // :114-115
      alpha = rr_old / pAp;                                                       // :118
      // ---------------- phase B: r -= alpha Ap ; r.r  (x += alpha p is deferred to the next phase A)
      acc = 0.0;
      pushed = false;
      for (int base = lt_last; base >= 0; base -= kMegaUnroll * gstep) {
        if (base != lt_last || (P.dbg_flags & 2)) {
#pragma unroll
          for (int u = 0; u < kMegaUnroll; ++u) {
            const int lt = base - u * gstep;
            const int row = tile_row(lt);
            if (row_ok(lt, row)) { rv[u] = mega_ld(P.r + row); av[u] = mega_ld(P.Ap + row); }
          }
        }
#pragma unroll
        for (int u = 0; u < kMegaUnroll; ++u) {
          const int lt = base - u * gstep;
          const int row = tile_row(lt);
          if (row_ok(lt, row)) {
            const double rn = rv[u] - alpha * av[u];                               // :122
            mega_st(P.r + row, rn);
            if (P.n_push > 0 && (row < P.nopush_lo || row >= P.nopush_hi)) pushed |= mega_push_r(P, row, rn);
            acc += rn * rn;
          }
        }
      }
      stamp(it, 3);
      if (!(P.dbg_flags & 1)) bulk_prime<1, C16>(P.A, ea, pipe);   // first tile of the next phase A: copies in flight
      const double rr = mega_allreduce(P, acc, e++, scratch, pushed, true, h0 + (unsigned long long)it + 1ull,
                                       push_base, &ok);
      stamp(it, 4);
      if (!ok) { status = PSB_MAXITER; k_final = it; break; }
      norm_r = sqrt(rr);                                                         // :125
      if (leader) P.hist[it] = norm_r;                                           // :126
      n_hist = it + 1;
      if ((norm_r <= P.tau * norm_b) || (!P.fail_on_maxiter && it == P.maxiter - 1)) {   // :129-131
        status = PSB_CONVERGED; k_final = it; flush_x = true; break;
      }
      if (it + 1 >= P.maxiter) { status = PSB_MAXITER; k_final = it; flush_x = true; break; }
      beta = rr / rr_old;                                                        // :135
      rr_old = rr;
    }
  }
  bulk_drain(pipe);
  if (flush_x) {                                            // the last x += alpha p  (:121)
    const double* p = P.pbuf[k_final & 1];
    for (int lt = blockIdx.x; lt < n_tiles; lt += gstep) {
      const int row = tile_row(lt);
      if (row_ok(lt, row)) P.x[row] = P.x[row] + alpha * mega_ld(p + row);
    }
  }
  if (leader) {
    P.st->norm_b = norm_b; P.st->norm_r = norm_r;
    P.st->status = status; P.st->k_final = k_final; P.st->n_hist = n_hist; P.st->done = ok ? 1 : 0;
    P.st->epochs_used = e - P.epoch0;
    P.st->halo_epochs_used = (unsigned int)n_hist + 1u;
  }
}

static void mega_caps(int tile_nnz, int* cap_v, int* cap_c, size_t* smem) {
  *cap_v = (tile_nnz + 2 + 1) & ~1;
  *cap_c = (tile_nnz + 6 + 3) & ~3;
  *smem = 2 * ((size_t)*cap_v * 8 + (size_t)*cap_c * 4 + (size_t)(kBlock + 4) * 4);
}

// Rows per tile (multiple of 4, <= 256) such that the CTAs of a `grid_max`-CTA grid get the same
// number of tiles: minimise rounds * (rows + overhead), the per-tile overhead (pipeline hand-over,
// barrier) weighed as 24 rows.  8 192 tiles of 256 rows on 740 CTAs are 12 rounds for 11.07 of
// work (8 % idle); 240-row tiles make it 12 rounds for 11.81.
static int balanced_tile_rows(long long n, long long grid_max) {
  int best = kBlock;
  long long best_cost = -1;
  for (int r = kBlock; r >= kBlock / 2; r -= 4) {
    const long long tiles = (n + r - 1) / r;
    const long long rounds = (tiles + grid_max - 1) / grid_max;
    const long long cost = rounds * (r + 24);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = r; }
  }
  return best;
}

static unsigned long long* g_timeline = nullptr;
static int g_tl_first = 0, g_tl_count = 0;

template <int MINB, bool C16, int G>
static int mega_launch_t(MegaParams& P, psb_csr* A, cudaStream_t stream) {
  static thread_local int per_sm = 0;
  static thread_local size_t cached_smem = 0;
  size_t smem_max;
  int cv, cc;
  mega_caps(A->max_tile_nnz[0], &cv, &cc, &smem_max);       // upper bound: full 256-row tiles
  if (per_sm == 0 || cached_smem != smem_max) {
    if (smem_max > 48 * 1024)
      PSB_CUDA(cudaFuncSetAttribute(pcg_mega_kernel<MINB, C16, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    PSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_mega_kernel<MINB, C16, G>, kBlock, smem_max));
    if (per_sm < 1) { set_error("pcg_mega_launch: kernel does not fit on an SM"); return PSB_ERR_UNSUPP; }
    cached_smem = smem_max;
  }
  const long long grid_max = std::min<long long>((long long)per_sm * sm_count(), (long long)sm_count() * 16);
  // ---- tile plan, cached on the matrix handle ----
  if (A->mega_grid != (int)grid_max || A->mega_tile_rows <= 0) {
    const char* env = getenv("PSB_MEGA_TILE_ROWS");
    // balanced tiles measured no better than full ones (profiles/round2_mega_timeline.md: the SMs,
    // not the CTAs, have to be balanced, and 5 CTAs per SM even out a missing tile): opt-in only
    int rows = env ? (atoi(env) == 0 ? balanced_tile_rows(P.n, grid_max) : atoi(env)) : kBlock;
    if (rows < 4 || rows > kBlock || (rows & 3)) rows = kBlock;
    int tile_nnz = A->max_tile_nnz[0];
    if (rows != kBlock) {
      int rc = csr_max_tile_nnz(A, rows, &tile_nnz, stream);
      if (rc != PSB_OK) return rc;
    }
    A->mega_grid = (int)grid_max; A->mega_tile_rows = rows; A->mega_tile_nnz = tile_nnz;
  }
  P.tile_rows = A->mega_tile_rows;
  size_t smem;
  mega_caps(A->mega_tile_nnz, &P.cap_v, &P.cap_c, &smem);
  if (smem > smem_max) { set_error("pcg_mega_launch: tile plan exceeds the shared-memory bound"); return PSB_ERR_UNSUPP; }
  const long long R = P.tile_rows;
  const long long tiles = (P.n + R - 1) / R;
  P.rot_t0 = 0; P.rot_t1 = 0;
  if (P.int_r1 > P.int_r0) {
    const long long t0 = (P.int_r0 + R - 1) / R;
    const long long t1 = P.int_r1 >= P.n ? tiles : P.int_r1 / R;
    if (t1 > t0) { P.rot_t0 = t0; P.rot_t1 = t1; }
  }
  {   // largest row interval that intersects no push range (rows inside it skip the range tests)
    long long best_lo = 0, best_hi = 0, cur = 0;
    long long lo[kMaxPush], hi[kMaxPush];
    int np = 0;
    for (int k = 0; k < P.n_push; ++k) if (P.push_cnt[k] > 0) { lo[np] = P.push_off[k]; hi[np] = P.push_off[k] + P.push_cnt[k]; ++np; }
    for (int a = 0; a < np; ++a) for (int b = a + 1; b < np; ++b) if (lo[b] < lo[a]) { std::swap(lo[a], lo[b]); std::swap(hi[a], hi[b]); }
    for (int k = 0; k <= np; ++k) {
      const long long gap_hi = k < np ? lo[k] : P.n;
      if (gap_hi - cur > best_hi - best_lo) { best_lo = cur; best_hi = gap_hi; }
      if (k < np) cur = std::max(cur, hi[k]);
    }
    P.nopush_lo = (int)best_lo; P.nopush_hi = (int)best_hi;
  }
  P.timeline = g_timeline; P.tl_first = g_tl_first; P.tl_count = g_tl_count;
  { static int flags = -1; if (flags < 0) { const char* e = getenv("PSB_MEGA_FLAGS"); flags = e ? atoi(e) : 0; } P.dbg_flags = flags; }
  const long long grid = std::max<long long>(1, std::min<long long>(grid_max, tiles));
  void* args[] = {(void*)&P};
  PSB_CUDA(cudaLaunchCooperativeKernel((const void*)pcg_mega_kernel<MINB, C16, G>, dim3((unsigned)grid), dim3(kBlock),
                                       args, smem_max, stream));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return PSB_OK;
}

int pcg_mega_launch(MegaParams& P, psb_csr* A, cudaStream_t stream) {
  // Kernel variant = (resident CTAs per SM it is compiled for, gathers per row loaded together).
  // Measured on B200 (profiles/round2_mega_timeline.md): 4 CTAs per SM with 64 registers and the
  // loads of a WHOLE stencil row (5 or 7 gathers = 10 / 14 loads) in flight at once beat 5 CTAs
  // with 48 registers and 4 gathers by 5 % (C3) - 8 % (3-D).  PSB_MEGA_MINB=5 selects the latter.
  static int minb = 0, g_env = 0;
  if (minb == 0) {
    const char* env = getenv("PSB_MEGA_MINB");
    minb = env ? atoi(env) : 4;
    if (minb != 4 && minb != 5) minb = 4;
    env = getenv("PSB_MEGA_G");
    g_env = env ? atoi(env) : -1;
  }
  P.A = *A;
  const bool c16 = P.A.colind16 != nullptr;
  int g = g_env > 0 ? g_env : std::min(std::max(A->max_row, 4), 8);
  if (minb == 5) g = 4;
#define PSB_MEGA_CASE(MB, GG)                                                        \
  if (minb == MB && g == GG)                                                         \
    return c16 ? mega_launch_t<MB, true, GG>(P, A, stream) : mega_launch_t<MB, false, GG>(P, A, stream);
  PSB_MEGA_CASE(5, 4)
  PSB_MEGA_CASE(4, 4) PSB_MEGA_CASE(4, 5) PSB_MEGA_CASE(4, 6) PSB_MEGA_CASE(4, 7) PSB_MEGA_CASE(4, 8)
#undef PSB_MEGA_CASE
  set_error("pcg_mega_launch: no kernel variant for PSB_MEGA_MINB=%d PSB_MEGA_G=%d", minb, g);
  return PSB_ERR_UNSUPP;
}

}  // namespace psb

// Profiling hook: the persistent PCG kernels launched afterwards record %globaltimer stamps of
// iterations [first_iter, first_iter + n_iters) into d_buf, laid out [iteration][CTA][6] u64:
// 0 phase A starts, 1 phase A done, 2 p.Ap reduced, 3 phase B done, 4 r.r reduced.  NULL disables.
extern "C" int psb_debug_mega_timeline(void* d_buf, int32_t first_iter, int32_t n_iters) {
  psb::g_timeline = (unsigned long long*)d_buf;
  psb::g_tl_first = first_iter;
  psb::g_tl_count = d_buf ? n_iters : 0;
  return PSB_OK;
}
