// Sparse triangular solve with the wavefront kept in SHARED memory: one CTA, or one thread-block
// cluster of 4 CTAs with the window replicated through distributed shared memory.
//
// Why: the factors the reference produces (SuperLU IC / ILUT / LU, ICPreconditioner.py:45-63,
// ILUTPreconditioner.py:51-78, VCycleManager.py:34-37) have thousands of dependency levels
// with only tens of rows each.  The grid-wide kernel in sptrsv.cu
// hands a finished x_j to its consumers through L2: measured 1.5 - 2 us per level on B200
// whatever the tuning -- the solve is bound by that latency, not by bytes.  A microbenchmark
// (tools/micro/smem_pingpong.cu) puts the hand-over between two warps of one CTA through shared
// memory at 145 - 200 cycles (0.08 - 0.1 us), so here the whole wavefront lives in one SM:
//
//  * rows are processed in the level-major order of the analysis; "position" q = rank of a row in
//    that order.  Measured on the reference's factors: > 99.9 % of the dependencies of an IC
//    factor, and every dependency on the previous level, lie within a few thousand positions.
//  * the last `wslots` results live in a circular window in shared memory, indexed by position
//    (x_q at slot q mod wslots), pre-filled with the NaN sentinel: the value is its own ready
//    flag, exactly as in the grid-wide kernel.  Dependencies further back than the window
//    ("far", encoded at analysis time) are read from the global x with relaxed loads -- they
//    were produced long ago and are never on the critical path.
//  * chunks (up to 32 short rows of ONE level, or one long row for the whole warp) are dealt
//    round-robin to the 16 warps.  A warp publishes the chunk it has started in `prog[]`; nobody
//    starts chunk g before every warp has started a chunk >= g - kAhead, and the warp starting
//    chunk g resets the slots chunk g + kAhead will use.  With near-ness limited to
//    wslots - 32 (2 kAhead + 2) positions a polled slot holds either the sentinel or the value of
//    exactly the awaited position (argument in DESIGN.md section 4).
//  * the (col, val) pairs of a chunk are copied into a per-warp staging buffer by the TMA engine
//    (cp.async.bulk + mbarrier) as soon as the previous chunk of the warp is done -- two rounds
//    after an L2 prefetch of the same bytes -- so that the critical window touches shared memory
//    only: four polls in flight, entries consumed in the order in which they become available
//    (oldest dependency level first, rows right-aligned in their chunk), and after a row's last
//    dependency arrives only multiply, subtract, scale and the shared-memory store remain.
//  * a warp whose chunk is far behind the wavefront (oldest chunk in work = min prog[]) sleeps
//    in proportion to the distance; only the next chunks spin, each on its own warp scheduler.
//  * arithmetic identical to the grid-wide kernel: b_i - sum L_ij x_j accumulated in
//    dependency-level order, product rounded first, times the reciprocal of the diagonal last
//    (invdiag is formed once, as scipy's spsolve_triangular does) -> bit-identical results, and
//    no fp64 division on the critical path.
//  * cluster variant (kC = 4): every CTA keeps a replica of the window and of prog[]; results,
//    resets and progress are stored into all replicas with st.shared::cluster; chunk g runs on
//    CTA g mod 4 (profiles/round1g_cluster.md).
#include "sptrsv.cuh"

#include <algorithm>
#include <cstdlib>

namespace psb {

namespace {

constexpr unsigned long long kSentinelBits = 0xFFF8DEADBEEF0B20ull;
constexpr int kCtaWarps = kTrsvCtaWarps;
constexpr int kCtaThreads = kCtaWarps * 32;
constexpr int kAhead = 2 * kCtaWarps;
constexpr int kSpinLimitCta = 1 << 24;
static_assert(kAhead == kTrsvAhead, "the analysis sizes the near limit with kTrsvAhead");

// Readiness test on the high word only (one 32-bit compare on the critical path): a value whose
// high word equals the sentinel's is a NaN with that very payload, which no arithmetic produces.
__device__ __forceinline__ bool there(double v) {
  return (unsigned int)__double2hiint(v) != (unsigned int)(kSentinelBits >> 32);
}
__device__ __forceinline__ double ld_x_relaxed(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_x_relaxed(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}
// shared memory through 32-bit addresses (a generic pointer costs an address conversion, i.e. a
// special-register read, on every trip of a spin loop)
__device__ __forceinline__ double lds_vol(uint32_t a) {
  double v;
  asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_vol(uint32_t a, double v) {
  asm volatile("st.volatile.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ int lds_vol_i32(uint32_t a) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_vol_i32(uint32_t a, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
// staged (col, val) pairs: written by the TMA engine, read after the mbarrier wait
__device__ __forceinline__ int lds_i32(uint32_t a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void mbar_init32(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect32(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try32(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s32(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct CtaView {
  int64_t n;
  int n_groups;
  int wslots;                 // power of two
  int wraps;                  // n > wslots: slots are reused
  int stage_len;              // entries per lane the staging buffer of a warp holds
  const int32_t* order;
  const double* diag;         // 1 / diagonal (1 for unit_diag)
  const int4* meta;           // per chunk: {base lo, base hi, item0, rows | has_far << 6 | len << 7}
  const int32_t* cols;        // >= 0: byte offset in the window (padding: the zero slot behind it); <= -2: far, row = -c - 2
  const double* vals;
  int* error;
  int near_chunks;            // chunks this close to the wavefront spin instead of sleeping
  int ns_per_chunk_q4;        // sleep per chunk of distance beyond that, in 1/16 ns
  long long* trace;           // debugging: 12 values per chunk (null: off)
};

// One dependency value.  The analysis stores, for every entry, where to look: a byte offset into the
// shared-memory window (near dependency; padding points at a slot that always holds 0.0, so that
// it subtracts 0 * 0) or, negative, the row of a far dependency to be read from the global vector.
template <bool kFar>
__device__ __forceinline__ double poll(int c, uint32_t wbase, const double* x) {
  if (kFar && c < 0) return ld_x_relaxed(x + (-c - 2));
  return lds_vol(wbase + (uint32_t)c);
}

// Per-lane tight spin on one dependency: load, one compare, branch.  No warp votes: the rows of
// a chunk are independent, so a lane never waits for a lane of its own warp and lanes that have
// their value simply wait at the reconvergence point.
#define PSB_TRSV_AWAIT(xv, c)                                                   \
  if (!there(xv)) {                                                             \
    int budget = kSpinLimitCta;                                                 \
    if (!kFar || c >= 0) {                                                      \
      const uint32_t a_ = wbase + (uint32_t)c;                                  \
      _Pragma("unroll 1")                                                       \
      do {                                                                      \
        xv = lds_vol(a_);                                                       \
        if (--budget == 0) { *error = 1; break; }                               \
      } while (!there(xv));                                                     \
    } else {                                                                    \
      const double* a_ = x + (-c - 2);                                          \
      _Pragma("unroll 1")                                                       \
      do {                                                                      \
        xv = ld_x_relaxed(a_);                                                  \
        if (--budget == 0) { *error = 1; break; }                               \
      } while (!there(xv));                                                     \
    }                                                                           \
  }

// (col, val) of entry e of the staged round into slot S; past the end the last entry is read again
// (never consumed: the walk stops at the end)
#define PSB_TRSV_LOAD(S, e)                                                     \
  {                                                                             \
    const uint32_t e_ = (uint32_t)min((int)(e), nr - 1);                        \
    c##S = lds_i32(sc + e_ * 128u);                                             \
    v##S = lds_f64(sv + e_ * 256u);                                             \
  }

// consume entry e from slot S; start the poll of entry e + 4 (slot P, whose (col, val) arrived four
// steps ago); fetch (col, val) of entry e + 8 into slot S
#define PSB_TRSV_STEP(S, P, e)                                                  \
  {                                                                             \
    PSB_TRSV_AWAIT(x##S, c##S)                                                  \
    if (kTrace) { ta[2] = ta[1]; ta[1] = ta[0]; ta[0] = clock64(); }            \
    acc = acc - v##S * x##S;                                                    \
    if ((e) + 1 >= nr) break;                                                   \
    x##P = poll<kFar>(c##P, wbase, x);                                          \
    PSB_TRSV_LOAD(S, (e) + 8)                                                   \
  }

// Walk the nr staged entries of this lane in the dependency order of the analysis (oldest level
// first: a row's newest dependency is its last operand) with a rotating software pipeline: while
// entry e is consumed the polls of e+1..e+3 are in flight and the (col, val) pairs of e+4..e+7 are
// in registers, so the common path never waits for a shared-memory load, has one branch per entry,
// and once the last dependency has arrived only multiply, subtract, scale and store remain.
template <bool kFar, bool kTrace>
__device__ __forceinline__ double walk(double acc, int nr, uint32_t sc, uint32_t sv, uint32_t wbase,
                                       const double* x, int* error, long long* ta) {
  int c0, c1, c2, c3, c4, c5, c6, c7;
  double v0, v1, v2, v3, v4, v5, v6, v7;
  PSB_TRSV_LOAD(0, 0) PSB_TRSV_LOAD(1, 1) PSB_TRSV_LOAD(2, 2) PSB_TRSV_LOAD(3, 3)
  PSB_TRSV_LOAD(4, 4) PSB_TRSV_LOAD(5, 5) PSB_TRSV_LOAD(6, 6) PSB_TRSV_LOAD(7, 7)
  double x0 = poll<kFar>(c0, wbase, x), x1 = poll<kFar>(c1, wbase, x);
  double x2 = poll<kFar>(c2, wbase, x), x3 = poll<kFar>(c3, wbase, x);
  double x4 = 0.0, x5 = 0.0, x6 = 0.0, x7 = 0.0;
#pragma unroll 1
  for (int k = 0;; k += 8) {
    PSB_TRSV_STEP(0, 4, k)
    PSB_TRSV_STEP(1, 5, k + 1)
    PSB_TRSV_STEP(2, 6, k + 2)
    PSB_TRSV_STEP(3, 7, k + 3)
    PSB_TRSV_STEP(4, 0, k + 4)
    PSB_TRSV_STEP(5, 1, k + 5)
    PSB_TRSV_STEP(6, 2, k + 6)
    PSB_TRSV_STEP(7, 3, k + 7)
  }
  return acc;
}

// ---- distributed shared memory (thread-block cluster): every CTA of the cluster keeps a replica of
// the window and of the progress array; a producer stores its result into all of them.
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f64(uint32_t a, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_s32(uint32_t a, int v) {
  asm volatile("st.shared::cluster.s32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// kC = 1: one CTA.  kC > 1: a cluster of kC CTAs on kC SMs; chunk g runs on CTA g mod kC, warp
// (g / kC) mod 16, so consecutive chunks -- the critical one and the ones spinning behind it -- sit
// on different SMs and 16 kC warps cover wide levels.  The price is the DSMEM hand-over (~215
// cycles instead of 38), so the cluster is used when a level has more chunks than one CTA can
// keep on separate schedulers (measured: 4.5 - 8.5 chunks per level 0.78 - 0.88 us per level
// against 1.0 - 1.7 for one CTA and 1.15 for the grid; profiles/round1g_cluster.md).
template <int kC, bool kTrace>
__global__ void __launch_bounds__(kCtaThreads, 1)
trsv_cta_kernel(const CtaView T, const double* __restrict__ rhs, double* x,
                const int32_t* __restrict__ rhs_map, double* out2,
                const int32_t* __restrict__ out_map, const int* d_skip) {
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;       // uniform over the whole cluster
  constexpr int kStride = kCtaWarps * kC;                   // chunks per round of all warps
  constexpr int kAheadC = 2 * kStride;                      // run-ahead allowance in chunks: two rounds (with one
                                                            // round every round ends in a cluster-wide wait)
  constexpr int kProg = kStride < 32 ? 32 : kStride;        // progress entries (one per warp of the cluster)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ int prog_raw[kProg];
  __shared__ __align__(8) unsigned long long bars_raw[kCtaWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cta = kC == 1 ? 0 : (int)cluster_rank();
  const int gw = warp * kC + cta;                           // this warp's slot in prog[]: first chunk it owns
  // the shared-window addresses are made opaque to the compiler: otherwise it re-derives them from
  // special registers (S2R, tens of cycles) right before each use, also on the critical path
  uint32_t wbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
  uint32_t prog_base = (uint32_t)__cvta_generic_to_shared(prog_raw);
  asm volatile("mov.u32 %0, %0;" : "+r"(wbase));
  asm volatile("mov.u32 %0, %0;" : "+r"(prog_base));
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(bars_raw) + 8u * warp;
  // replicas: cluster-space addresses of every CTA's window, most urgent consumer first (the next
  // chunk runs on the next CTA); lane c < kC also holds the address of CTA c's progress array
  uint32_t rwin[kC];
  uint32_t rprog = 0;
  if (kC > 1) {
#pragma unroll
    for (int c = 0; c < kC; ++c) rwin[c] = mapa_u32(wbase, (uint32_t)((cta + 1 + c) % kC));
    rprog = mapa_u32(prog_base, (uint32_t)(lane % kC));
  }
  // layout: window | 128 bytes whose first 8 are the always-zero slot of the padding entries |
  // per warp: stage_len x 32 columns (int32) then stage_len x 32 values (fp64)
  const uint32_t stage_bytes = (uint32_t)T.stage_len * 384u;
  const uint32_t sc_base = wbase + (uint32_t)T.wslots * 8u + 128u + (uint32_t)warp * stage_bytes;
  const uint32_t sv_base = sc_base + (uint32_t)T.stage_len * 128u;
  const uint32_t sc = sc_base + 4u * lane, sv = sv_base + 8u * lane;
  const int wmask = T.wslots - 1;
  const double kNotReady = __longlong_as_double((long long)kSentinelBits);
  const int fill = (int)min((int64_t)T.wslots, T.n);
  for (int i = threadIdx.x; i < fill; i += kCtaThreads) sts_vol(wbase + ((uint32_t)i << 3), kNotReady);
  if (threadIdx.x == 0) sts_vol(wbase + (uint32_t)T.wslots * 8u, 0.0);
  for (int i = threadIdx.x; i < kProg; i += kCtaThreads)
    prog_raw[i] = (i < kStride && i < T.n_groups) ? i - kStride : INT32_MAX;
  if (lane == 0) mbar_init32(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (kC > 1) cluster_sync_all();        // nobody stores into a replica before its owner has initialised it

  // oldest chunk in work anywhere in the cluster (local replica: may lag, never runs ahead)
  auto low_water = [&]() -> int {
    int v = INT32_MAX;
#pragma unroll
    for (int j = 0; j < kProg / 32; ++j) v = min(v, lds_vol_i32(prog_base + 4u * (uint32_t)(lane + 32 * j)));
    return __reduce_min_sync(0xffffffffu, v);
  };
  auto publish = [&](int value) {
    if (kC == 1) { if (lane == 0) sts_vol_i32(prog_base + 4u * (uint32_t)gw, value); }
    else if (lane < kC) st_cluster_s32(rprog + 4u * (uint32_t)gw, value);
  };

  const int4 kNone = make_int4(0, 0, 0, 0);
  int g = gw;
  bool ok = true;
  int4 m0 = g < T.n_groups ? __ldg(T.meta + g) : kNone;
  int4 m1 = g + kStride < T.n_groups ? __ldg(T.meta + g + kStride) : kNone;
  int4 m2 = g + 2 * kStride < T.n_groups ? __ldg(T.meta + g + 2 * kStride) : kNone;
  uint32_t parity = 0;
  // stage the first round of the first chunk
  if (g < T.n_groups && lane == 0) {
    const int64_t b0 = ((int64_t)m0.y << 32) | (unsigned int)m0.x;
    const int n0 = min(m0.w >> 7, T.stage_len);
    if (n0 > 0) {
      mbar_expect32(bar, (uint32_t)n0 * 384u);
      bulk_g2s32(sc_base, T.cols + b0, (uint32_t)n0 * 128u, bar);
      bulk_g2s32(sv_base, T.vals + b0, (uint32_t)n0 * 256u, bar);
    }
  }
  for (; ok && g < T.n_groups; g += kStride) {
    // meta of the chunk three rounds ahead (arrives while this chunk is worked on)
    const int4 m3 = g + 3 * kStride < T.n_groups ? __ldg(T.meta + g + 3 * kStride) : kNone;
    const int4 mr = m2;                         // chunk g + kAheadC: its slots are reset below
    const int64_t base = ((int64_t)m0.y << 32) | (unsigned int)m0.x;
    const int item0 = m0.z;
    const int rows = m0.w & 63;                 // 0: one long row for the warp
    const bool has_far = (m0.w & 64) != 0;      // some entry of the chunk reads the global vector
    const int len = m0.w >> 7;                  // entries per lane
    const bool is_long = rows == 0;
    const bool owner = is_long ? (lane == 0) : (lane < rows);
    long long t_start = 0, t_ready = 0, t_woke = 0, t_staged = 0;
    long long ta[3] = {0, 0, 0};
    if (kTrace) t_start = clock64();

    int row = 0, q = 0;
    double d = 1.0;
    if (owner) {
      q = item0 + (is_long ? 0 : lane);
      row = T.order[q];
      d = T.diag[q];
    }

    if (T.wraps) {
      // nobody runs more than kAheadC chunks ahead of the slowest warp (nothing to protect during
      // the first round: the window is >= 64 kStride slots, no slot is reused yet) ...
      int spins = 0;
      while (g >= kStride) {
        if (low_water() >= g - kAheadC) break;
        if (++spins > kSpinLimitCta) { *T.error = 2; ok = false; break; }
        __nanosleep(64);
      }
      // ... so the readers of what the slots of chunk g + kAheadC held one lap ago are done: reset them
      if (g + kAheadC < T.n_groups) {
        const int r_item = mr.z, r_rows = (mr.w & 63) == 0 ? 1 : (mr.w & 63);
        if (lane < r_rows) {
          const uint32_t off = (uint32_t)((r_item + lane) & wmask) << 3;
          if (kC == 1) sts_vol(wbase + off, kNotReady);
          else {
#pragma unroll
            for (int c = 0; c < kC; ++c) st_cluster_f64(rwin[c] + off, kNotReady);
          }
        }
      }
      if (kC > 1) asm volatile("fence.acq_rel.cluster;" ::: "memory");   // resets before the progress store, everywhere
      __syncwarp();                      // the resets of all lanes before the progress store
    }
    publish(g);                          // also the wavefront estimate of the sleepers

    double acc = 0.0;
    if (owner) acc = rhs_map ? rhs[rhs_map[row]] : rhs[row];
    const uint32_t wout = (uint32_t)(q & wmask) << 3;
    double* xout = x + row;

    // prefetch the chunk of two rounds ahead into L2: cols (128 B per entry row), vals (256 B)
    if (g + 2 * kStride < T.n_groups) {
      const int64_t pbase = ((int64_t)m2.y << 32) | (unsigned int)m2.x;
      const int plen = m2.w >> 7;
      for (int i = lane; i < plen; i += 32) prefetch_l2(T.cols + pbase + (int64_t)i * 32);
      for (int i = lane; i < 2 * plen; i += 32) prefetch_l2(T.vals + pbase + (int64_t)i * 16);
      if (lane == 0) prefetch_l2(T.order + m2.z);
      if (lane == 1) prefetch_l2(T.diag + m2.z);
      if (lane == 2) prefetch_l2(T.diag + m2.z + 16);
    }

    // A spinning warp competes for issue slots with the warps on the critical path: while this
    // chunk is far behind the wavefront (oldest chunk in work = min prog[]) sleep in proportion
    // to the distance; only the next few levels spin on their dependencies.
    if (len > 0) {
      for (;;) {
        const int ahead = g - low_water() - T.near_chunks;
        if (ahead <= 0) break;
        __nanosleep((unsigned)((min(ahead, 512) * T.ns_per_chunk_q4) >> 4));
      }
    }
    if (kTrace) t_woke = clock64();

    for (int k0 = 0; k0 < len; k0 += T.stage_len) {
      const int nr = min(len - k0, T.stage_len);           // entries per lane staged in this round
      if (k0 > 0) {
        // a row longer than the staging buffer: next round (the lanes are done with the last one)
        __syncwarp();
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_expect32(bar, (uint32_t)nr * 384u);
          bulk_g2s32(sc_base, T.cols + base + (int64_t)k0 * 32, (uint32_t)nr * 128u, bar);
          bulk_g2s32(sv_base, T.vals + base + (int64_t)k0 * 32, (uint32_t)nr * 256u, bar);
        }
      }
      {
        int spins = 0;
        while (!mbar_try32(bar, parity)) {
          if (++spins > kSpinLimitCta) { *T.error = 3; ok = false; break; }
        }
        parity ^= 1u;
      }
      if (kTrace && k0 == 0) t_staged = clock64();
      if (has_far) acc = walk<true, kTrace>(acc, nr, sc, sv, wbase, x, T.error, ta);
      else         acc = walk<false, kTrace>(acc, nr, sc, sv, wbase, x, T.error, ta);
    }
    if (is_long) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    const double r = acc * d;            // d = 1 / diagonal (scipy: x = y * invdiag); 1 for unit_diag
    if (owner) {
      if (kC == 1) sts_vol(wbase + wout, r);
      else {
#pragma unroll
        for (int c = 0; c < kC; ++c) st_cluster_f64(rwin[c] + wout, r);   // next CTA first, own replica last
      }
    }
    if (kTrace) t_ready = clock64();
    // everything below is off the critical path
    if (owner) {
      st_x_relaxed(xout, r);
      if (out2 != nullptr) out2[out_map[row]] = r;
    }
    __syncwarp();
    if (lane == 0 && g + kStride < T.n_groups) {            // stage the first round of the next chunk
      const int64_t b1 = ((int64_t)m1.y << 32) | (unsigned int)m1.x;
      const int n1 = min(m1.w >> 7, T.stage_len);
      if (n1 > 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect32(bar, (uint32_t)n1 * 384u);
        bulk_g2s32(sc_base, T.cols + b1, (uint32_t)n1 * 128u, bar);
        bulk_g2s32(sv_base, T.vals + b1, (uint32_t)n1 * 256u, bar);
      }
    }
    if (kTrace && lane == 0) {
      long long* tr = T.trace + 12 * (int64_t)g;
      tr[0] = t_start; tr[1] = 0; tr[2] = t_woke; tr[3] = t_staged;
      tr[4] = t_ready; tr[5] = clock64(); tr[6] = 0; tr[7] = len;
      tr[8] = ta[0]; tr[9] = ta[1]; tr[10] = ta[2]; tr[11] = 0;
    }
    m0 = m1; m1 = m2; m2 = m3;
  }
  __syncwarp();
  publish(INT32_MAX);
  if (kC > 1) cluster_sync_all();        // no CTA leaves while another may still store into its replica
}

}  // namespace

template <int kC>
static int launch_cta(const CtaView& V, size_t smem, bool trace, const double* rhs, double* x,
                      const int32_t* rhs_map, double* out2, const int32_t* out_map, const int* d_skip,
                      cudaStream_t st) {
  static thread_local bool configured = false;
  if (!configured) {
    PSB_CUDA(cudaFuncSetAttribute(trsv_cta_kernel<kC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsvSmemBudget));
    PSB_CUDA(cudaFuncSetAttribute(trsv_cta_kernel<kC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsvSmemBudget));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kC);
  cfg.blockDim = dim3(kCtaThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = kC > 1 ? 1 : 0;
  if (trace) PSB_CUDA(cudaLaunchKernelEx(&cfg, trsv_cta_kernel<kC, true>, V, rhs, x, rhs_map, out2, out_map, d_skip));
  else       PSB_CUDA(cudaLaunchKernelEx(&cfg, trsv_cta_kernel<kC, false>, V, rhs, x, rhs_map, out2, out_map, d_skip));
  PSB_LAUNCH_CHECK();
  return PSB_OK;
}

int trsv_solve_cta(const psb_trsv* T, int cluster, const double* rhs, double* x, const int32_t* rhs_map,
                   double* out2, const int32_t* out_map, const int* d_skip, cudaStream_t st) {
  const size_t smem = (size_t)T->wslots * sizeof(double) + 128 + (size_t)kCtaWarps * T->stage_len * 384;
  if (cluster && !T->cluster_ok) {
    set_error("trsv_solve: this factor was not analysed for the cluster kernel");
    return PSB_ERR_UNSUPP;
  }
  const int kc = cluster ? kTrsvClusterSize : 1;
  // chunks per level decide who spins (the next 2 levels) and how long the others sleep (half of
  // an optimistic 128 ns per level of distance)
  const double cpl = std::max(1.0, (double)T->n_groups / std::max(T->n_levels, 1));
  // Measured on B200 (profiles/round1d_trsv.md): with one chunk per level the next 3 chunks spin,
  // each on a warp scheduler of its own (warp = chunk mod 16, scheduler = warp mod 4), and the
  // critical warp has the fourth to itself; a fourth spinner shares its scheduler and costs a
  // factor 2.  A pause inside the spin loop (nanosleep 20 - 50 ns) oversleeps and is worse.
  // Swept again in round 2 on IC factors of 256^2 / 512^2 / 1024^2 (1.0 / 1.8 / 3.4 chunks per level,
  // tools/trsv_probe.py with PSB_PROBE_SWEEP=1): the one-CTA kernel wants 2.5 levels of spinners once a
  // level has more than one chunk, and a sleep of 64 ns per CHUNK of distance, not per level
  // (IC 1024^2: 0.76 / 0.90 -> 0.66 / 0.68 us per level; IC 512^2: 0.56 -> 0.52 / 0.55; IC 256^2
  // unchanged at 0.38).  The cluster of 4 (64 warps, 16 schedulers) wants FOUR levels of spinners, up to
  // about half its warps, and 128 ns per level: Gauss-Seidel triangles of 256^2 / 384^2 / 512^2 grids
  // 0.79 / 0.82 / 0.88 -> 0.60 / 0.63 / 0.66 us per level (PSB_PROBE_KERNEL=cluster).
  // PSB_TRSV_NEAR_LEVELS / PSB_TRSV_SLEEP_NS (ns per level of distance): A/B knobs.
  double near_levels = kc == 1 ? (cpl >= 1.5 ? 2.5 : 2.0) : 4.0;
  double sleep_ns = kc == 1 ? 64.0 * cpl : 128.0;
  if (const char* e = getenv("PSB_TRSV_NEAR_LEVELS")) near_levels = atof(e);
  if (const char* e = getenv("PSB_TRSV_SLEEP_NS")) sleep_ns = atof(e);
  int near_chunks = (int)(near_levels * cpl + 1.5);
  if (kc > 1) near_chunks = std::min(near_chunks, 36);
  const int ns_q4 = std::max(1, (int)(16.0 * sleep_ns / cpl));
  CtaView V{T->n, T->n_groups, T->wslots, T->n > T->wslots ? 1 : 0, T->stage_len, T->d_order, T->d_diag,
            reinterpret_cast<const int4*>(T->d_wmeta), T->d_wcols, T->d_vals, T->d_error,
            near_chunks, ns_q4, T->d_trace};
  if (kc == 1) return launch_cta<1>(V, smem, T->d_trace != nullptr, rhs, x, rhs_map, out2, out_map, d_skip, st);
  return launch_cta<kTrsvClusterSize>(V, smem, T->d_trace != nullptr, rhs, x, rhs_map, out2, out_map, d_skip, st);
}

}  // namespace psb
