// Sparse triangular solves for the IC / ILUT preconditioners (and the AMG Gauss-Seidel
// smoother and coarse solve): x = T^-1 b for a lower or upper triangular CSR factor.
//
// Replaces scipy.sparse.linalg.spsolve_triangular (PySolvers/Linear/ICPreconditioner.py:61-63)
// and SuperLU.solve (PySolvers/Linear/ILUTPreconditioner.py:67,78), both SuperLU gstrs on
// one CPU thread.  The factors come from the reference's own SuperLU setup on the host and
// are uploaded once.
//
// The reference's IC factor of the 2-D Laplacian has ~11.8 m dependency levels with only
// ~m/12 rows each: a kernel launch (or a grid barrier) per
// level is hopeless, the solve is bound by the LATENCY of the dependency chain, not by HBM.
// Design:
//  * analysis (host, once): level(i) = 1 + max level of the rows it depends on; rows sorted
//    level-major (ascending inside a level); the strictly triangular part is repacked in
//    that order as SELL-32 (entry k of 32 consecutive items is contiguous) so every load of
//    the solve is coalesced; the diagonal is split out.
//  * solve: ONE persistent launch.  Warps claim work chunks with an atomic counter.  A chunk
//    is either 32 consecutive SHORT rows (<= 32 off-diagonal entries): each THREAD owns one
//    row and accumulates b_i - sum L_ij x_j sequentially in the order in which the dependencies
//    become available (ascending level, ties in stored order) and multiplies by the reciprocal
//    of the diagonal last (scipy's spsolve_triangular forms invdiag = 1/diag once and ends with
//    x = y * invdiag), so the
//    dependency that arrives last -- the one nearest the diagonal -- is also the last operand
//    and everything else is folded in beforehand; or ONE LONG row (exact LU factors of the
//    AMG coarse level have rows of hundreds of entries) spread over the 32 lanes of the warp
//    and combined with a shuffle tree.  Every lane keeps a window of 4 entries in registers:
//    the 4 dependencies are polled with independent loads and consumed in order as far as
//    they are ready, the next entries are fetched behind them.
//    Readiness travels with the data: x is pre-filled with a NaN sentinel and a consumer
//    polls x[j] (ld.volatile, L2) until it changes -- one L2 round trip per level on the
//    critical path, no flags, no fences, no grid barrier.  The poll loop never blocks a
//    lane on another (each trip is non-blocking), so dependencies inside a warp resolve.
//    Items are claimed in dependency order, hence every awaited row is owned by a warp that
//    is already running: no deadlock whatever the residency.  A bounded spin count turns a
//    would-be hang into an error code.
#include "sptrsv.cuh"
#include "prec.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>

namespace psb {

static constexpr unsigned long long kSentinelBits = 0xFFF8DEADBEEF0B20ull;   // a quiet NaN payload
static constexpr int kSpinLimit = 1 << 22;   // polls per lane before giving up (>= 1 s)
static constexpr int kIdleTrips = 10;        // a warp this many polls (~3 us) without progress is >= 2 levels
                                              // behind the wavefront: it may sleep between polls

__device__ __forceinline__ double ld_volatile(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile(double* p, double v) {
  asm volatile("st.volatile.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ bool is_ready(double v) {
  return (unsigned long long)__double_as_longlong(v) != kSentinelBits;
}

__global__ void __launch_bounds__(kBlock)
trsv_prepare_kernel(double* __restrict__ x, int64_t n, unsigned int* counter, const int* d_skip) {
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;
  const double s = __longlong_as_double((long long)kSentinelBits);
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    x[i] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0u;
}

struct TrsvView {
  int64_t n;
  int n_groups;
  const int32_t* order;
  const double* diag;
  const int64_t* grp_ptr;
  const int32_t* grp_item;
  const int32_t* grp_rows;
  const int32_t* cols;
  const double* vals;
  const int32_t* row_cnt;    // off-diagonal entries of item q
  unsigned int* counter;
  int* error;
  int unit_diag;
  long long* trace;          // debugging (psb_trsv_set_trace): per chunk {claimed, done, first item}, null: off
  int idle_trips;            // polls without progress before a warp starts to sleep between polls
  unsigned int sleep_cap;    // longest sleep between two polls (ns)
};

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ double ld_relaxed(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}

// kTrace: %globaltimer stamps per chunk (tools/trsv_levels.py)
template <bool kTrace>
__global__ void __launch_bounds__(kBlock)
trsv_solve_kernel(const TrsvView T, const double* __restrict__ rhs, double* x,
                  const int32_t* __restrict__ rhs_map, double* out2,
                  const int32_t* __restrict__ out_map, const int* d_skip) {
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;
  const int lane = threadIdx.x & 31;
  const double kNotReady = __longlong_as_double((long long)kSentinelBits);
  for (;;) {
    unsigned int g = 0;
    if (lane == 0) g = atomicAdd(T.counter, 1u);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= (unsigned int)T.n_groups) return;
    const int64_t base = T.grp_ptr[g];
    int len = (int)((T.grp_ptr[g + 1] - base) >> 5);           // entries per lane
    const int rows = T.grp_rows[g];                            // 0: one long row for the warp; < 0: -rows rows of 8 lanes
    const int item0 = T.grp_item[g];
    if (kTrace && lane == 0) { T.trace[3 * (int64_t)g] = global_ns(); T.trace[3 * (int64_t)g + 2] = item0; }
    const bool is_sub = rows < 0;
    const bool is_long = rows <= 0;                            // lanes share rows: reduce before the store
    const bool owner = rows == 0 ? (lane == 0) : (is_sub ? ((lane & 7) == 0 && (lane >> 3) < -rows) : (lane < rows));
    int row = 0;
    double acc = 0.0, d = 1.0;
    if (owner) {
      const int q = item0 + (rows == 0 ? 0 : (is_sub ? (lane >> 3) : lane));
      row = T.order[q];
      acc = rhs_map ? rhs[rhs_map[row]] : rhs[row];
      d = T.diag[q];
    }
    // short rows are RIGHT-aligned in their chunk (every row's newest dependency in the last
    // entry row, for the lock-step walk of the one-CTA kernel): this lane's list starts further in
    const int width = len;
    const int first = (is_long || !owner) ? 0 : width - T.row_cnt[item0 + lane];
    if (is_sub && (lane >> 3) >= -rows) len = 0;              // lanes of an absent row
    const int32_t* cp = T.cols + base + lane + (int64_t)first * 32;
    const double*  vp = T.vals + base + lane + (int64_t)first * 32;
    // window of 4 entries: (c0,v0) is the next one to consume
    len -= first;
    if (!is_long && !owner) len = 0;
    int k = 0;
    int c0 = -1, c1 = -1, c2 = -1, c3 = -1;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    if (len > 0) { c0 = cp[0];  v0 = vp[0]; }
    if (len > 1) { c1 = cp[32]; v1 = vp[32]; }
    if (len > 2) { c2 = cp[64]; v2 = vp[64]; }
    if (len > 3) { c3 = cp[96]; v3 = vp[96]; }
    int spins = 0, idle = 0;
    bool fed = !(c0 >= 0);            // all entries of this lane consumed (padding ends a lane's list)
    bool done = false;
    while (!__all_sync(0xffffffffu, done)) {
      bool made_progress = false;
      if (!fed) {
        // waiting mode: one poll of the next dependency, nothing else on the path
        double x0 = ld_relaxed(x + c0);
        if (is_ready(x0)) {
          made_progress = true;
          // streaming mode: the following three entries are polled together and consumed in
          // order as far as they are ready; new entries are fetched behind them
          double x1 = c1 >= 0 ? ld_relaxed(x + c1) : kNotReady;
          double x2 = c2 >= 0 ? ld_relaxed(x + c2) : kNotReady;
          double x3 = c3 >= 0 ? ld_relaxed(x + c3) : kNotReady;
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            if (c0 >= 0 && is_ready(x0)) {
              acc = acc - v0 * x0;                    // stored order, product rounded first
              ++k;
              c0 = c1; v0 = v1; x0 = x1;
              c1 = c2; v1 = v2; x1 = x2;
              c2 = c3; v2 = v3; x2 = x3;
              const int nk = k + 3;
              if (nk < len) { c3 = cp[(int64_t)nk * 32]; v3 = vp[(int64_t)nk * 32]; } else { c3 = -1; v3 = 0.0; }
              x3 = kNotReady;
            }
          }
          if (c0 < 0) fed = true;
        } else if (++spins > kSpinLimit) {
          // a dependency never became ready: raise the flag AND poison the row with an ordinary
          // quiet NaN (not the sentinel, so consumers do not wait for it in turn) -- the
          // failure is then visible in the result itself
          *T.error = 1; fed = true;
          acc = __longlong_as_double(0x7ff8000000000000ll);
        }
      }
      // warps whose dependencies are still levels away back off instead of hammering L2
      if (__any_sync(0xffffffffu, made_progress)) idle = 0;
      else if (++idle > T.idle_trips) __nanosleep(min(32u * (unsigned)(idle - T.idle_trips), T.sleep_cap));
      if (!done) {
        if (is_long) {
          // the row is complete when every lane has consumed its share
          if (__all_sync(0xffffffffu, fed)) {
            double t = acc;
            if (!is_sub) { t += __shfl_xor_sync(0xffffffffu, t, 16); t += __shfl_xor_sync(0xffffffffu, t, 8); }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (owner) {
              const double r = T.unit_diag ? t : t * d;       // d = 1 / diagonal
              st_relaxed(x + row, r);
              if (out2 != nullptr) out2[out_map[row]] = r;
            }
            done = true;
          }
        } else if (fed) {
          if (owner) {
            const double r = T.unit_diag ? acc : acc * d;   // times the reciprocal of the diagonal, last
            st_relaxed(x + row, r);
            if (out2 != nullptr) out2[out_map[row]] = r;
          }
          done = true;
        }
      }
    }
    if (kTrace && lane == 0) T.trace[3 * (int64_t)g + 1] = global_ns();
  }
}

// PSB_TRSV_KERNEL=grid|cta overrides the analysis (A/B measurements, tests of both kernels)
static int trsv_kernel_override() {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("PSB_TRSV_KERNEL");
    v = e == nullptr ? -1 : (strcmp(e, "cta") == 0 ? PSB_TRSV_CTA : (strcmp(e, "cluster") == 0 ? PSB_TRSV_CLUSTER
                                              : (strcmp(e, "grid") == 0 ? PSB_TRSV_GRID : -1)));
  }
  return v;
}

int trsv_solve(const psb_trsv* T, const double* rhs, double* x, const int32_t* rhs_map,
               double* out2, const int32_t* out_map, const int* d_skip, cudaStream_t st) {
  if (T->n == 0) return PSB_OK;
  const int forced = T->forced_kernel >= 0 ? T->forced_kernel : trsv_kernel_override();
  int kernel = forced >= 0 ? forced : T->kernel;
  if (!T->cta_ok) kernel = PSB_TRSV_GRID;       // wide levels packed with 8 lanes per row: the grid kernel's format
  if (kernel == PSB_TRSV_CTA || kernel == PSB_TRSV_CLUSTER) {
    if (T->n_far > 0) {      // far dependencies are polled in the global vector: sentinel first
      const int fg = (int)std::min<int64_t>((T->n + kBlock * 4 - 1) / (kBlock * 4), (int64_t)sm_count() * 8);
      trsv_prepare_kernel<<<std::max(fg, 1), kBlock, 0, st>>>(x, T->n, T->d_counter, d_skip);
      PSB_LAUNCH_CHECK();
    }
    return trsv_solve_cta(T, kernel == PSB_TRSV_CLUSTER ? 1 : 0, rhs, x, rhs_map, out2, out_map, d_skip, st);
  }
  const int fill_grid = (int)std::min<int64_t>((T->n + kBlock * 4 - 1) / (kBlock * 4), (int64_t)sm_count() * 8);
  trsv_prepare_kernel<<<std::max(fill_grid, 1), kBlock, 0, st>>>(x, T->n, T->d_counter, d_skip);
  PSB_LAUNCH_CHECK();
  static thread_local int per_sm = 0;
  if (per_sm == 0) {
    PSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trsv_solve_kernel<false>, kBlock, 0));
    if (per_sm < 1) per_sm = 1;
  }
  // Only the warps working a few dozen levels ahead of the wavefront do useful work; the rest
  // would just poll.  Size the grid for `kLookahead` levels of average width.
  int64_t kLookahead = 48;
  { const char* e = getenv("PSB_TRSV_LOOKAHEAD"); if (e) kLookahead = std::max(1, atoi(e)); }
  const int64_t chunks_per_level = T->n_groups / std::max(T->n_levels, 1) + 1;
  int64_t warps_needed = std::min<int64_t>(T->n_groups, chunks_per_level * kLookahead);
  int64_t grid = std::min<int64_t>((int64_t)per_sm * sm_count(), (warps_needed + kWarps - 1) / kWarps);
  TrsvView V{T->n, T->n_groups, T->d_order, T->d_diag, T->d_grp_ptr, T->d_grp_item, T->d_grp_rows,
             T->d_cols, T->d_vals, T->d_row_cnt, T->d_counter, T->d_error, T->unit_diag, T->d_trace,
             kIdleTrips, 256u};
  // A/B knobs of the back-off (defaults above)
  if (const char* e = getenv("PSB_TRSV_IDLE_TRIPS")) V.idle_trips = std::max(0, atoi(e));
  if (const char* e = getenv("PSB_TRSV_SLEEP_CAP")) V.sleep_cap = (unsigned int)std::max(32, atoi(e));
  auto kern = T->d_trace != nullptr ? trsv_solve_kernel<true> : trsv_solve_kernel<false>;
  kern<<<(int)std::max<int64_t>(grid, 1), kBlock, 0, st>>>(V, rhs, x, rhs_map, out2, out_map, d_skip);
  PSB_LAUNCH_CHECK();
  return PSB_OK;
}

// ---------------------------------------------------------------------------
// host analysis
// ---------------------------------------------------------------------------
static void free_trsv(psb_trsv* T) {
  if (!T) return;
  cudaFree(T->d_order); cudaFree(T->d_grp_ptr); cudaFree(T->d_grp_item); cudaFree(T->d_grp_rows);
  cudaFree(T->d_cols); cudaFree(T->d_wcols); cudaFree(T->d_wmeta); cudaFree(T->d_vals); cudaFree(T->d_diag); cudaFree(T->d_row_cnt); cudaFree(T->d_counter); cudaFree(T->d_error);
  delete T;
}

template <typename T>
static cudaError_t upload(T** dptr, const std::vector<T>& h, cudaStream_t st) {
  cudaError_t e = cudaMalloc((void**)dptr, std::max<size_t>(h.size(), 1) * sizeof(T));
  if (e != cudaSuccess) return e;
  if (!h.empty()) e = cudaMemcpyAsync(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st);
  return e;
}

// ---------------------------------------------------------------------------
// preconditioner objects
// ---------------------------------------------------------------------------
struct IcPrec : psb_prec {
  psb_trsv* L = nullptr;     // not owned
  psb_trsv* Lt = nullptr;
  double* tmp = nullptr;     // owned, n doubles
  ~IcPrec() override { cudaFree(tmp); }
  const char* kind() const override { return "ic"; }
  int check_error() override {
    const int a = trsv_take_error(L), b = trsv_take_error(Lt);
    return a | b;
  }
  int apply(const double* r, double* z, const int* d_skip, cudaStream_t st) override {
    int rc = trsv_solve(L, r, tmp, nullptr, nullptr, nullptr, d_skip, st);    // u = L^-1 r
    if (rc != PSB_OK) return rc;
    return trsv_solve(Lt, tmp, z, nullptr, nullptr, nullptr, d_skip, st);     // z = L^-T u
  }
};

// x = Pc U^-1 L^-1 Pr v with Pr[perm_r[i], i] = 1 and Pc[i, perm_c[i]] = 1
// (SuperLU.solve).
struct IluPrec : psb_prec {
  psb_trsv* L = nullptr;     // unit lower, not owned
  psb_trsv* U = nullptr;
  int32_t* iperm_r = nullptr;   // owned: row r of L takes v[iperm_r[r]]
  int32_t* iperm_c = nullptr;   // owned: result[iperm_c[r]] = z[r]
  double* tmp1 = nullptr;       // owned
  double* tmp2 = nullptr;
  ~IluPrec() override { cudaFree(iperm_r); cudaFree(iperm_c); cudaFree(tmp1); cudaFree(tmp2); }
  const char* kind() const override { return "ilu"; }
  int check_error() override {
    const int a = trsv_take_error(L), b = trsv_take_error(U);
    return a | b;
  }
  int apply(const double* r, double* z, const int* d_skip, cudaStream_t st) override {
    int rc = trsv_solve(L, r, tmp1, iperm_r, nullptr, nullptr, d_skip, st);
    if (rc != PSB_OK) return rc;
    return trsv_solve(U, tmp1, tmp2, nullptr, z, iperm_c, d_skip, st);
  }
};

}  // namespace psb

using namespace psb;

extern "C" int psb_trsv_create(int64_t n, const int32_t* h_rowptr, const int32_t* h_colind,
                               const double* h_vals, int lower, int unit_diag, void* stream,
                               psb_trsv_t* out) {
  PSB_REQUIRE(out != nullptr && n >= 0, PSB_ERR_ARG, "psb_trsv_create: bad argument");
  PSB_REQUIRE(n < (int64_t)INT32_MAX, PSB_ERR_UNSUPP, "psb_trsv_create: int32 row range exceeded");
  PSB_REQUIRE(h_rowptr != nullptr, PSB_ERR_ARG, "psb_trsv_create: rowptr is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  psb_trsv* T = new (std::nothrow) psb_trsv();
  PSB_REQUIRE(T != nullptr, PSB_ERR_ARG, "psb_trsv_create: out of host memory");
  T->n = n; T->lower = lower ? 1 : 0; T->unit_diag = unit_diag ? 1 : 0;

  // ---- levels: level(i) = 1 + max level over the rows i depends on ------------------
  std::vector<int32_t> level((size_t)n, 0);
  std::vector<double> diag_by_row((size_t)n, 1.0);
  std::vector<int32_t> off_count((size_t)n, 0);
  int32_t max_level = -1;
  bool missing_diag = false;
  auto visit = [&](int64_t i) {
    int32_t lv = 0, cnt = 0;
    bool have_d = false;
    for (int32_t p = h_rowptr[i]; p < h_rowptr[i + 1]; ++p) {
      const int32_t j = h_colind[p];
      if (j == i) { diag_by_row[i] = h_vals[p]; have_d = true; continue; }
      const bool dep = lower ? (j < i) : (j > i);
      if (!dep) continue;                      // entries on the wrong side are ignored
      ++cnt;
      lv = std::max(lv, level[j] + 1);
    }
    if (!have_d && !unit_diag) missing_diag = true;
    if (unit_diag) diag_by_row[i] = 1.0;
    level[i] = lv; off_count[i] = cnt;
    max_level = std::max(max_level, lv);
  };
  if (lower) for (int64_t i = 0; i < n; ++i) visit(i);
  else       for (int64_t i = n - 1; i >= 0; --i) visit(i);
  if (missing_diag) {
    delete T;
    set_error("psb_trsv_create: a row has no diagonal entry");
    return PSB_ERR_ARG;
  }
  T->n_levels = (int)(max_level + 1);

  // ---- level-major order (counting sort, ascending row id inside a level) -----------
  T->h_level_ptr.assign((size_t)T->n_levels + 1, 0);
  for (int64_t i = 0; i < n; ++i) T->h_level_ptr[level[i] + 1]++;
  for (int l = 0; l < T->n_levels; ++l) T->h_level_ptr[l + 1] += T->h_level_ptr[l];
  T->h_level_rows.assign((size_t)n, 0);
  {
    std::vector<int32_t> cursor(T->h_level_ptr.begin(), T->h_level_ptr.end() - (T->n_levels ? 1 : 0));
    for (int64_t i = 0; i < n; ++i) T->h_level_rows[cursor[level[i]]++] = (int32_t)i;
  }

  // ---- row classes ------------------------------------------------------------------------------
  // narrow levels (IC / ILUT / chains: the latency of the dependency chain decides, one-CTA / cluster
  // kernels): thread per row up to 32 entries, warp per row beyond.
  // WIDE levels (> kTrsvClusterMaxChunksPerLevel chunks of 32 rows per level: the grid kernel; e.g. the
  // leading blocks of the AMG coarse LU factors, 8 000 rows per level, 46 entries per row): a lane
  // walks its entries four L2 round trips at a time, so a 32-entry row on one lane is 8 trips deep and
  // every level waited for its longest thread-per-row row (measured 4 - 5 us per level,
  // profiles/round2_amg.md).  There: thread per row up to 8 entries, EIGHT lanes per row (four rows
  // per warp) up to 128, warp per row beyond (thresholds swept on the Bratu-2048^2 coarse factors:
  // (8, 128) 0.89 ms per coarse solve, (8, 64) 0.92, (4, 32) 1.02, thread per row up to 32: 1.02).
  const bool wide = (double)n / 32.0 / std::max(T->n_levels, 1) > kTrsvClusterMaxChunksPerLevel &&
                    getenv("PSB_TRSV_NO_SUBWARP") == nullptr;
  int kShortRow = wide ? 8 : 32;               // thread per row up to this many off-diagonal entries
  int kLongRow = wide ? 128 : 32;              // more than this -> warp per row; in between: 8 lanes per row
  if (wide) {                                  // A/B measurements
    if (const char* e = getenv("PSB_TRSV_SHORT")) kShortRow = std::max(1, std::min(32, atoi(e)));
    if (const char* e = getenv("PSB_TRSV_LONG")) kLongRow = std::max(kShortRow, atoi(e));
  }
  auto cls = [&](int32_t i) { return off_count[i] <= kShortRow ? 0 : (off_count[i] <= kLongRow ? 1 : 2); };
  // ---- processing order: inside a level the short rows first, then the medium, then the long ones
  std::vector<int32_t> order((size_t)n);
  {
    int64_t q = 0;
    for (int l = 0; l < T->n_levels; ++l) {
      const int32_t a0 = T->h_level_ptr[l], b0 = T->h_level_ptr[l + 1];
      for (int c = 0; c < 3; ++c)
        for (int32_t p = a0; p < b0; ++p) { const int32_t i = T->h_level_rows[p]; if (cls(i) == c) order[q++] = i; }
    }
  }
  // ---- chunks: up to 32 consecutive short rows (SELL-32), up to 4 medium rows on 8 lanes each
  // (grp_rows = -rows), or one long row over 32 lanes (grp_rows = 0) ---------------------------------
  std::vector<int64_t> grp_ptr(1, 0);
  std::vector<int32_t> grp_item, grp_rows;
  T->n_subwarp = 0;
  for (int64_t q = 0; q < n;) {
    const int c0 = cls(order[q]);
    const int32_t lv = level[order[q]];
    if (c0 == 2) {
      const int64_t w = (off_count[order[q]] + 31) / 32;
      grp_item.push_back((int32_t)q); grp_rows.push_back(0);
      grp_ptr.push_back(grp_ptr.back() + w * 32);
      ++T->n_long;
      ++q;
    } else if (c0 == 1) {
      int cnt = 0;
      int32_t w = 0;
      while (q + cnt < n && cnt < 4 && cls(order[q + cnt]) == 1 && level[order[q + cnt]] == lv) {
        w = std::max(w, (off_count[order[q + cnt]] + 7) / 8);
        ++cnt;
      }
      grp_item.push_back((int32_t)q); grp_rows.push_back(-cnt);
      grp_ptr.push_back(grp_ptr.back() + (int64_t)w * 32);
      ++T->n_subwarp;
      q += cnt;
    } else {
      // a chunk never spans two levels: its rows are independent of each other, so the lanes of
      // a warp never wait for one another (the one-CTA kernel relies on it)
      int cnt = 0;
      int32_t w = 0;
      while (q + cnt < n && cnt < 32 && cls(order[q + cnt]) == 0 && level[order[q + cnt]] == lv) {
        w = std::max(w, off_count[order[q + cnt]]);
        ++cnt;
      }
      grp_item.push_back((int32_t)q); grp_rows.push_back(cnt);
      grp_ptr.push_back(grp_ptr.back() + (int64_t)w * 32);
      q += cnt;
    }
  }
  T->n_groups = (int)grp_item.size();
  T->nnz_packed = grp_ptr.back();
  std::vector<int32_t> cols((size_t)T->nnz_packed, -1);
  std::vector<double> vals((size_t)T->nnz_packed, 0.0);
  std::vector<double> diag((size_t)n, 1.0);
  std::vector<int32_t> row_cnt((size_t)n, 0);
  int64_t nnz_off = 0;
  std::vector<std::pair<int32_t, int32_t>> deps;
  for (int g = 0; g < T->n_groups; ++g) {
    const int nrows = grp_rows[g] == 0 ? 1 : std::abs(grp_rows[g]);
    const int64_t width = (grp_ptr[g + 1] - grp_ptr[g]) / 32;
    for (int l = 0; l < nrows; ++l) {
      const int64_t q = grp_item[g] + l;
      const int32_t i = order[q];
      row_cnt[q] = off_count[i];
      diag[q] = unit_diag ? 1.0 : 1.0 / diag_by_row[i];   // invdiag = 1/diag like scipy (IEEE division, once)
      // the dependencies of a row in the order in which they become available: ascending level,
      // ties in stored order -- the newest one is the last operand (oracle.precond.trsv_rowwise)
      deps.clear();
      for (int32_t p = h_rowptr[i]; p < h_rowptr[i + 1]; ++p) {
        const int32_t j = h_colind[p];
        if (j == i) continue;
        const bool dep = lower ? (j < i) : (j > i);
        if (dep) deps.push_back(std::make_pair(level[j], p));
      }
      std::sort(deps.begin(), deps.end());
      int k = 0;
      for (const auto& dp : deps) {
        const int32_t p = dp.second;
        // short rows: lane l, right-aligned (the row's last entry in the chunk's last entry row);
        // long row: entry k spread as (k / 32, k % 32)
        // medium row l of a sub-warp chunk: entry k on lane 8 l + k % 8, entry row k / 8
        const int64_t pos = grp_rows[g] == 0 ? grp_ptr[g] + k
                          : grp_rows[g] < 0 ? grp_ptr[g] + (int64_t)(k / 8) * 32 + 8 * l + (k % 8)
                                            : grp_ptr[g] + (int64_t)(width - (int64_t)deps.size() + k) * 32 + l;
        cols[pos] = h_colind[p];
        vals[pos] = h_vals[p];
        ++k; ++nnz_off;
      }
    }
  }
  T->nnz_off = nnz_off;

  // ---- the same entries for the shared-memory window kernel: dependencies as positions ------
  // near (<= near_limit positions back): polled in the window; far: polled in the global vector
  // Shared memory of the one-CTA kernel = window + 16 staging buffers (384 B per entry row).  A
  // vector of up to 16 384 rows lives in the window whole; beyond that the window wraps around
  // and is sized so that a whole chunk of typical length can be staged: 16 384 slots while the
  // chunks are short, 8 192 slots (26 entries per lane staged) for the longer rows of IC factors.
  // With more than kTrsvCtaMaxChunksPerLevel chunks per level the cluster kernel is the candidate:
  // its 64 warps need a reuse slack of 32 (2 * 128 + 2) positions, i.e. the large window.
  const double mean_len = (double)T->nnz_packed / 32.0 / std::max(T->n_groups, 1);
  const double cpl = (double)T->n_groups / std::max(T->n_levels, 1);
  const bool cluster_cand = cpl > kTrsvCtaMaxChunksPerLevel && cpl <= kTrsvClusterMaxChunksPerLevel;
  if (n <= kTrsvMaxSlots) { T->wslots = 32; while (T->wslots < n) T->wslots <<= 1; }
  else T->wslots = (cluster_cand || mean_len <= 12.0) ? kTrsvMaxSlots : kTrsvMaxSlots / 2;
  T->stage_len = (int)std::min<int64_t>(32, (kTrsvSmemBudget - 128 - (int64_t)T->wslots * 8) / (kTrsvCtaWarps * 384));
  const int ahead = cluster_cand ? 2 * kTrsvCtaWarps * kTrsvClusterSize : kTrsvAhead;
  const int64_t near_limit = n <= T->wslots ? n : (int64_t)T->wslots - 32 * (2 * ahead + 2);
  T->cluster_ok = (n <= T->wslots || cluster_cand) ? 1 : 0;
  // entry encoding: byte offset into the window (near), the always-zero slot right behind the window
  // (padding), or -(row) - 2 (far)
  const int32_t zero_off = T->wslots * 8;
  const int32_t wmask = T->wslots - 1;
  std::vector<int32_t> wcols((size_t)T->nnz_packed, zero_off);
  std::vector<int32_t> wmeta((size_t)T->n_groups * 4);
  {
    std::vector<int32_t> pos_of((size_t)n);
    for (int64_t q = 0; q < n; ++q) pos_of[order[q]] = (int32_t)q;
    for (int g = 0; g < T->n_groups; ++g) {
      const int nrows = grp_rows[g] == 0 ? 1 : grp_rows[g];
      const int64_t width = (grp_ptr[g + 1] - grp_ptr[g]) / 32;
      int has_far = 0;
      for (int l = 0; l < nrows && T->n_subwarp == 0; ++l) {   // sub-warp chunks: grid kernel only
        const int64_t q = grp_item[g] + l;
        for (int64_t k = 0; k < (grp_rows[g] == 0 ? width * 32 : width); ++k) {
          const int64_t at = grp_rows[g] == 0 ? grp_ptr[g] + k : grp_ptr[g] + k * 32 + l;
          const int32_t j = cols[at];
          if (j < 0) continue;
          const int64_t dist = q - pos_of[j];
          T->max_dist = std::max(T->max_dist, dist);
          if (dist <= near_limit) wcols[at] = (pos_of[j] & wmask) * 8;
          else { wcols[at] = -j - 2; ++T->n_far; has_far = 1; }
        }
      }
      wmeta[4 * (size_t)g + 0] = (int32_t)(uint32_t)(grp_ptr[g] & 0xffffffffll);
      wmeta[4 * (size_t)g + 1] = (int32_t)(grp_ptr[g] >> 32);
      wmeta[4 * (size_t)g + 2] = grp_item[g];
      wmeta[4 * (size_t)g + 3] = (int32_t)(std::max(grp_rows[g], 0) | (has_far << 6) | (width << 7));
    }
  }
  // One CTA wins while the solve is bound by the latency of the dependency chain and a level has
  // at most ~3 chunks (the critical warps and the next spinners keep their warp schedulers); up to
  // ~12 chunks per level the 4-CTA cluster covers the wavefront with 64 warps at the price of
  // the DSMEM hand-over; wider levels, or many dependencies outside the window, go to the grid.
  // Thresholds measured on B200 (profiles/round1d_trsv.md, round1g_cluster.md).
  {
    const bool local = (double)T->n_far <= 0.02 * (double)std::max<int64_t>(T->nnz_off, 1);
    // The cluster pays off for SHORT rows (Gauss-Seidel triangles, 2 entries per row: 0.77 - 0.87 us per
    // level against 0.93 - 1.7 on one CTA at 4.5 - 8.5 chunks per level).  With the rows of an IC factor
    // (~17 entries) one CTA stays ahead further up: IC 1536^2, 4.3 chunks per level, 0.87 against 1.15 us
    // per level; IC 2048^2, 6.4 chunks per level, 1.22 against 1.30 (profiles/round2_amg.md).
    const bool short_rows = (double)T->nnz_off <= 6.0 * (double)std::max<int64_t>(n, 1);
    if (T->n_subwarp > 0) { T->cluster_ok = 0; T->cta_ok = 0; }
    if (!local || T->n_subwarp > 0) T->kernel = PSB_TRSV_GRID;
    else if (cpl <= kTrsvCtaMaxChunksPerLevel) T->kernel = PSB_TRSV_CTA;
    else if (cluster_cand && !short_rows && cpl <= 8.0) T->kernel = PSB_TRSV_CTA;
    else if (cluster_cand) T->kernel = PSB_TRSV_CLUSTER;
    else T->kernel = PSB_TRSV_GRID;
  }

  cudaError_t e = upload(&T->d_order, order, st);
  if (e == cudaSuccess) e = upload(&T->d_wcols, wcols, st);
  if (e == cudaSuccess) e = upload(&T->d_wmeta, wmeta, st);
  if (e == cudaSuccess) e = upload(&T->d_grp_ptr, grp_ptr, st);
  if (e == cudaSuccess) e = upload(&T->d_grp_item, grp_item, st);
  if (e == cudaSuccess) e = upload(&T->d_grp_rows, grp_rows, st);
  if (e == cudaSuccess) e = upload(&T->d_cols, cols, st);
  if (e == cudaSuccess) e = upload(&T->d_vals, vals, st);
  if (e == cudaSuccess) e = upload(&T->d_diag, diag, st);
  if (e == cudaSuccess) e = upload(&T->d_row_cnt, row_cnt, st);
  if (e == cudaSuccess) e = cudaMalloc((void**)&T->d_counter, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMalloc((void**)&T->d_error, sizeof(int));
  if (e == cudaSuccess) e = cudaMemsetAsync(T->d_error, 0, sizeof(int), st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);      // host vectors die at return
  if (e != cudaSuccess) {
    set_error("psb_trsv_create: %s", cudaGetErrorString(e));
    free_trsv(T);
    return PSB_ERR_CUDA;
  }
  *out = T;
  return PSB_OK;
}

extern "C" int psb_trsv_destroy(psb_trsv_t T) {
  free_trsv(T);
  return PSB_OK;
}

extern "C" int psb_trsv_info(psb_trsv_t T, int64_t info[8]) {
  PSB_REQUIRE(T && info, PSB_ERR_ARG, "psb_trsv_info: NULL argument");
  info[0] = T->n; info[1] = T->n_levels; info[2] = T->nnz_off; info[3] = T->nnz_packed;
  info[4] = T->lower; info[5] = T->unit_diag; info[6] = T->n_groups; info[7] = T->n_long;
  return PSB_OK;
}

extern "C" int psb_trsv_info2(psb_trsv_t T, int64_t info[8]) {
  PSB_REQUIRE(T && info, PSB_ERR_ARG, "psb_trsv_info2: NULL argument");
  info[0] = T->kernel; info[1] = T->wslots; info[2] = T->n_far; info[3] = T->max_dist;
  info[4] = T->forced_kernel; info[5] = T->stage_len; info[6] = T->cluster_ok; info[7] = 0;
  return PSB_OK;
}

extern "C" int psb_trsv_set_trace(psb_trsv_t T, long long* d_trace) {
  PSB_REQUIRE(T != nullptr, PSB_ERR_ARG, "psb_trsv_set_trace: NULL argument");
  T->d_trace = d_trace;
  return PSB_OK;
}

extern "C" int psb_trsv_set_kernel(psb_trsv_t T, int kernel) {
  PSB_REQUIRE(T != nullptr, PSB_ERR_ARG, "psb_trsv_set_kernel: NULL argument");
  PSB_REQUIRE(kernel >= -1 && kernel <= PSB_TRSV_CLUSTER, PSB_ERR_ARG, "psb_trsv_set_kernel: unknown kernel");
  PSB_REQUIRE(kernel != PSB_TRSV_CLUSTER || T->cluster_ok, PSB_ERR_UNSUPP,
              "psb_trsv_set_kernel: this factor was not analysed for the cluster kernel");
  PSB_REQUIRE(kernel != PSB_TRSV_CTA || T->cta_ok, PSB_ERR_UNSUPP,
              "psb_trsv_set_kernel: this factor has wide levels packed for the grid kernel (8 lanes per row)");
  T->forced_kernel = kernel;
  return PSB_OK;
}

extern "C" int psb_trsv_get_levels(psb_trsv_t T, int32_t* h_level_ptr, int32_t* h_level_rows) {
  PSB_REQUIRE(T && h_level_ptr && h_level_rows, PSB_ERR_ARG, "psb_trsv_get_levels: NULL argument");
  std::copy(T->h_level_ptr.begin(), T->h_level_ptr.end(), h_level_ptr);
  std::copy(T->h_level_rows.begin(), T->h_level_rows.end(), h_level_rows);
  return PSB_OK;
}

extern "C" int psb_trsv_solve(psb_trsv_t T, const double* d_b, double* d_x, void* stream) {
  PSB_REQUIRE(T && d_b && d_x, PSB_ERR_ARG, "psb_trsv_solve: NULL argument");
  PSB_REQUIRE(d_b != d_x, PSB_ERR_ARG, "psb_trsv_solve: x must not alias b");
  return trsv_solve(T, d_b, d_x, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int psb_trsv_error(psb_trsv_t T, int32_t* h_flag) {
  PSB_REQUIRE(T && h_flag, PSB_ERR_ARG, "psb_trsv_error: NULL argument");
  *h_flag = trsv_take_error(T);            // reported once, then cleared
  return PSB_OK;
}

extern "C" int psb_ic_create(psb_trsv_t L, psb_trsv_t Lt, psb_prec_t* out) {
  PSB_REQUIRE(L && Lt && out, PSB_ERR_ARG, "psb_ic_create: NULL argument");
  PSB_REQUIRE(L->n == Lt->n && L->lower && !Lt->lower, PSB_ERR_ARG,
              "psb_ic_create: need a lower and an upper factor of the same size");
  IcPrec* P = new (std::nothrow) IcPrec();
  PSB_REQUIRE(P != nullptr, PSB_ERR_ARG, "psb_ic_create: out of host memory");
  P->n = L->n; P->L = L; P->Lt = Lt;
  cudaError_t e = cudaMalloc((void**)&P->tmp, std::max<int64_t>(L->n, 1) * sizeof(double));
  if (e != cudaSuccess) { delete P; set_error("psb_ic_create: %s", cudaGetErrorString(e)); return PSB_ERR_CUDA; }
  *out = P;
  return PSB_OK;
}

extern "C" int psb_ilu_create(psb_trsv_t L, psb_trsv_t U, const int32_t* h_perm_r,
                              const int32_t* h_perm_c, void* stream, psb_prec_t* out) {
  PSB_REQUIRE(L && U && h_perm_r && h_perm_c && out, PSB_ERR_ARG, "psb_ilu_create: NULL argument");
  PSB_REQUIRE(L->n == U->n && L->lower && !U->lower, PSB_ERR_ARG,
              "psb_ilu_create: need a lower and an upper factor of the same size");
  const int64_t n = L->n;
  IluPrec* P = new (std::nothrow) IluPrec();
  PSB_REQUIRE(P != nullptr, PSB_ERR_ARG, "psb_ilu_create: out of host memory");
  P->n = n; P->L = L; P->U = U;
  // (Pr v)[perm_r[i]] = v[i]  ->  row r of the L solve reads v[iperm_r[r]]
  // (Pc z)[i] = z[perm_c[i]]  ->  row r of the U solve writes result[iperm_c[r]]
  std::vector<int32_t> ipr((size_t)n), ipc((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    if (h_perm_r[i] < 0 || h_perm_r[i] >= n || h_perm_c[i] < 0 || h_perm_c[i] >= n) {
      delete P; set_error("psb_ilu_create: permutation entry out of range"); return PSB_ERR_ARG;
    }
    ipr[h_perm_r[i]] = (int32_t)i;
    ipc[h_perm_c[i]] = (int32_t)i;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = upload(&P->iperm_r, ipr, st);
  if (e == cudaSuccess) e = upload(&P->iperm_c, ipc, st);
  if (e == cudaSuccess) e = cudaMalloc((void**)&P->tmp1, std::max<int64_t>(n, 1) * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&P->tmp2, std::max<int64_t>(n, 1) * sizeof(double));
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { delete P; set_error("psb_ilu_create: %s", cudaGetErrorString(e)); return PSB_ERR_CUDA; }
  *out = P;
  return PSB_OK;
}

extern "C" int psb_prec_apply(psb_prec_t P, const double* d_r, double* d_z, void* stream) {
  PSB_REQUIRE(P && d_r && d_z, PSB_ERR_ARG, "psb_prec_apply: NULL argument");
  PSB_REQUIRE(d_r != d_z, PSB_ERR_ARG, "psb_prec_apply: z must not alias r");
  return P->apply(d_r, d_z, nullptr, (cudaStream_t)stream);
}

extern "C" int psb_prec_destroy(psb_prec_t P) {
  delete P;
  return PSB_OK;
}
