// CSR SpMV for sm_100a: y = A x (+ fused epilogues), fp64 values / int32 indices.
//
// Replaces scipy's sequential csr_matvec behind `A*x`
// (PySolvers/Linear/IterativeLinearSolver.py:104 and every call site listed in
// SURVEY.md section 2.1).  Two kernels, picked once per matrix from row-length
// statistics gathered at psb_csr_create:
//
//  STREAM (short rows: stencils, FE matrices, AMG operators) -- the default
//    A persistent grid of CTAs walks tiles of 256 (or 512) consecutive rows.  The
//    tile's nonzeros are one CONTIGUOUS slice of vals/colind, so one elected thread
//    stages rowptr/colind/vals of the NEXT tile into shared memory with three bulk
//    async copies (cp.async.bulk + mbarrier complete_tx, the TMA engine: SASS UBLKCP)
//    while all threads work on the current one: double-buffered, no LSU wavefronts
//    and no registers spent on the streaming part, L2 evict-first for the matrix so
//    x stays cached.  Then one thread per row reads its (col, val) pairs from shared
//    memory, gathers x[col] (adjacent rows of a banded matrix hit adjacent x: the
//    gathers coalesce) and accumulates IN STORED ORDER starting from +0 -- the same
//    sequence of roundings as scipy's csr_matvec, so y is bit-identical to the
//    reference.  HBM traffic = 12 B/nnz + 4 B/row (rowptr) + x once + y once.
//
//  STREAM_LSU: the first-generation variant (coalesced 128-bit LDG of the tile,
//    products parked in shared memory); kept for arrays that are not 16-byte aligned
//    and for A/B timing.  ncu showed it L1-wavefront bound (profiles/round1_notes.md).
//
//  VECTOR (long or very uneven rows)
//    2..32 lanes per row, strided accumulate, shuffle reduction.
//
// The fused x.y (p'Ap of PCG, PCGSolver.py:113) is reduced deterministically:
// fixed tree inside the CTA, one partial per CTA, and the CTA that draws the last
// ticket adds the partials in index order.
#include "spmv.cuh"
#include "spmv_bulk.cuh"

#include <algorithm>
#include <new>

namespace psb {

// ---------------------------------------------------------------------------
// row statistics (one pass over rowptr)
// ---------------------------------------------------------------------------
__global__ void csr_stats_kernel(const int* __restrict__ rowptr, int64_t n_rows,
                                 int* __restrict__ stats) {
  int m_row = 0, m_t256 = 0, m_t512 = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rows;
       i += (int64_t)gridDim.x * blockDim.x) {
    int a = rowptr[i];
    m_row = max(m_row, rowptr[i + 1] - a);
    if ((i & 255) == 0) {
      int64_t e = min(i + 256, n_rows);
      m_t256 = max(m_t256, rowptr[e] - a);
    }
    if ((i & 511) == 0) {
      int64_t e = min(i + 512, n_rows);
      m_t512 = max(m_t512, rowptr[e] - a);
    }
  }
  atomicMax(&stats[0], m_row);
  atomicMax(&stats[1], m_t256);
  atomicMax(&stats[2], m_t512);
}

// most nonzeros in a tile of `tile_rows` consecutive rows (tiles start at row 0 of the view)
__global__ void csr_tile_stats_kernel(const int* __restrict__ rowptr, int64_t n_rows, int tile_rows,
                                      int* __restrict__ out) {
  int m = 0;
  const int64_t n_tiles = (n_rows + tile_rows - 1) / tile_rows;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n_tiles;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = t * tile_rows, e = min(a + tile_rows, n_rows);
    m = max(m, rowptr[e] - rowptr[a]);
  }
  atomicMax(out, m);
}

static int g_cols16 = -1;     // 16-bit column distances for banded matrices: -1 ask PSB_SPMV_C16, 0 off, 1 on

// colind[k] - row as int16 for every entry; *fail is raised when one does not fit
__global__ void csr_delta16_kernel(const int* __restrict__ rowptr, const int* __restrict__ colind,
                                   int64_t n_rows, short* __restrict__ out, int* __restrict__ fail) {
  bool bad = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rows;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int a = rowptr[i], b = rowptr[i + 1];
    for (int k = a; k < b; ++k) {
      const long long d = (long long)colind[k] - i;
      if (d < -32768 || d > 32767) bad = true;
      out[k] = (short)d;
    }
  }
  if (bad) *fail = 1;
}

template <int EPI>
__device__ __forceinline__ void finish_dot(double acc, const psb_csr A, const EpiArgs& ea,
                                           double* scratch) {
  if (EPI != EPI_DOT && EPI != EPI_DOT_PUP && EPI != EPI_RESID_NORM) return;
  double t = block_sum(acc, scratch);
  if (threadIdx.x == 0) A.partials[blockIdx.x] = t;
  if (last_block(A.ticket)) {
    double total = sum_partials(A.partials, gridDim.x, scratch);
    if (EPI == EPI_RESID_NORM) {               // end of a V-cycle (VCycleSolver.py:87-91)
      if (threadIdx.x == 0 && ea.amg_state == nullptr && ea.dot != nullptr) {
        // row-partitioned: only this rank's part of ||r||^2; the cycle is finished after the all-reduce
        *ea.dot = ea.dot_accumulate ? *ea.dot + total : total;
      }
      if (threadIdx.x == 0 && ea.amg_state != nullptr) {
        AmgState* st = ea.amg_state;
        const double nr = sqrt(total);
        const int k = st->cycles;
        st->norm_r = nr;
        if (ea.amg_hist != nullptr) ea.amg_hist[k] = nr;
        st->cycles = k + 1;
        if (nr < st->tau * st->norm_b) { st->status = PSB_CONVERGED; st->skip = 1; }
        else if (k + 1 >= st->maxiter) { st->skip = 1; }
      }
      return;
    }
    if (threadIdx.x == 0) {
      if (ea.dot_accumulate) total = *ea.dot + total;
      *ea.dot = total;
      for (int q = 0; q < ea.push_n; ++q) peer_push(ea.push_slots[q], total, ea.push_epoch);   // NVLink
    }
  }
}

// ---------------------------------------------------------------------------
// STREAM kernel
// ---------------------------------------------------------------------------
template <int EPI, int RPT, bool VEC>
__global__ void __launch_bounds__(kBlock)
spmv_stream_kernel(const psb_csr A, const double* __restrict__ x, double* __restrict__ y,
                   const EpiArgs ea, const int* __restrict__ d_skip, int prod_cap) {
  constexpr int R = kBlock * RPT;
  extern __shared__ double smem[];
  double* prod = smem;                                   // [prod_cap]
  int*    rp   = reinterpret_cast<int*>(smem + prod_cap); // [R + 1]
  __shared__ double scratch[kWarps];

  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;

  const int tid = threadIdx.x;
  const int64_t n_tiles = (A.n_rows + R - 1) / R;
  double acc = 0.0;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * R;
    const int nr = (int)min((int64_t)R, A.n_rows - row0);
    for (int i = tid; i <= nr; i += kBlock) rp[i] = ld_stream_i(A.rowptr + row0 + i);
    __syncthreads();
    const int s = rp[0];
    const int len = rp[nr] - s;

    // ---- phase 1: stream the tile's nonzeros, products -> shared memory ----
    if (VEC) {
      const int base0 = s & ~3;                       // 16-byte aligned start
      const int nquads = (s - base0 + len + 3) >> 2;
      for (int q = tid; q < nquads; q += kBlock) {
        const int g = base0 + (q << 2);
        int c[4];
        double v[4];
        if ((int64_t)g + 4 <= A.nnz) {
          int4 ci = ld_stream_i4(A.colind + g);
          double2 v01 = ld_stream2(A.vals + g);
          double2 v23 = ld_stream2(A.vals + g + 2);
          c[0] = ci.x; c[1] = ci.y; c[2] = ci.z; c[3] = ci.w;
          v[0] = v01.x; v[1] = v01.y; v[2] = v23.x; v[3] = v23.y;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            bool ok = (int64_t)g + j < A.nnz;
            c[j] = ok ? ld_stream_i(A.colind + g + j) : 0;
            v[j] = ok ? ld_stream(A.vals + g + j) : 0.0;
          }
        }
        double xv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = g + j - s;
          xv[j] = (k >= 0 && k < len) ? __ldg(x + c[j]) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = g + j - s;
          if (k >= 0 && k < len) prod[k] = v[j] * xv[j];
        }
      }
    } else {
      for (int i0 = tid; i0 < len; i0 += 4 * kBlock) {
        int c[4];
        double v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + j * kBlock;
          const bool ok = i < len;
          c[j] = ok ? ld_stream_i(A.colind + s + i) : 0;
          v[j] = ok ? ld_stream(A.vals + s + i) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + j * kBlock;
          if (i < len) prod[i] = v[j] * __ldg(x + c[j]);
        }
      }
    }
    __syncthreads();

    // ---- phase 2: one thread per row, sequential sum in stored order --------
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int lr = tid + j * kBlock;
      if (lr < nr) {
        const int a = rp[lr] - s, b = rp[lr + 1] - s;
        double sum = 0.0;
        for (int k = a; k < b; ++k) sum += prod[k];
        epilogue<EPI>(A.row_off + row0 + lr, sum, x, y, ea, acc);
      }
    }
    __syncthreads();
  }
  finish_dot<EPI>(acc, A, ea, scratch);
}

// ---------------------------------------------------------------------------
// STREAM kernel, bulk-async staged (default): thin wrapper over bulk_pass (spmv_bulk.cuh)
// ---------------------------------------------------------------------------
template <int EPI, int RPT, bool C16>
__global__ void __launch_bounds__(kBlock)
spmv_bulk_kernel(const psb_csr A, const double* x, double* y, const EpiArgs ea,
                 const int* __restrict__ d_skip, int cap_v, int cap_c) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double scratch[kWarps];
  __shared__ __align__(8) BulkShared bsh;
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;
  double beta = 0.0;
  if (EPI == EPI_DOT_PUP) beta = ld_cg(ea.beta_num) / ld_cg(ea.beta_den);     // PCGSolver.py:135
  BulkPipe P;
  bulk_pipe_init(P, smem_raw, &bsh, cap_v, cap_c);
  double acc = 0.0;
  // all loads of a row of <= 8 entries in flight at once; the fused direction update doubles the
  // loads per gather, so it stays at 4
  bulk_pass<EPI, RPT, C16, (EPI == EPI_DOT_PUP ? 4 : 8)>(A, x, y, ea, beta, P, acc);
  finish_dot<EPI>(acc, A, ea, scratch);
}

// ---------------------------------------------------------------------------
// VECTOR kernel: W lanes per row
// ---------------------------------------------------------------------------
template <int EPI, int W>
__global__ void __launch_bounds__(kBlock)
spmv_vector_kernel(const psb_csr A, const double* __restrict__ x, double* __restrict__ y,
                   const EpiArgs ea, const int* __restrict__ d_skip) {
  __shared__ double scratch[kWarps];
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;
  constexpr int ROWS = kBlock / W;
  const int lane = threadIdx.x % W;
  const int sub  = threadIdx.x / W;
  double acc = 0.0;
  const int64_t n_groups = (A.n_rows + ROWS - 1) / ROWS;
  for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int64_t row = grp * ROWS + sub;
    double sum = 0.0;
    if (row < A.n_rows) {
      const int a = A.rowptr[row], b = A.rowptr[row + 1];
      int k = a + lane;
      // long rows: four entries per lane in flight (values, columns, then the four gathers), added
      // in the same order as the plain loop
      for (; k + 3 * W < b; k += 4 * W) {
        const double v0 = ld_stream(A.vals + k), v1 = ld_stream(A.vals + k + W);
        const double v2 = ld_stream(A.vals + k + 2 * W), v3 = ld_stream(A.vals + k + 3 * W);
        const int c0 = ld_stream_i(A.colind + k), c1 = ld_stream_i(A.colind + k + W);
        const int c2 = ld_stream_i(A.colind + k + 2 * W), c3 = ld_stream_i(A.colind + k + 3 * W);
        const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
        sum += v0 * x0; sum += v1 * x1; sum += v2 * x2; sum += v3 * x3;
      }
      for (; k < b; k += W)
        sum += ld_stream(A.vals + k) * __ldg(x + ld_stream_i(A.colind + k));
    }
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o, W);
    if (row < A.n_rows && lane == 0) epilogue<EPI>(A.row_off + row, sum, x, y, ea, acc);
  }
  finish_dot<EPI>(acc, A, ea, scratch);
}

// ---------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------
struct LaunchCfg { int grid; size_t smem; };

template <typename K>
static int occupancy_per_sm(K kernel, size_t smem, int* per_sm_out) {
  if (smem > 48 * 1024)
    PSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  PSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, smem));
  *per_sm_out = per_sm < 1 ? 1 : per_sm;
  return PSB_OK;
}

static inline int grid_for(int per_sm, int64_t work_items, int max_grid) {
  int64_t g = (int64_t)per_sm * sm_count();
  g = std::min<int64_t>(g, work_items);
  g = std::min<int64_t>(g, max_grid);
  return (int)std::max<int64_t>(g, 1);
}

template <int EPI, int RPT, bool VEC>
static int launch_stream(const psb_csr* A, const double* x, double* y, const EpiArgs& ea,
                         const int* d_skip, cudaStream_t st) {
  constexpr int R = kBlock * RPT;
  const int cap = (A->max_tile_nnz[RPT - 1] + 1) & ~1;   // keep rp[] 8-byte aligned
  const size_t smem = (size_t)cap * sizeof(double) + (size_t)(R + 1) * sizeof(int);
  static thread_local int per_sm = 0;
  static thread_local size_t cached_smem = 0;
  const int64_t tiles = (A->n_rows + R - 1) / R;
  if (per_sm == 0 || cached_smem != smem) {
    int rc = occupancy_per_sm(spmv_stream_kernel<EPI, RPT, VEC>, smem, &per_sm);
    if (rc != PSB_OK) return rc;
    cached_smem = smem;
  }
  spmv_stream_kernel<EPI, RPT, VEC><<<grid_for(per_sm, tiles, A->max_grid), kBlock, smem, st>>>(
      *A, x, y, ea, d_skip, cap);
  PSB_LAUNCH_CHECK();
  return PSB_OK;
}

static inline void bulk_caps(const psb_csr* A, int rpt, int* cap_v, int* cap_c, size_t* smem) {
  const int mt = A->max_tile_nnz[rpt - 1];
  *cap_v = (mt + 2 + 1) & ~1;                  // tile nnz + alignment slack, even
  *cap_c = (mt + 6 + 3) & ~3;                  // multiple of 4 ints keeps rp[] 16-byte aligned
  const size_t stage = (size_t)*cap_v * 8 + (size_t)*cap_c * 4 + (size_t)(kBlock * rpt + 4) * 4;
  *smem = 2 * stage;
}

template <int EPI, int RPT, bool C16>
static int launch_bulk_t(const psb_csr* A, const double* x, double* y, const EpiArgs& ea,
                         const int* d_skip, cudaStream_t st) {
  constexpr int R = kBlock * RPT;
  int cap_v, cap_c;
  size_t smem;
  bulk_caps(A, RPT, &cap_v, &cap_c, &smem);
  static thread_local int per_sm = 0;
  static thread_local size_t cached_smem = 0;
  const int64_t tiles = (A->n_rows + R - 1) / R;
  if (per_sm == 0 || cached_smem != smem) {
    int rc = occupancy_per_sm(spmv_bulk_kernel<EPI, RPT, C16>, smem, &per_sm);
    if (rc != PSB_OK) return rc;
    cached_smem = smem;
  }
  spmv_bulk_kernel<EPI, RPT, C16><<<grid_for(per_sm, tiles, A->max_grid), kBlock, smem, st>>>(
      *A, x, y, ea, d_skip, cap_v, cap_c);
  PSB_LAUNCH_CHECK();
  return PSB_OK;
}

template <int EPI, int RPT>
static int launch_bulk(const psb_csr* A, const double* x, double* y, const EpiArgs& ea,
                       const int* d_skip, cudaStream_t st) {
  if (A->colind16 != nullptr) return launch_bulk_t<EPI, RPT, true>(A, x, y, ea, d_skip, st);
  return launch_bulk_t<EPI, RPT, false>(A, x, y, ea, d_skip, st);
}

template <int EPI, int W>
static int launch_vector(const psb_csr* A, const double* x, double* y, const EpiArgs& ea,
                         const int* d_skip, cudaStream_t st) {
  constexpr int ROWS = kBlock / W;
  static thread_local int per_sm = 0;
  const int64_t groups = (A->n_rows + ROWS - 1) / ROWS;
  if (per_sm == 0) {
    int rc = occupancy_per_sm(spmv_vector_kernel<EPI, W>, 0, &per_sm);
    if (rc != PSB_OK) return rc;
  }
  spmv_vector_kernel<EPI, W><<<grid_for(per_sm, groups, A->max_grid), kBlock, 0, st>>>(*A, x, y, ea, d_skip);
  PSB_LAUNCH_CHECK();
  return PSB_OK;
}

// ---------------------------------------------------------------------------
// MERGE kernel (merge-path SpMV, Merrill & Garland): skewed row-length histograms
// ---------------------------------------------------------------------------
// A few very long rows among short ones defeat both layouts above: a thread (or sub-warp) per row
// serialises on the long rows, and a STREAM tile that contains one does not fit shared memory.  Here
// the WORK is split evenly instead of the rows: the n_rows row-ends and the nnz entries are merged
// into one list of n_rows + nnz items, every CTA takes kMergeTile consecutive items and every thread
// kMergeItems of them, wherever the row boundaries fall (two binary searches per CTA, one per thread).
// A thread adds the products of its segment in stored order; the pieces of a row that spans threads
// are combined by a segmented scan in thread order, the pieces of a row that spans CTAs by a second
// small kernel in CTA order: deterministic, but associated differently from the sequential sum, so
// this kind agrees with scipy to rounding, not bit for bit (like VECTOR).  The epilogue runs as a
// separate pass over the rows (the sums are complete only after the cross-CTA fix-up).
constexpr int kMergeItems = 7;
constexpr int kMergeTile = kBlock * kMergeItems;

// first coordinate (row, nz) on diagonal d of the merge of row_end[0..n_rows) with 0..nnz)
__device__ __forceinline__ void merge_search(int64_t d, const int* __restrict__ row_end, int64_t n_rows, int64_t nnz,
                                             int64_t& row, int64_t& nz) {
  int64_t lo = d > nnz ? d - nnz : 0, hi = d < n_rows ? d : n_rows;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)row_end[mid] <= d - mid - 1) lo = mid + 1; else hi = mid;
  }
  row = lo; nz = d - lo;
}

// row coordinate of the merge path at the start of every tile (and at its end)
__global__ void __launch_bounds__(kBlock)
merge_coords_kernel(const int* __restrict__ rowptr, int64_t n_rows, int64_t nnz, int64_t n_tiles, int* __restrict__ tile_row) {
  const int64_t total = n_rows + nnz;
  for (int64_t t = blockIdx.x * (int64_t)kBlock + threadIdx.x; t <= n_tiles; t += (int64_t)gridDim.x * kBlock) {
    int64_t r, k;
    merge_search(min(t * kMergeTile, total), rowptr + 1, n_rows, nnz, r, k);
    tile_row[t] = (int)r;
  }
}

__global__ void __launch_bounds__(kBlock)
spmv_merge_kernel(const psb_csr A, const double* x, double* __restrict__ ysum, int* __restrict__ carry_row,
                  double* __restrict__ carry_val, const int* __restrict__ d_skip) {
  __shared__ int s_end[kMergeTile + 2];
  __shared__ double s_acc[kMergeTile + 2];
  __shared__ double s_prod[kMergeTile];                      // the tile's products, staged with coalesced loads
  __shared__ int s_key[kBlock];
  __shared__ double s_val[kBlock];
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;
  const int tid = threadIdx.x;
  const int* row_end = A.rowptr + 1;
  const int64_t total = A.n_rows + A.nnz;
  const int64_t n_tiles = (total + kMergeTile - 1) / kMergeTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t d0 = tile * kMergeTile, d1 = min(d0 + (int64_t)kMergeTile, total);
    // tile boundaries on the merge path depend on the structure only: searched once when the kind is
    // set (a binary search in global memory here cost ~20 dependent loads per tile, the whole CTA waiting)
    const int64_t r0 = A.merge_tile_row[tile], r1 = A.merge_tile_row[tile + 1];
    const int64_t k0 = d0 - r0;
    const int n_tk = (int)((d1 - r1) - k0);                   // entries of the tile
    const int n_tr = (int)(r1 - r0) + 1;                      // rows touched; the last one may be partial
#pragma unroll
    for (int u = 0; u <= kMergeItems; ++u) {                 // n_tr <= kMergeTile + 1: independent loads, unrolled
      const int i = tid + u * kBlock;
      if (i < n_tr) {
        s_end[i] = r0 + i < A.n_rows ? row_end[r0 + i] : INT32_MAX;
        s_acc[i] = 0.0;
      }
    }
    // a thread's segment is kMergeItems CONSECUTIVE entries: read straight from global memory that is
    // a 56-byte stride between lanes (measured 1 TB/s); staged here, consecutive lanes read
    // consecutive entries
    {
      // all of a thread's (up to kMergeItems) entries in flight: columns and values, then the gathers
      int cc[kMergeItems];
      double vv[kMergeItems], xx[kMergeItems];
#pragma unroll
      for (int u = 0; u < kMergeItems; ++u) {
        const int i = tid + u * kBlock;
        cc[u] = i < n_tk ? A.colind[k0 + i] : 0;
        vv[u] = i < n_tk ? A.vals[k0 + i] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kMergeItems; ++u) xx[u] = (tid + u * kBlock) < n_tk ? ld_ca(x + cc[u]) : 0.0;
#pragma unroll
      for (int u = 0; u < kMergeItems; ++u) {
        const int i = tid + u * kBlock;
        if (i < n_tk) s_prod[i] = vv[u] * xx[u];
      }
    }
    __syncthreads();
    // this thread's segment
    const int64_t d = min(d0 + (int64_t)tid * kMergeItems, d1);
    const int64_t de = min(d + (int64_t)kMergeItems, d1);
    int64_t lo = max(d - A.nnz, r0), hi = min(d, r1);       // search restricted to the tile's rows
    lo = max(lo, (int64_t)0);
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)s_end[mid - r0] <= d - mid - 1) lo = mid + 1; else hi = mid;
    }
    int64_t r = lo, k = d - lo;
    double acc = 0.0;
    for (int64_t it = d; it < de; ++it) {
      if (k < (int64_t)s_end[r - r0]) {
        acc += s_prod[k - k0];                                 // stored order inside the segment
        ++k;
      } else {
        s_acc[r - r0] = acc;                                   // the row ends here: its tail part
        acc = 0.0;
        ++r;
      }
    }
    s_key[tid] = (int)(r - r0);
    s_val[tid] = acc;                                          // carry: head part of row r
    __syncthreads();
    // carries of consecutive threads with the same row are one run: segmented scan per 32, runs
    // added into the row's slot in thread order (warp 0, eight rounds: deterministic)
    if (tid < 32) {
      for (int c = 0; c < kBlock; c += 32) {
        const int key = s_key[c + tid];
        double v = s_val[c + tid];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double up = __shfl_up_sync(0xffffffffu, v, o);
          const int kup = __shfl_up_sync(0xffffffffu, key, o);
          if (tid >= o && kup == key) v = up + v;
        }
        const int knext = __shfl_down_sync(0xffffffffu, key, 1);
        if (tid == 31 || knext != key) s_acc[key] = s_acc[key] + v;   // run tail; later rounds add behind it
        __syncwarp();
      }
    }
    __syncthreads();
    for (int i = tid; i < n_tr - 1; i += kBlock) ysum[r0 + i] = s_acc[i];   // rows that END in this tile
    if (tid == 0) {
      const bool open = r1 < A.n_rows;                         // the last row continues in the next tile
      carry_row[tile] = open ? (int)r1 : -1;
      carry_val[tile] = open ? s_acc[n_tr - 1] : 0.0;
    }
    __syncthreads();
  }
}

// adds the head parts that earlier CTAs computed to the row's tail part, in CTA order
__global__ void __launch_bounds__(kBlock)
spmv_merge_fixup_kernel(int64_t n_tiles, const int* __restrict__ carry_row, const double* __restrict__ carry_val,
                        double* __restrict__ ysum, const int* __restrict__ d_skip) {
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;
  for (int64_t t = blockIdx.x * (int64_t)kBlock + threadIdx.x; t < n_tiles; t += (int64_t)gridDim.x * kBlock) {
    const int row = carry_row[t];
    if (row < 0 || (t > 0 && carry_row[t - 1] == row)) continue;   // not the first CTA of this row's run
    double sum = 0.0;
    for (int64_t u = t; u < n_tiles && carry_row[u] == row; ++u) sum += carry_val[u];
    ysum[row] = sum + ysum[row];
  }
}

// the epilogue of a merge-path product: one thread per row on the finished sums
template <int EPI>
__global__ void __launch_bounds__(kBlock)
spmv_merge_epi_kernel(const psb_csr A, const double* __restrict__ ysum, const double* x, double* y,
                      const EpiArgs ea, const int* __restrict__ d_skip) {
  __shared__ double scratch[kWarps];
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < A.n_rows; i += (int64_t)gridDim.x * kBlock)
    epilogue<EPI>(A.row_off + i, ysum[i], x, y, ea, acc);
  finish_dot<EPI>(acc, A, ea, scratch);
}

template <int EPI>
static int launch_merge(const psb_csr* A, const double* x, double* y, const EpiArgs& ea, const int* d_skip,
                        cudaStream_t st) {
  const int64_t total = A->n_rows + A->nnz;
  const int64_t n_tiles = (total + kMergeTile - 1) / kMergeTile;
  if (A->merge_ysum == nullptr || n_tiles > A->merge_tiles) {
    set_error("spmv_launch: the matrix has no merge-path workspace (psb_csr_set_kind(PSB_SPMV_MERGE) allocates it)");
    return PSB_ERR_ARG;
  }
  static thread_local int per_sm = 0;
  if (per_sm == 0) {
    int rc = occupancy_per_sm(spmv_merge_kernel, 0, &per_sm);
    if (rc != PSB_OK) return rc;
  }
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)per_sm * sm_count(), n_tiles));
  spmv_merge_kernel<<<grid, kBlock, 0, st>>>(*A, x, A->merge_ysum, A->merge_carry_row, A->merge_carry_val, d_skip);
  PSB_LAUNCH_CHECK();
  const int g2 = (int)std::max<int64_t>(1, std::min<int64_t>((n_tiles + kBlock - 1) / kBlock, (int64_t)sm_count() * 4));
  spmv_merge_fixup_kernel<<<g2, kBlock, 0, st>>>(n_tiles, A->merge_carry_row, A->merge_carry_val, A->merge_ysum, d_skip);
  PSB_LAUNCH_CHECK();
  const int g3 = grid_for(8, (A->n_rows + kBlock - 1) / kBlock, A->max_grid);
  spmv_merge_epi_kernel<EPI><<<g3, kBlock, 0, st>>>(*A, A->merge_ysum, x, y, ea, d_skip);
  PSB_LAUNCH_CHECK();
  return PSB_OK;
}

template <int EPI>
static int launch_epi(const psb_csr* A, const double* x, double* y, const EpiArgs& ea,
                      const int* d_skip, cudaStream_t st) {
  if (A->n_rows == 0) return PSB_OK;
  if (A->kind == PSB_SPMV_MERGE) return launch_merge<EPI>(A, x, y, ea, d_skip, st);
  if (A->kind == PSB_SPMV_STREAM) {
    return A->rpt == 2 ? launch_bulk<EPI, 2>(A, x, y, ea, d_skip, st)
                       : launch_bulk<EPI, 1>(A, x, y, ea, d_skip, st);
  }
  if (A->kind == PSB_SPMV_STREAM_LSU) {
    if (A->rpt == 2)
      return A->vec_loads ? launch_stream<EPI, 2, true>(A, x, y, ea, d_skip, st)
                          : launch_stream<EPI, 2, false>(A, x, y, ea, d_skip, st);
    return A->vec_loads ? launch_stream<EPI, 1, true>(A, x, y, ea, d_skip, st)
                        : launch_stream<EPI, 1, false>(A, x, y, ea, d_skip, st);
  }
  switch (A->vec_width) {
    case 2:  return launch_vector<EPI, 2>(A, x, y, ea, d_skip, st);
    case 4:  return launch_vector<EPI, 4>(A, x, y, ea, d_skip, st);
    case 8:  return launch_vector<EPI, 8>(A, x, y, ea, d_skip, st);
    case 16: return launch_vector<EPI, 16>(A, x, y, ea, d_skip, st);
    default: return launch_vector<EPI, 32>(A, x, y, ea, d_skip, st);
  }
}

psb_csr csr_row_view(const psb_csr* A, int64_t r0, int64_t r1) {
  psb_csr v = *A;
  v.rowptr = A->rowptr + r0;
  v.n_rows = r1 - r0;
  v.row_off = A->row_off + r0;
  // max_tile_nnz was measured on tiles aligned to 256 / 512 parent rows; the view's tiles start
  // at r0.  A misaligned tile straddles two aligned ones, so twice the parent's figure (or the
  // longest row times the tile height, if smaller) bounds it -- the stages are sized from this.
  for (int i = 0; i < 2; ++i) {
    const int64_t rows = (int64_t)kBlock * (i + 1);
    if (r0 % rows != 0) {
      const int64_t twice = 2 * (int64_t)A->max_tile_nnz[i];
      const int64_t by_row = rows * (int64_t)A->max_row;
      v.max_tile_nnz[i] = (int)std::min<int64_t>(twice, by_row);
    }
  }
  v.mega_grid = 0; v.mega_tile_rows = 0; v.mega_tile_nnz = 0;
  // the merge path (tile coordinates, carries) belongs to the whole matrix: a row range of it runs
  // on the sub-warp kernel
  if (v.kind == PSB_SPMV_MERGE && (r0 != 0 || r1 != A->n_rows)) v.kind = PSB_SPMV_VECTOR;
  return v;
}

int csr_max_tile_nnz(const psb_csr* A, int tile_rows, int* out, cudaStream_t st) {
  if (tile_rows < 1) { set_error("csr_max_tile_nnz: bad tile size"); return PSB_ERR_ARG; }
  *out = 0;
  if (A->n_rows == 0) return PSB_OK;
  int* d = nullptr;
  PSB_CUDA(cudaMalloc(&d, sizeof(int)));
  cudaError_t e = cudaMemsetAsync(d, 0, sizeof(int), st);
  if (e == cudaSuccess) {
    const int64_t tiles = (A->n_rows + tile_rows - 1) / tile_rows;
    const int grid = (int)std::min<int64_t>((tiles + kBlock - 1) / kBlock, (int64_t)sm_count() * 8);
    csr_tile_stats_kernel<<<std::max(grid, 1), kBlock, 0, st>>>(A->rowptr, A->n_rows, tile_rows, d);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    e = cudaPeekAtLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) { set_error("csr_max_tile_nnz: %s", cudaGetErrorString(e)); return PSB_ERR_CUDA; }
  return PSB_OK;
}

int spmv_launch(const psb_csr* A, Epi epi, const double* x, double* y, const EpiArgs& ea,
                const int* d_skip, cudaStream_t st) {
  if (epi == EPI_DOT_PUP) {
    if (A->kind != PSB_SPMV_STREAM) { set_error("spmv_launch: EPI_DOT_PUP needs the STREAM kernel"); return PSB_ERR_UNSUPP; }
    if (A->n_rows == 0) return PSB_OK;
    return A->rpt == 2 ? launch_bulk<EPI_DOT_PUP, 2>(A, x, y, ea, d_skip, st)
                       : launch_bulk<EPI_DOT_PUP, 1>(A, x, y, ea, d_skip, st);
  }
  switch (epi) {
    case EPI_STORE:  return launch_epi<EPI_STORE>(A, x, y, ea, d_skip, st);
    case EPI_DOT:    return launch_epi<EPI_DOT>(A, x, y, ea, d_skip, st);
    case EPI_RESID:  return launch_epi<EPI_RESID>(A, x, y, ea, d_skip, st);
    case EPI_ADD:    return launch_epi<EPI_ADD>(A, x, y, ea, d_skip, st);
    case EPI_JACOBI: return launch_epi<EPI_JACOBI>(A, x, y, ea, d_skip, st);
    case EPI_RESID_NORM: return launch_epi<EPI_RESID_NORM>(A, x, y, ea, d_skip, st);
    default: break;
  }
  set_error("spmv_launch: unknown epilogue %d", (int)epi);
  return PSB_ERR_ARG;
}

// shared memory budgets: the bulk-copy kernel keeps its bandwidth with two CTAs per SM (restriction
// operator of the Bratu-2048^2 hierarchy, 15 entries per row, 92 KB per CTA: 32.8 us against 58.9 us
// for the LSU kernel it got under the former 72 KB limit); the LSU kernel wants >= 3 CTAs per SM
static constexpr size_t kBulkSmemBudget = 110 * 1024;
static constexpr size_t kLsuSmemBudget = 48 * 1024;

static size_t lsu_smem(const psb_csr* A, int rpt) {
  return (size_t)((A->max_tile_nnz[rpt - 1] + 1) & ~1) * 8 + (size_t)(kBlock * rpt + 1) * 4;
}

// workspace of the merge-path kind: row sums + one carry per CTA tile
static int merge_alloc(psb_csr* A) {
  const int64_t tiles = (A->n_rows + A->nnz + kMergeTile - 1) / kMergeTile + 1;
  if (A->merge_ysum != nullptr && A->merge_tiles >= tiles) return PSB_OK;
  cudaFree(A->merge_ysum); cudaFree(A->merge_carry_row); cudaFree(A->merge_carry_val); cudaFree(A->merge_tile_row);
  A->merge_ysum = nullptr; A->merge_carry_row = nullptr; A->merge_carry_val = nullptr; A->merge_tile_row = nullptr;
  A->merge_tiles = 0;
  PSB_CUDA(cudaMalloc((void**)&A->merge_ysum, (size_t)std::max<int64_t>(A->n_rows, 1) * sizeof(double)));
  PSB_CUDA(cudaMalloc((void**)&A->merge_carry_row, (size_t)tiles * sizeof(int)));
  PSB_CUDA(cudaMalloc((void**)&A->merge_carry_val, (size_t)tiles * sizeof(double)));
  PSB_CUDA(cudaMalloc((void**)&A->merge_tile_row, (size_t)(tiles + 1) * sizeof(int)));
  // set-up time: the legacy default stream orders this after the uploads of the caller's streams
  const int64_t n_tiles = (A->n_rows + A->nnz + kMergeTile - 1) / kMergeTile;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_tiles + kBlock) / kBlock, (int64_t)sm_count() * 4));
  PSB_CUDA(cudaDeviceSynchronize());
  merge_coords_kernel<<<grid, kBlock>>>(A->rowptr, A->n_rows, A->nnz, n_tiles, A->merge_tile_row);
  PSB_LAUNCH_CHECK();
  PSB_CUDA(cudaDeviceSynchronize());
  A->merge_tiles = tiles;
  return PSB_OK;
}

static void choose_kernel(psb_csr* A) {
  const double mean = A->n_rows ? (double)A->nnz / (double)A->n_rows : 0.0;
  int w = 2;
  while (w < 32 && w < mean) w <<= 1;
  A->vec_width = w;
  A->kind = PSB_SPMV_VECTOR; A->rpt = 1;
  // skewed histogram: a few rows far longer than the rest (>= 512 entries and >= 16 x the mean) would
  // serialise a thread-per-row or sub-warp-per-row kernel -> split the work, not the rows.  Only when
  // the longest row is a visible share of the matrix (>= nnz / 512): the sub-warp kernel walks a row at
  // ~0.1 us per 16 entries, so shorter rows hide behind the streaming of the rest, and the merge
  // kernels run at ~70 % of its bandwidth (U12 block of the Bratu-2048^2 coarse LU, 691 200 x 8 192,
  // 10.9 M entries: VECTOR 52 us, MERGE 75 us)
  if (A->max_row >= 512 && (double)A->max_row >= 16.0 * std::max(mean, 2.0) &&
      (int64_t)A->max_row * 512 >= A->nnz && merge_alloc(A) == PSB_OK) {
    A->kind = PSB_SPMV_MERGE;
    return;
  }
  if (mean > 32.0) return;
  if (A->vec_loads) {                                  // bulk copies need 16-byte alignment
    int cv, cc; size_t smem;
    bulk_caps(A, 1, &cv, &cc, &smem);
    if (smem <= kBulkSmemBudget) { A->kind = PSB_SPMV_STREAM; A->rpt = 1; return; }
  }
  if (lsu_smem(A, 2) <= kLsuSmemBudget) { A->kind = PSB_SPMV_STREAM_LSU; A->rpt = 2; }
  else if (lsu_smem(A, 1) <= kLsuSmemBudget) { A->kind = PSB_SPMV_STREAM_LSU; A->rpt = 1; }
}

}  // namespace psb

using namespace psb;

extern "C" int psb_csr_create(int64_t n_rows, int64_t n_cols, int64_t nnz,
                              const int32_t* d_rowptr, const int32_t* d_colind,
                              const double* d_vals, void* stream, psb_csr_t* out) {
  PSB_REQUIRE(out != nullptr, PSB_ERR_ARG, "psb_csr_create: out is NULL");
  PSB_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0, PSB_ERR_ARG, "psb_csr_create: negative size");
  PSB_REQUIRE(nnz < (int64_t)INT32_MAX && n_rows < (int64_t)INT32_MAX && n_cols < (int64_t)INT32_MAX,
              PSB_ERR_UNSUPP, "psb_csr_create: int32 index range exceeded");
  PSB_REQUIRE(d_rowptr != nullptr, PSB_ERR_ARG, "psb_csr_create: rowptr is NULL");
  PSB_REQUIRE(nnz == 0 || (d_colind != nullptr && d_vals != nullptr), PSB_ERR_ARG,
              "psb_csr_create: colind/vals NULL");
  cudaStream_t st = (cudaStream_t)stream;
  psb_csr* A = new (std::nothrow) psb_csr();
  PSB_REQUIRE(A != nullptr, PSB_ERR_ARG, "psb_csr_create: out of host memory");
  A->n_rows = n_rows; A->n_cols = n_cols; A->nnz = nnz; A->row_off = 0;
  A->rowptr = d_rowptr; A->colind = d_colind; A->vals = d_vals;
  A->vec_loads = aligned16(d_colind) && aligned16(d_vals) && aligned16(d_rowptr);   // rowptr is bulk-copied too
  A->max_grid = sm_count() * 16;
  A->partials = nullptr; A->ticket = nullptr; A->colind16 = nullptr;
  A->mega_grid = 0; A->mega_tile_rows = 0; A->mega_tile_nnz = 0;
  A->merge_ysum = nullptr; A->merge_carry_row = nullptr; A->merge_carry_val = nullptr; A->merge_tiles = 0;
  A->merge_tile_row = nullptr;
  int h_stats[4] = {0, 0, 0, 0};        // longest row, fullest 256- / 512-row tile, "a column delta does not fit 16 bits"
  int* d_stats = nullptr;
  cudaError_t e = cudaMalloc(&A->partials, sizeof(double) * A->max_grid);
  if (e == cudaSuccess) e = cudaMalloc(&A->ticket, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemsetAsync(A->ticket, 0, sizeof(unsigned int), st);
  if (e == cudaSuccess) e = cudaMalloc(&d_stats, sizeof(h_stats));
  if (e == cudaSuccess) e = cudaMemsetAsync(d_stats, 0, sizeof(h_stats), st);
  if (e == cudaSuccess && n_rows > 0) {
    int grid = (int)std::min<int64_t>((n_rows + kBlock - 1) / kBlock, (int64_t)sm_count() * 8);
    csr_stats_kernel<<<grid, kBlock, 0, st>>>(d_rowptr, n_rows, d_stats);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    e = cudaPeekAtLastError();
    // OPTIONAL (psb_csr_set_cols16 / PSB_SPMV_C16=1, off by default).  Banded matrices
    // (|col - row| < 2^15, e.g. every 2-D stencil up to 32 767 columns per grid line): a second
    // copy of the column indices as 16-bit distances from the diagonal lets the STREAM kernels
    // move 10 instead of 12 bytes per entry.  Built speculatively in the same stream, judged with
    // the statistics (one synchronisation), dropped if a distance does not fit or another kernel
    // is chosen.  Measured at C3 (profiles/round1f_notes.md): 12.5 % fewer bytes buy 4 % on the
    // stand-alone SpMV and nothing in the persistent PCG kernel, and the extra pass costs 0.8 %
    // end to end -- at 93 - 97 % of the copy bandwidth the kernels are no longer purely DRAM-bound.
    if (g_cols16 < 0) { const char* v = getenv("PSB_SPMV_C16"); g_cols16 = (v && atoi(v) != 0) ? 1 : 0; }
    if (e == cudaSuccess && g_cols16 > 0 && n_rows == n_cols && nnz >= (1 << 16) && A->vec_loads &&
        nnz <= 32 * n_rows) {
      short* d16 = nullptr;
      if (cudaMalloc(&d16, sizeof(short) * (size_t)nnz) == cudaSuccess) {
        csr_delta16_kernel<<<grid, kBlock, 0, st>>>(d_rowptr, d_colind, n_rows, d16, d_stats + 3);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        e = cudaPeekAtLastError();
        A->colind16 = d16;
      } else {
        (void)cudaGetLastError();      // no memory for the copy: carry on with 32-bit indices
      }
    }
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (d_stats) cudaFree(d_stats);
  if (e != cudaSuccess) {
    set_error("psb_csr_create: %s", cudaGetErrorString(e));
    if (A->partials) cudaFree(A->partials);
    if (A->ticket) cudaFree(A->ticket);
    if (A->colind16) cudaFree((void*)A->colind16);
    delete A;
    return PSB_ERR_CUDA;
  }
  A->max_row = h_stats[0];
  A->max_tile_nnz[0] = h_stats[1];
  A->max_tile_nnz[1] = h_stats[2];
  choose_kernel(A);
  if (A->colind16 != nullptr && (h_stats[3] != 0 || A->kind != PSB_SPMV_STREAM)) {
    cudaFree((void*)A->colind16);
    A->colind16 = nullptr;
  }
  *out = A;
  return PSB_OK;
}

extern "C" int psb_csr_set_cols16(int enable) {
  g_cols16 = enable ? 1 : 0;
  return PSB_OK;
}

extern "C" int psb_csr_destroy(psb_csr_t A) {
  if (A == nullptr) return PSB_OK;
  if (A->partials) cudaFree(A->partials);
  if (A->ticket) cudaFree(A->ticket);
  if (A->colind16) cudaFree((void*)A->colind16);
  cudaFree(A->merge_ysum); cudaFree(A->merge_carry_row); cudaFree(A->merge_carry_val); cudaFree(A->merge_tile_row);
  delete A;
  return PSB_OK;
}

extern "C" int psb_csr_info(psb_csr_t A, int64_t info[8]) {
  PSB_REQUIRE(A != nullptr && info != nullptr, PSB_ERR_ARG, "psb_csr_info: NULL argument");
  info[0] = A->kind;
  info[1] = A->max_row;
  info[2] = A->max_tile_nnz[0];
  info[3] = (int64_t)kBlock * A->rpt;
  info[4] = A->vec_width;
  info[5] = A->max_grid;
  info[6] = (A->vec_loads ? 1 : 0) | (A->colind16 ? 2 : 0);
  info[7] = A->max_tile_nnz[1];
  return PSB_OK;
}

extern "C" int psb_csr_set_kind(psb_csr_t A, int kind) {
  PSB_REQUIRE(A != nullptr, PSB_ERR_ARG, "psb_csr_set_kind: NULL handle");
  const int base = kind & 15;
  const int rpt = (kind & PSB_SPMV_TILE512) ? 2 : 1;
  if (base == PSB_SPMV_VECTOR) { A->kind = base; return PSB_OK; }
  if (base == PSB_SPMV_MERGE) {
    int rc = merge_alloc(A);
    if (rc != PSB_OK) return rc;
    A->kind = base;
    return PSB_OK;
  }
  if (base == PSB_SPMV_STREAM) {
    PSB_REQUIRE(A->vec_loads, PSB_ERR_UNSUPP, "psb_csr_set_kind: arrays are not 16-byte aligned");
    int cv, cc; size_t smem;
    bulk_caps(A, rpt, &cv, &cc, &smem);
    PSB_REQUIRE(smem <= (size_t)max_optin_smem(), PSB_ERR_UNSUPP,
                "psb_csr_set_kind: tile does not fit in shared memory");
    A->kind = base; A->rpt = rpt;
    return PSB_OK;
  }
  if (base == PSB_SPMV_STREAM_LSU) {
    PSB_REQUIRE(lsu_smem(A, rpt) <= (size_t)max_optin_smem(), PSB_ERR_UNSUPP,
                "psb_csr_set_kind: tile does not fit in shared memory");
    A->kind = base; A->rpt = rpt;
    return PSB_OK;
  }
  set_error("psb_csr_set_kind: unknown kind %d", kind);
  return PSB_ERR_ARG;
}

extern "C" int psb_spmv(psb_csr_t A, const double* d_x, double* d_y, void* stream) {
  PSB_REQUIRE(A && d_x && d_y, PSB_ERR_ARG, "psb_spmv: NULL argument");
  return spmv_launch(A, EPI_STORE, d_x, d_y, EpiArgs(), nullptr, (cudaStream_t)stream);
}

extern "C" int psb_spmv_dot(psb_csr_t A, const double* d_x, double* d_y, double* d_dot, void* stream) {
  PSB_REQUIRE(A && d_x && d_y && d_dot, PSB_ERR_ARG, "psb_spmv_dot: NULL argument");
  PSB_REQUIRE(A->n_rows == A->n_cols, PSB_ERR_ARG, "psb_spmv_dot: matrix must be square");
  EpiArgs ea; ea.dot = d_dot;
  return spmv_launch(A, EPI_DOT, d_x, d_y, ea, nullptr, (cudaStream_t)stream);
}

extern "C" int psb_spmv_residual(psb_csr_t A, const double* d_x, const double* d_f, double* d_y,
                                 void* stream) {
  PSB_REQUIRE(A && d_x && d_y && d_f, PSB_ERR_ARG, "psb_spmv_residual: NULL argument");
  EpiArgs ea; ea.f = d_f;
  return spmv_launch(A, EPI_RESID, d_x, d_y, ea, nullptr, (cudaStream_t)stream);
}

extern "C" int psb_spmv_add(psb_csr_t A, const double* d_x, double* d_y, void* stream) {
  PSB_REQUIRE(A && d_x && d_y, PSB_ERR_ARG, "psb_spmv_add: NULL argument");
  return spmv_launch(A, EPI_ADD, d_x, d_y, EpiArgs(), nullptr, (cudaStream_t)stream);
}

extern "C" int psb_jacobi_sweep(psb_csr_t A, const double* d_dinv, double omega, const double* d_f,
                                const double* d_x, double* d_xnew, void* stream) {
  PSB_REQUIRE(A && d_dinv && d_f && d_x && d_xnew, PSB_ERR_ARG, "psb_jacobi_sweep: NULL argument");
  PSB_REQUIRE(d_x != d_xnew, PSB_ERR_ARG, "psb_jacobi_sweep: x_new must not alias x");
  PSB_REQUIRE(A->n_rows == A->n_cols, PSB_ERR_ARG, "psb_jacobi_sweep: matrix must be square");
  EpiArgs ea; ea.f = d_f; ea.dinv = d_dinv; ea.omega = omega;
  return spmv_launch(A, EPI_JACOBI, d_x, d_xnew, ea, nullptr, (cudaStream_t)stream);
}
