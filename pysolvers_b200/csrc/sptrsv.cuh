// Level-scheduled, synchronisation-free sparse triangular solve (sptrsv.cu).
#pragma once
#include "common.cuh"

#include <vector>

struct psb_trsv {
  int64_t n = 0, nnz_off = 0, nnz_packed = 0;
  int lower = 1, unit_diag = 0;
  int n_levels = 0;
  int n_groups = 0;                   // work chunks (32 short rows, or one long row per warp)
  int n_long = 0;                     // rows handled by a whole warp
  int n_subwarp = 0;                  // chunks of up to 4 rows on 8 lanes each (wide levels: grid kernel only)
  int cta_ok = 1;                     // the packing can be run by the one-CTA kernel
  // host copies kept for inspection / bit-exact tests
  std::vector<int32_t> h_level_ptr;   // n_levels + 1
  std::vector<int32_t> h_level_rows;  // n, level-major, ascending inside a level
  // device arrays (owned).  Processing order = level-major, inside a level the short rows
  // first (ascending), then the long ones; item q is the q-th row in that order.
  int32_t* d_order = nullptr;     // [n]           row id of item q
  double*  d_diag = nullptr;      // [n]           1 / diagonal of item q (1 for unit_diag)
  int64_t* d_grp_ptr = nullptr;   // [n_groups+1]  start of chunk g in cols/vals
  int32_t* d_grp_item = nullptr;  // [n_groups]    first item of chunk g
  int32_t* d_grp_rows = nullptr;  // [n_groups]    rows in the chunk (1..32), 0 = one long row, -r = r rows of 8 lanes each
  int32_t* d_row_cnt = nullptr;   // [n]           off-diagonal entries of item q
  int32_t* d_cols = nullptr;      // [nnz_packed]  short rows: lane l of chunk g holds its row RIGHT-aligned, entry k of
                                  //               cnt at grp_ptr[g] + 32 (width - cnt + k) + l; -1 = padding (in front)
  double*  d_vals = nullptr;      // [nnz_packed]
  unsigned int* d_counter = nullptr;   // next unclaimed chunk
  int* d_error = nullptr;              // set when a spin timed out
  // shared-memory window kernel (sptrsv_cta.cu): dependencies as POSITIONS in the processing order
  int32_t* d_wcols = nullptr;     // [nnz_packed]  >= 0 byte offset in the window (padding: the zero slot); <= -2 far (row = -c - 2)
  int32_t* d_wmeta = nullptr;     // [4 n_groups]  {base lo, base hi, first item, rows | has_far << 6 | entries-per-lane << 7}
  int wslots = 0;                 // window size in doubles (power of two)
  int stage_len = 0;              // entries per lane of a warp's staging buffer
  int64_t n_far = 0;              // dependencies older than the window (read from the global vector)
  int64_t max_dist = 0;           // largest position distance of a dependency
  int cluster_ok = 0;             // the window data respect the (larger) reuse slack of the cluster kernel
  int kernel = 0;                 // PSB_TRSV_GRID / PSB_TRSV_CTA / PSB_TRSV_CLUSTER chosen by the analysis
  long long* d_trace = nullptr;   // not owned: psb_trsv_set_trace (debugging: per-chunk time stamps)
  int forced_kernel = -1;         // psb_trsv_set_kernel: >= 0 overrides the analysis
};

enum { PSB_TRSV_GRID = 0, PSB_TRSV_CTA = 1, PSB_TRSV_CLUSTER = 2 };

namespace psb {

constexpr int kTrsvCtaWarps = 16;        // warps of the one-CTA kernel (512 threads: 128 registers each)
constexpr int kTrsvAhead = 2 * kTrsvCtaWarps;   // chunks a warp may run ahead of the slowest one (2 rounds)
constexpr int kTrsvMaxSlots = 16384;     // largest window of the shared-memory kernel: 128 KB of fp64
constexpr int kTrsvSmemBudget = 224 * 1024;   // window + 16 staging buffers must fit (227 KB opt-in limit)
constexpr int kTrsvClusterSize = 4;      // CTAs of the cluster kernel: 64 warps on 4 SMs (8 CTAs would need a reuse
                                         // slack of 32 (2 * 256 + 2) positions, more than the 16 384-slot window)
constexpr double kTrsvCtaMaxChunksPerLevel = 3.5;       // mean chunks per level up to which ONE CTA is used
constexpr double kTrsvClusterMaxChunksPerLevel = 12.0;  // ... up to which the 4-CTA cluster is used; beyond: the grid

// x = T^-1 rhs, enqueued on st.  rhs_map (nullable): row r takes rhs[rhs_map[r]].
// out2/out_map (nullable): additionally out2[out_map[r]] = x[r].
// x is overwritten with a sentinel first and must not alias rhs.
int trsv_solve(const psb_trsv* T, const double* rhs, double* x, const int32_t* rhs_map,
               double* out2, const int32_t* out_map, const int* d_skip, cudaStream_t st);

// the same solve by ONE CTA, or one cluster of CTAs, with the wavefront in shared memory (sptrsv_cta.cu)
int trsv_solve_cta(const psb_trsv* T, int cluster, const double* rhs, double* x, const int32_t* rhs_map,
                   double* out2, const int32_t* out_map, const int* d_skip, cudaStream_t st);

// Reads a factor's device error flag (synchronising) and clears it once it has been reported,
// so that one timed-out solve does not condemn every later use of the same factor.
inline int trsv_take_error(const psb_trsv* T) {
  int v = 0;
  if (T == nullptr || T->d_error == nullptr) return 0;
  if (cudaMemcpy(&v, T->d_error, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
  if (v != 0) cudaMemset(T->d_error, 0, sizeof(int));
  return v;
}

}  // namespace psb
