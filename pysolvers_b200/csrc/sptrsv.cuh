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
  // host copies kept for inspection / bit-exact tests
  std::vector<int32_t> h_level_ptr;   // n_levels + 1
  std::vector<int32_t> h_level_rows;  // n, level-major, ascending inside a level
  // device arrays (owned).  Processing order = level-major, inside a level the short rows
  // first (ascending), then the long ones; item q is the q-th row in that order.
  int32_t* d_order = nullptr;     // [n]           row id of item q
  double*  d_diag = nullptr;      // [n]           diagonal of item q (1 for unit_diag)
  int64_t* d_grp_ptr = nullptr;   // [n_groups+1]  start of chunk g in cols/vals
  int32_t* d_grp_item = nullptr;  // [n_groups]    first item of chunk g
  int32_t* d_grp_rows = nullptr;  // [n_groups]    rows in the chunk (1..32), or 0 = one long row
  int32_t* d_cols = nullptr;      // [nnz_packed]  entry k of lane l at grp_ptr[g] + 32k + l; -1 = padding
  double*  d_vals = nullptr;      // [nnz_packed]
  unsigned int* d_counter = nullptr;   // next unclaimed chunk
  int* d_error = nullptr;              // set when a spin timed out
};

namespace psb {

// x = T^-1 rhs, enqueued on st.  rhs_map (nullable): row r takes rhs[rhs_map[r]].
// out2/out_map (nullable): additionally out2[out_map[r]] = x[r].
// x is overwritten with a sentinel first and must not alias rhs.
int trsv_solve(const psb_trsv* T, const double* rhs, double* x, const int32_t* rhs_map,
               double* out2, const int32_t* out_map, const int* d_skip, cudaStream_t st);

}  // namespace psb
