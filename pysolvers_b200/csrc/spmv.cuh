// Internal interface of the CSR SpMV family (spmv.cu) used by the solvers.
#pragma once
#include "common.cuh"

struct psb_csr {
  int64_t n_rows, n_cols, nnz;
  int64_t row_off;      // row-range views: vectors are indexed at row_off + row (rowptr is pre-offset)
  const int*    rowptr;
  const int*    colind;
  const short*  colind16;   // owned, nullable: colind[k] - (row_off + row) as int16 when every |delta| fits
                            // (banded matrices): the STREAM kernels then stream 10 instead of 12 bytes per entry
  const double* vals;
  int  kind;            // PSB_SPMV_STREAM | PSB_SPMV_VECTOR
  int  max_row;         // longest row
  int  max_tile_nnz[2]; // most nnz in a tile of 256 / 512 consecutive rows
  int  rpt;             // STREAM: rows per thread (1 -> 256-row tiles, 2 -> 512)
  int  vec_width;       // VECTOR: lanes per row
  bool vec_loads;       // arrays 16-byte aligned -> 128-bit loads of vals/colind
  double*       partials;   // per-CTA partial sums of the fused dot
  unsigned int* ticket;     // last-CTA ticket
  int  max_grid;        // CTAs the partial buffer can hold
  // merge-path kind: per-row sums before the epilogue and the per-CTA carries (owned, lazily allocated)
  double* merge_ysum;
  int*    merge_carry_row;
  double* merge_carry_val;
  int64_t merge_tiles;
  int*    merge_tile_row;      // [tiles + 1] row coordinate of every tile boundary on the merge path (structure only)
  // persistent PCG kernel: tile plan cached per (grid) -- rows per tile and the fullest such tile
  int  mega_grid, mega_tile_rows, mega_tile_nnz;
};

namespace psb {

enum Epi : int {
  EPI_STORE = 0,   // y = A x
  EPI_DOT   = 1,   // y = A x ; dot = x . y
  EPI_RESID = 2,   // y = f - A x
  EPI_ADD   = 3,   // y += A x
  EPI_JACOBI = 4,  // y = x + omega * dinv .* (f - A x)
  EPI_DOT_PUP = 5, // PCG: p = x + beta*pold formed on the fly (gathers and own row), pnew = p,
                   // y = A p, dot = p . y   (STREAM kernel only)
  EPI_RESID_NORM = 6,  // y = f - A x ; sum of y^2 handed to the finish hook below (AMG cycle end)
  EPI_COUNT
};

// State of a V-cycle solve (amg.cu); lives here because the residual SpMV that ends a cycle
// finishes it in its own last CTA (EPI_RESID_NORM) instead of a separate pass over r.
struct AmgState {
  double norm_b, norm_r, tau;
  double bb, rr;
  int skip;        // != 0: every kernel of the remaining cycles is a no-op
  int cycles;      // cycles completed
  int status;
  int maxiter;
};

struct EpiArgs {
  const double* f     = nullptr;   // RESID, JACOBI
  const double* dinv  = nullptr;   // JACOBI
  double        omega = 1.0;       // JACOBI
  double*       dot   = nullptr;   // DOT: where the last CTA writes x . y
  int dot_accumulate  = 0;         // DOT: add to *dot (row-range launches chained on one stream)
  // DOT_PUP: beta = *beta_num / *beta_den, p = x + beta * pold, stored to pnew
  const double* pold = nullptr;
  double*       pnew = nullptr;
  const double* beta_num = nullptr;
  const double* beta_den = nullptr;
  // RESID_NORM: the last CTA ends the V-cycle -- ||r||, history, strict '<' test (VCycleSolver.py:87-91)
  AmgState*     amg_state = nullptr;
  double*       amg_hist = nullptr;
  // RESID_NORM: when not null, the first Jacobi sweep of the NEXT cycle is formed from the residual
  // just computed, jac_out = x + omega * dinv .* r (the arithmetic of EPI_JACOBI, which would
  // recompute the same r = f - A x with another pass over the matrix)
  double*       jac_out = nullptr;
  // DOT_PUP, persistent PCG: deferred solution update xsol += alpha_prev * pold on the tile's own rows
  double*       xsol = nullptr;
  double        alpha_prev = 0.0;
  // DOT_PUP, multi-GPU: rows in [pp_off, pp_off+pp_cnt) of the new p are also stored into the
  // neighbour's halo at pp_remote
  int pp_n = 0;
  long long pp_off[4] = {0, 0, 0, 0}, pp_cnt[4] = {0, 0, 0, 0};
  double* pp_remote[4] = {nullptr, nullptr, nullptr, nullptr};
  int pp_skip_lo = 0, pp_skip_hi = 0;   // rows in [pp_skip_lo, pp_skip_hi) are in no push range: two compares per row
  // ---- multi-GPU over NVLink peer memory (dist.cu) ------------------------------------
  // once the dot is final, store it (epoch-tagged, see peer_push) into this rank's slot in
  // every rank's memory
  unsigned long long* const* push_slots = nullptr;   // device array [push_n] of peer addresses
  int push_n = 0;
  unsigned int push_epoch = 0;
  // before touching x: wait until every wait_flags[i] >= wait_value (halo pushed by peers)
  const unsigned long long* wait_flags = nullptr;
  int wait_n = 0;
  unsigned long long wait_value = 0;
  int* error_flag = nullptr;       // set when a wait times out
  // tile order of the STREAM kernel: tiles [rot_t0, rot_t1) (rows without halo columns) are
  // processed first, the wait above happens only when a CTA reaches the first other tile
  long long rot_t0 = 0, rot_t1 = 0;
};

// Enqueue one SpMV-shaped kernel.  `d_skip` (nullable): the kernel is a no-op
// when *d_skip != 0 (solver loops keep launching after convergence).
int spmv_launch(const psb_csr* A, Epi epi, const double* x, double* y,
                const EpiArgs& ea, const int* d_skip, cudaStream_t stream);

}  // namespace psb

// ---- row-partitioned operators (dist.cu); opaque here -----------------------------------------
struct psb_dist;
struct psb_comm;
namespace psb {
int dist_spmv_epi(psb_dist* D, Epi epi, double* d_x_ext, double* d_y, const EpiArgs& ea, const int* d_skip,
                  cudaStream_t st);
int dist_allreduce(psb_comm* c, double* d_buf, int count, cudaStream_t st);
int dist_allgather_slices(psb_comm* c, double* d_full, const int64_t* starts, cudaStream_t st);
int dist_rank(const psb_comm* c);
int dist_nranks(const psb_comm* c);
int64_t dist_n_loc(const psb_dist* D);
int64_t dist_n_own(const psb_dist* D);
int64_t dist_n_halo(const psb_dist* D);
psb_comm* dist_comm(const psb_dist* D);

// View of rows [r0, r1) of A (r0 a multiple of 4 keeps the bulk copies aligned).  Shares
// the parent's arrays and reduction scratch; y, f, dinv and x[row] stay indexed by the
// parent's row numbers.
psb_csr csr_row_view(const psb_csr* A, int64_t r0, int64_t r1);

}  // namespace psb
