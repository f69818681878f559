// Persistent single-launch PCG (pcg_mega.cu): parameters shared with pcg.cu and dist.cu.
#pragma once
#include "spmv.cuh"

namespace psb {

constexpr int kRing = 4;         // reduction slots are reused every kRing epochs
constexpr int kMaxRanks = 32;
constexpr int kMaxPush = 4;
// One all-reduce slot (two 8-byte epoch-tagged words) per 128-byte line: every CTA of a rank polls
// the slots of all ranks, and with the slots of 8 ranks in ONE line ~4 700 lanes hammered a single L2
// slice -- the peers' incoming stores queued behind them (reduce latency 4.5 us at 2 GPUs, 9 - 14 us at
// 8, profiles/round2_mega_timeline.md).  Spread over lines, i.e. L2 slices, each sees one lane per CTA.
constexpr int kSlotWords = 16;

struct MegaState {
  double norm_b, norm_r;
  int status, k_final, n_hist, done;
  unsigned int epochs_used;      // reduction epochs consumed by the launch
  unsigned int halo_epochs_used;
};

struct MegaParams {
  psb_csr A;                     // local rows; STREAM kind, 256-row tiles
  long long n;                   // local length of the vectors
  const double* b;
  double* x;
  double* r;                     // n (+ halo when row-partitioned)
  double* Ap;
  double* pbuf[2];               // p ping-pong, n (+ halo)
  long long n_halo;
  double* hist;
  MegaState* st;
  double* partials;
  unsigned int* ticket;
  // all-reduce slots (epoch-tagged, see common.cuh peer_push / peer_wait)
  const unsigned long long* my_slots;            // local ring: slot(e, q) at (e % kRing) * ring_words + q * kSlotWords
  unsigned long long* const* slot_ptrs;          // device [kRing * nranks]: my slot in rank q's memory
  int ring_words;                                // words between ring entries (kMaxRanks * kSlotWords; kSlotWords on one GPU)
  int nranks;
  unsigned int epoch0;
  // halo pushes to the neighbours (row-partitioned runs)
  int n_push;
  long long push_off[kMaxPush], push_cnt[kMaxPush];
  double* push_r[kMaxPush];                      // neighbour's r halo slice
  double* push_p[2][kMaxPush];                   // neighbour's p halo slice, per ping-pong buffer
  unsigned long long* push_flag[kMaxPush];
  int n_wait;
  const unsigned long long* my_flags;
  unsigned long long halo_epoch0;
  long long int_r0, int_r1;                      // rows [int_r0, int_r1) touch no halo column (set by the caller)
  int maxiter;
  double tau;
  int fail_on_maxiter;
  int* error;
  // ---- filled by pcg_mega_launch ----
  int tile_rows;                                 // rows per SpMV tile, balanced over the grid
  long long rot_t0, rot_t1;                      // interior tiles [rot_t0, rot_t1) run first
  int cap_v, cap_c;
  unsigned long long* timeline;                  // profiling (psb_debug_mega_timeline), nullable
  int tl_first, tl_count;
  int nopush_lo, nopush_hi;                      // rows in [nopush_lo, nopush_hi) lie in no push range (cheap per-row test)
  int dbg_flags;                                 // PSB_MEGA_FLAGS: 1 no early tile staging, 2 no phase-B loads across the barrier
};

// Whole PCG solve (identity preconditioner) in ONE cooperative launch.  Fills *P.st (device).
// `A` is the handle P.A was copied from: the tile plan (rows per tile, fullest tile) is cached on it.
int pcg_mega_launch(MegaParams& P, psb_csr* A, cudaStream_t stream);

// most nonzeros in any tile of `tile_rows` consecutive rows starting at row 0 (spmv.cu; synchronises)
int csr_max_tile_nnz(const psb_csr* A, int tile_rows, int* out, cudaStream_t stream);

}  // namespace psb
