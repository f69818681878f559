// Persistent single-launch PCG (pcg_mega.cu): parameters shared with pcg.cu and dist.cu.
#pragma once
#include "spmv.cuh"

namespace psb {

constexpr int kRing = 4;         // reduction slots are reused every kRing epochs
constexpr int kMaxRanks = 32;
constexpr int kMaxPush = 4;

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
  const unsigned long long* my_slots;            // local ring: slot(e, q) at ((e % kRing) * kMaxRanks + q) * 2
  unsigned long long* const* slot_ptrs;          // device [kRing * nranks]: my slot in rank q's memory
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
  long long rot_t0, rot_t1;                      // interior tiles first
  int maxiter;
  double tau;
  int fail_on_maxiter;
  int cap_v, cap_c;
  int* error;
};

// Whole PCG solve (identity preconditioner) in ONE cooperative launch.  Fills *st (device).
int pcg_mega_launch(const MegaParams& P, cudaStream_t stream);
// shared-memory bytes / staging capacities of the STREAM pipeline for A (256-row tiles)
void pcg_mega_caps(const psb_csr* A, int* cap_v, int* cap_c, size_t* smem);

}  // namespace psb
