// Preconditioner objects living behind psb_prec_t.
#pragma once
#include "common.cuh"

// z = M^-1 r, enqueued on `st`.  When d_skip is non-null the kernels are no-ops
// if *d_skip != 0 (a solver loop that has already converged keeps launching).
struct psb_prec {
  int64_t n = 0;
  virtual ~psb_prec() {}
  virtual int apply(const double* d_r, double* d_z, const int* d_skip, cudaStream_t st) = 0;
  virtual const char* kind() const = 0;
  // non-zero when a device-side failure was recorded (synchronises)
  virtual int check_error() { return 0; }
};

namespace psb {

// Reduction scratch that several kernels of one solve share: `partials` holds
// slots * max_grid doubles, `ticket` one counter per slot.
struct ReduceBuf {
  double* partials = nullptr;
  unsigned int* ticket = nullptr;
  int max_grid = 0;
};

// grid for the n-long streaming kernels: resident CTAs of the device, capped by work
int stream_grid(int64_t n, int max_grid);

}  // namespace psb
