// Device-side residual and Jacobian of the Bratu problem (configuration 5's nonlinear function),
// so that the Newton outer loop can keep u, F and J in HBM.
//
//   evalF:  F = A u - alpha exp(-u)                  examples/FDBratu2D.py:20-21
//   evalJ:  J = A with diag += alpha exp(-u)          examples/FDBratu2D.py:23-29
// The structure of J never changes: only the diagonal VALUES of the device CSR are rewritten.
#include "common.cuh"
#include "prec.cuh"

namespace psb {

__global__ void __launch_bounds__(kBlock)
bratu_residual_kernel(int64_t n, const double* __restrict__ Au, const double* __restrict__ u,
                      double alpha, double* __restrict__ F) {
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    F[i] = Au[i] - alpha * exp(-u[i]);
}

__global__ void __launch_bounds__(kBlock)
bratu_jacobian_kernel(int64_t n, const int64_t* __restrict__ diag_pos, const double* __restrict__ a_diag,
                      const double* __restrict__ u, double alpha, double* __restrict__ vals) {
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    vals[diag_pos[i]] = a_diag[i] + alpha * exp(-u[i]);
}

}  // namespace psb

using namespace psb;

extern "C" int psb_bratu_residual(int64_t n, const double* d_Au, const double* d_u, double alpha,
                                  double* d_F, void* stream) {
  PSB_REQUIRE(n >= 0 && (n == 0 || (d_Au && d_u && d_F)), PSB_ERR_ARG, "psb_bratu_residual: bad argument");
  if (n == 0) return PSB_OK;
  bratu_residual_kernel<<<stream_grid(n, sm_count() * 16), kBlock, 0, (cudaStream_t)stream>>>(n, d_Au, d_u, alpha, d_F);
  PSB_LAUNCH_CHECK();
  return PSB_OK;
}

extern "C" int psb_bratu_jacobian(int64_t n, const int64_t* d_diag_pos, const double* d_a_diag,
                                  const double* d_u, double alpha, double* d_vals, void* stream) {
  PSB_REQUIRE(n >= 0 && (n == 0 || (d_diag_pos && d_a_diag && d_u && d_vals)), PSB_ERR_ARG,
              "psb_bratu_jacobian: bad argument");
  if (n == 0) return PSB_OK;
  bratu_jacobian_kernel<<<stream_grid(n, sm_count() * 16), kBlock, 0, (cudaStream_t)stream>>>(
      n, d_diag_pos, d_a_diag, d_u, alpha, d_vals);
  PSB_LAUNCH_CHECK();
  return PSB_OK;
}
