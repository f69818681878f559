// Benchmark-input generators on the device: the finite-difference
// Laplacians of examples/FDLaplacian2D.py:5-23 (2-D, 5-point) and its builder-defined 7-point
// 3-D extension, assembled directly in HBM as CSR with the reference's STORED column order
// ([k, k-m, k+m, k-1, k+1] resp. [k, k-m^2, k+m^2, k-m, k+m, k-1, k+1], out-of-grid neighbours
// left out), for a slab of rows [row_lo, row_hi) with global column numbers.  Bit-identical to
// pysolvers_b200.problems.fd_laplacian_2d / _3d: the two distinct values are computed by the
// caller with the reference's own expression (-4/h/h, 1/h/h) and passed in.
#include "common.cuh"
#include "prec.cuh"

namespace psb {

// entries stored in rows [0, k) of the full grid matrix
__host__ __device__ inline long long stencil_prefix(int dim, long long m, long long k) {
  if (dim == 2) {
    const long long n = m * m;
    if (k > n) k = n;
    return 5 * k - (k < m ? k : m)                         // rows with iy = 0 lack k - m
           - (k > (m - 1) * m ? k - (m - 1) * m : 0)      // rows with iy = m-1 lack k + m
           - (k + m - 1) / m                               // rows with ix = 0 lack k - 1
           - k / m;                                        // rows with ix = m-1 lack k + 1
  }
  const long long mm = m * m, n = mm * m;
  if (k > n) k = n;
  const long long planes = k / mm, rem = k % mm;
  return 7 * k - (k < mm ? k : mm)                                        // iz = 0
         - (k > (m - 1) * mm ? k - (m - 1) * mm : 0)                      // iz = m-1
         - (planes * m + (rem < m ? rem : m))                             // iy = 0
         - (planes * m + (rem > (m - 1) * m ? rem - (m - 1) * m : 0))     // iy = m-1
         - (k + m - 1) / m                                                // ix = 0
         - k / m;                                                         // ix = m-1
}

template <int DIM>
__global__ void __launch_bounds__(kBlock)
stencil_fill_kernel(long long m, long long row_lo, long long row_hi, double diag, double off,
                    int* __restrict__ rowptr, int* __restrict__ colind, double* __restrict__ vals) {
  const long long base = stencil_prefix(DIM, m, row_lo);
  const long long mm = m * m;
  for (long long k = row_lo + blockIdx.x * (long long)kBlock + threadIdx.x; k <= row_hi;
       k += (long long)gridDim.x * kBlock) {
    long long p = stencil_prefix(DIM, m, k) - base;
    rowptr[k - row_lo] = (int)p;
    if (k == row_hi) break;
    const long long ix = k % m, iy = (k / m) % m, iz = k / mm;
    colind[p] = (int)k; vals[p] = diag; ++p;
    if (DIM == 3) {
      if (iz > 0)     { colind[p] = (int)(k - mm); vals[p] = off; ++p; }
      if (iz < m - 1) { colind[p] = (int)(k + mm); vals[p] = off; ++p; }
    }
    if (iy > 0)     { colind[p] = (int)(k - m); vals[p] = off; ++p; }
    if (iy < m - 1) { colind[p] = (int)(k + m); vals[p] = off; ++p; }
    if (ix > 0)     { colind[p] = (int)(k - 1); vals[p] = off; ++p; }
    if (ix < m - 1) { colind[p] = (int)(k + 1); vals[p] = off; ++p; }
  }
}

}  // namespace psb

using namespace psb;

extern "C" int64_t psb_stencil_nnz(int dim, int64_t m, int64_t row_lo, int64_t row_hi) {
  if ((dim != 2 && dim != 3) || m < 1 || row_lo < 0 || row_hi < row_lo) return -1;
  return (int64_t)(stencil_prefix(dim, m, row_hi) - stencil_prefix(dim, m, row_lo));
}

extern "C" int psb_stencil_fill(int dim, int64_t m, int64_t row_lo, int64_t row_hi, double diag, double off,
                                int32_t* d_rowptr, int32_t* d_colind, double* d_vals, void* stream) {
  PSB_REQUIRE(dim == 2 || dim == 3, PSB_ERR_ARG, "psb_stencil_fill: dim must be 2 or 3");
  PSB_REQUIRE(m >= 1 && row_lo >= 0 && row_hi >= row_lo, PSB_ERR_ARG, "psb_stencil_fill: bad row range");
  const long long n = dim == 2 ? (long long)m * m : (long long)m * m * m;
  PSB_REQUIRE(row_hi <= n && n < (long long)INT32_MAX, PSB_ERR_UNSUPP, "psb_stencil_fill: int32 index range exceeded");
  PSB_REQUIRE(psb_stencil_nnz(dim, m, row_lo, row_hi) < (int64_t)INT32_MAX, PSB_ERR_UNSUPP,
              "psb_stencil_fill: int32 index range exceeded");
  PSB_REQUIRE(d_rowptr && (row_hi == row_lo || (d_colind && d_vals)), PSB_ERR_ARG, "psb_stencil_fill: NULL array");
  const int grid = stream_grid(row_hi - row_lo + 1, sm_count() * 16);
  if (dim == 2)
    stencil_fill_kernel<2><<<grid, kBlock, 0, (cudaStream_t)stream>>>(m, row_lo, row_hi, diag, off, d_rowptr, d_colind, d_vals);
  else
    stencil_fill_kernel<3><<<grid, kBlock, 0, (cudaStream_t)stream>>>(m, row_lo, row_hi, diag, off, d_rowptr, d_colind, d_vals);
  PSB_LAUNCH_CHECK();
  return PSB_OK;
}
