// Shared device/host helpers for libpysolv_b200 (sm_100a only).
//
// Arithmetic policy: the whole library is compiled with -fmad=false so that a*b+c
// rounds twice, exactly like the numpy ufunc loops and scipy's csr_matvec the
// reference executes on x86-64 (no FMA contraction there).  The path is HBM-bound
// fp64 work, so giving up FMA costs nothing; it removes one source of divergence
// from the reference's residual histories ("Arithmetic").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <string>

#include "../../include/pysolv_b200.h"

namespace psb {

// ----------------------------------------------------------------------------
// host side: error text, launch accounting, device properties
// ----------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;   // kernels launched by this library
int  sm_count();                            // multiprocessors of the current device
int  max_optin_smem();                      // bytes of opt-in shared memory per block

#define PSB_CUDA(expr)                                                          \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      psb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,              \
                     cudaGetErrorString(_e));                                   \
      return PSB_ERR_CUDA;                                                      \
    }                                                                           \
  } while (0)

#define PSB_LAUNCH_CHECK()                                                      \
  do {                                                                          \
    psb::g_launches.fetch_add(1, std::memory_order_relaxed);                    \
    cudaError_t _e = cudaPeekAtLastError();                                     \
    if (_e != cudaSuccess) {                                                    \
      psb::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,          \
                     cudaGetErrorString(_e));                                   \
      return PSB_ERR_CUDA;                                                      \
    }                                                                           \
  } while (0)

#define PSB_REQUIRE(cond, code, msg)                                            \
  do {                                                                          \
    if (!(cond)) {                                                              \
      psb::set_error("%s:%d: %s", __FILE__, __LINE__, msg);                     \
      return code;                                                              \
    }                                                                           \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ----------------------------------------------------------------------------
// device side
// ----------------------------------------------------------------------------
constexpr int kBlock = 256;          // threads per CTA for the streaming kernels
constexpr int kWarps = kBlock / 32;

// Streaming 128-bit loads/stores that do not allocate in L1 (each element is
// touched once per kernel; L1 is kept for the x gathers of the SpMV).
__device__ __forceinline__ double2 ld_stream2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
               : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int4 ld_stream_i4(const int* p) {
  int4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream_i(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
// plain (coherent) variants for data that an earlier block of the SAME kernel wrote
__device__ __forceinline__ double ld_cg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ int    ld_cg(const int* p)    { return __ldcg(p); }

// L1-cached load on the coherent path (not .nc): for vectors a peer GPU or an earlier phase of
// the same kernel may have written before this CTA first touches them
__device__ __forceinline__ double ld_ca(const double* p) {
  double v;
  asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_vol(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_vol(double* p, double v) {
  asm volatile("st.volatile.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_vol_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
constexpr int kPeerSpinLimit = 1 << 24;   // ~5-10 s of polling before a wait is declared dead

// Scalar exchange between GPUs over NVLink peer memory, NCCL-LL style: a double travels as two
// 8-byte words, each carrying 32 data bits and the 32-bit epoch of the reduction it belongs
// to.  8-byte stores are atomic, so a reader that sees the expected epoch in both words has the
// whole value; epochs only grow, so slots never need to be reset.
__device__ __forceinline__ void peer_push(unsigned long long* slot, double v, unsigned int epoch) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned long long tag = (unsigned long long)epoch << 32;
  asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(slot), "l"(tag | (bits & 0xffffffffull)) : "memory");
  asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(slot + 1), "l"(tag | (bits >> 32)) : "memory");
}
// returns false on time-out
__device__ __forceinline__ bool peer_wait(const unsigned long long* slot, unsigned int epoch, double* out) {
  int spins = 0;
  for (;;) {
    const unsigned long long w0 = ld_vol_u64(slot), w1 = ld_vol_u64(slot + 1);
    if ((unsigned int)(w0 >> 32) == epoch && (unsigned int)(w1 >> 32) == epoch) {
      *out = __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
      return true;
    }
    if (++spins > kPeerSpinLimit) { *out = 0.0; return false; }
  }
}

__device__ __forceinline__ void st_stream2(double* p, double2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};"
               :: "l"(p), "d"(v.x), "d"(v.y) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;   // butterfly: every lane holds the same, order-fixed sum
}

// Deterministic CTA reduction (fixed tree: xor-butterfly inside warps, then the
// kWarps warp sums added in warp order).  Every thread returns the total.
// `scratch` must hold kWarps doubles; safe to call repeatedly (syncs both sides).
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) t += scratch[w];
  return t;
}

// "Last CTA finishes the sum" ticket.  Each CTA first publishes its partial(s),
// then takes a ticket; the CTA that draws the last ticket sees every partial.
// atomicInc wraps the counter back to 0, so it is ready for the next launch.
__device__ __forceinline__ bool last_block(unsigned int* ticket) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicInc(ticket, gridDim.x - 1);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// Sum `count` per-CTA partials (stride 1) in a fixed order; all threads of the
// calling CTA participate and receive the total.
__device__ __forceinline__ double sum_partials(const double* partials, int count,
                                               double* scratch) {
  double a = 0.0;
  for (int i = threadIdx.x; i < count; i += kBlock) a += ld_cg(partials + i);
  return block_sum(a, scratch);
}

}  // namespace psb
