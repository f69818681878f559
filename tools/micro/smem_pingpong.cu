// Hand-over latency between two warps of one CTA through shared memory (B200 microbenchmark).
// Variants: 0 = bare token; 1 = fp64 mul+sub+mul between receive and send; 2 = variant 1 plus a
// st.relaxed.gpu global store after the shared store; 3 = variant 2 with N-2 extra warps spinning
// on a slot that never changes (issue-slot contention).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double lds_vol(uint32_t a) {
  double v; asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ void sts_vol(uint32_t a, double v) {
  asm volatile("st.volatile.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}

template <int VAR>
__global__ void pingpong(int rounds, double* gout, long long* cycles, int spinners_sleep) {
  __shared__ double slots[4096];
  __shared__ int stop;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(slots);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double NaN = __longlong_as_double(0xFFF8DEADBEEF0B20ll);
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) slots[i] = NaN;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp >= 2) {               // contention: spin on a slot nobody writes until told to stop
    if (VAR == 3) {
      volatile int* s = &stop;
      while (!*s) {
        double v = lds_vol(base + 8 * 4095);
        if ((unsigned)__double2hiint(v) != 0xFFF8DEADu) break;
        if (spinners_sleep) __nanosleep(spinners_sleep);
      }
    }
    return;
  }
  // chain: value i is produced by warp (i & 1) from value i-1; slot i (mod 2048), lane 0 only matters
  long long t0 = clock64();
  double acc = 1.0;
  if (warp == 0 && lane == 0) sts_vol(base, 1.0);
  for (int i = 1 + warp; i <= rounds; i += 2) {
    const uint32_t a = base + 8 * ((i - 1) & 2047);
    double v;
    do { v = lds_vol(a); } while ((unsigned)__double2hiint(v) == 0xFFF8DEADu);
    if (VAR >= 1) { acc = 3.0 - 0.5 * v; v = acc * 0.999; }
    sts_vol(base + 8 * (i & 2047), v);
    sts_vol(base + 8 * ((i + 1024) & 2047), NaN);      // reset a slot half a lap ahead
    if (VAR >= 2 && lane == 0) st_relaxed(gout + (i & 1023), v);
  }
  long long t1 = clock64();
  if (lane == 0) cycles[warp] = t1 - t0;
  __syncwarp();
  if (warp == 0 && lane == 0) { stop = 1; }
}

int main() {
  double* gout; long long* cyc;
  cudaMalloc(&gout, 8192); cudaMalloc(&cyc, 64);
  const int rounds = 100000;
  for (int var = 0; var <= 3; ++var) {
    for (int threads : {64, 512, 1024}) {
      if (var < 3 && threads != 64) continue;
      for (int sl : {0, 100}) {
        if (var < 3 && sl) continue;
        switch (var) {
          case 0: pingpong<0><<<1, threads>>>(rounds, gout, cyc, sl); break;
          case 1: pingpong<1><<<1, threads>>>(rounds, gout, cyc, sl); break;
          case 2: pingpong<2><<<1, threads>>>(rounds, gout, cyc, sl); break;
          case 3: pingpong<3><<<1, threads>>>(rounds, gout, cyc, sl); break;
        }
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
        printf("variant %d threads %4d spinner-sleep %3d ns: %.1f cycles per hand-over (%s)\n", var, threads, sl,
               (double)h[0] / rounds, cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
