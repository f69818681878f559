// Bulk-async (TMA engine) staged STREAM SpMV: the tile pipeline as a device function, shared by
// the stand-alone kernel (spmv.cu) and the persistent PCG kernel (pcg_mega.cu).
#pragma once
#include "spmv.cuh"

#ifndef PSB_PUP_WINDOW
#define PSB_PUP_WINDOW 0
#endif

namespace psb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// global -> shared bulk copy (TMA engine), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                         uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
               "[%0], [%1], %2, [%3], %4;"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}


// ---------------------------------------------------------------------------
// epilogue shared by both kernels
// ---------------------------------------------------------------------------
template <int EPI>
__device__ __forceinline__ void epilogue(int64_t row, double sum, const double* x, double* y,
                                         const EpiArgs& ea, double& acc) {
  if (EPI == EPI_STORE) {
    y[row] = sum;
  } else if (EPI == EPI_DOT) {
    y[row] = sum;
    acc += __ldg(x + row) * sum;
  } else if (EPI == EPI_RESID) {
    y[row] = ea.f[row] - sum;
  } else if (EPI == EPI_ADD) {
    y[row] = y[row] + sum;
  } else if (EPI == EPI_JACOBI) {
    double r = ea.f[row] - sum;
    double d = ea.dinv[row] * r;
    if (ea.omega != 1.0) d = ea.omega * d;
    y[row] = __ldg(x + row) + d;
  }
}


// Multi-GPU: block until the neighbours have pushed this iteration's halo (dist.cu, pcg_mega.cu).
__device__ __forceinline__ void wait_for_halo(const EpiArgs& ea) {
  if (ea.wait_n == 0) return;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ea.wait_n; ++i) {
      int spins = 0;
      while (ld_vol_u64(ea.wait_flags + i) < ea.wait_value) {
        if (++spins > kPeerSpinLimit) { if (ea.error_flag) *ea.error_flag = 1; break; }
      }
    }
  }
  __syncthreads();
  __threadfence();          // acquire: the halo data stored before the flag is visible to weak loads
}

// Double-buffered staging area of one CTA: two stages of
//   vals[cap_v] doubles | cols[cap_c] ints | rp[R + 4] ints     (all 16-byte aligned)
struct BulkPipe {
  unsigned char* smem;      // dynamic shared memory
  uint64_t* full;           // two mbarriers
  int cap_v, cap_c;
  uint32_t phase_bits;      // parity of each stage's barrier; carried across passes
  uint64_t pol;             // L2 evict-first policy (thread 0)
};

__device__ __forceinline__ void bulk_pipe_init(BulkPipe& P, unsigned char* smem, uint64_t* full,
                                               int cap_v, int cap_c) {
  P.smem = smem; P.full = full; P.cap_v = cap_v; P.cap_c = cap_c; P.phase_bits = 0u; P.pol = 0;
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    P.pol = policy_evict_first();
  }
  __syncthreads();
}

// One pass over all tiles of A assigned to this CTA (grid-stride).  `beta` is used by
// EPI_DOT_PUP only.  x may have been written earlier by this same kernel (persistent PCG):
// it is read through the coherent path.
template <int EPI, int RPT, bool C16 = false>
__device__ __forceinline__ void bulk_pass(const psb_csr& A, const double* x, double* y,
                                          const EpiArgs& ea, const double beta, BulkPipe& P,
                                          double& acc) {
  constexpr int R = kBlock * RPT;
  const uint64_t pol = P.pol;
  uint64_t* const full = P.full;
  unsigned char* const smem_raw = P.smem;
  const int cap_v = P.cap_v, cap_c = P.cap_c;
  const int tid = threadIdx.x;
  const size_t stage_bytes = (size_t)cap_v * 8 + (size_t)cap_c * 4 + (size_t)(R + 4) * 4;
  const int64_t n_tiles = (A.n_rows + R - 1) / R;
  // logical -> physical tile: interior tiles [rot_t0, rot_t1) first (multi-GPU overlap)
  const int64_t n_int = ea.rot_t1 - ea.rot_t0;
  auto phys = [&](int64_t t) -> int64_t {
    return t < n_int ? ea.rot_t0 + t : (t < ea.rot_t1 ? t - n_int : t);
  };
  bool waited = (ea.wait_n == 0);
  const int nnz_v_lim = (int)(A.nnz & ~(int64_t)1);          // bulk copies stop at the last
  constexpr int kCA = C16 ? 7 : 3;                           // column entries per 16 bytes, minus 1
  const int nnz_c_lim = (int)(A.nnz & ~(int64_t)kCA);        // whole 16-byte chunk of each array
  const int rp_lim = (int)((A.n_rows + 1) & ~(int64_t)3);

  auto stage_vals = [&](int st) { return reinterpret_cast<double*>(smem_raw + st * stage_bytes); };
  auto stage_cols = [&](int st) { return reinterpret_cast<int*>(smem_raw + st * stage_bytes + (size_t)cap_v * 8); };
  auto stage_rp = [&](int st) {
    return reinterpret_cast<int*>(smem_raw + st * stage_bytes + (size_t)cap_v * 8 + (size_t)cap_c * 4);
  };


  // producer (thread 0): stage tile `t` whose nonzero range is [s, e)
  auto issue = [&](int64_t lt, int st, int s, int e) {
    const int64_t row0 = phys(lt) * R;
    const int nr = (int)min((int64_t)R, A.n_rows - row0);
    const int v0 = s & ~1, c0 = s & ~kCA;
    const int v1 = min((e + 1) & ~1, nnz_v_lim);
    const int c1 = min((e + kCA) & ~kCA, nnz_c_lim);
    const int r1 = (int)min((int64_t)((nr + 1 + 3) & ~3), (int64_t)rp_lim - row0);
    const uint32_t bv = v1 > v0 ? (uint32_t)(v1 - v0) * 8u : 0u;
    const uint32_t bc = c1 > c0 ? (uint32_t)(c1 - c0) * (C16 ? 2u : 4u) : 0u;
    const uint32_t br = r1 > 0 ? (uint32_t)r1 * 4u : 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&full[st], bv + bc + br);
    if (bv) bulk_g2s(stage_vals(st), A.vals + v0, bv, &full[st], pol);
    if (bc) {
      if (C16) bulk_g2s(stage_cols(st), A.colind16 + c0, bc, &full[st], pol);
      else     bulk_g2s(stage_cols(st), A.colind + c0, bc, &full[st], pol);
    }
    if (br) bulk_g2s(stage_rp(st), A.rowptr + row0, br, &full[st], pol);
  };
  auto tile_bounds = [&](int64_t lt, int& s, int& e) {
    const int64_t row0 = phys(lt) * R;
    s = ld_stream_i(A.rowptr + row0);
    e = ld_stream_i(A.rowptr + min(row0 + R, A.n_rows));
  };

  int64_t tile = blockIdx.x;
  int s_next = 0, e_next = 0;       // bounds of the tile to be staged next (thread 0 only)
  if (tid == 0 && tile < n_tiles) {
    int s, e;
    tile_bounds(tile, s, e);
    issue(tile, 0, s, e);
    if (tile + gridDim.x < n_tiles) tile_bounds(tile + gridDim.x, s_next, e_next);
  }

  uint32_t phase_bits = P.phase_bits;
  int st = 0;
  for (; tile < n_tiles; tile += gridDim.x, st ^= 1) {
    const int64_t nxt = tile + gridDim.x;
    if (tid == 0 && nxt < n_tiles) {
      issue(nxt, st ^ 1, s_next, e_next);              // stage st^1 was released by the
      if (nxt + gridDim.x < n_tiles)                   // __syncthreads of the previous trip
        tile_bounds(nxt + gridDim.x, s_next, e_next);
    }
    while (!mbar_try_wait(&full[st], (phase_bits >> st) & 1u)) {}
    phase_bits ^= (1u << st);
    if (!waited && tile >= n_int) { wait_for_halo(ea); waited = true; }   // uniform per CTA

    const int64_t row0 = phys(tile) * R;
    const int nr = (int)min((int64_t)R, A.n_rows - row0);
    const double* sv = stage_vals(st);
    const int*    sc = stage_cols(st);
    int*          rp = stage_rp(st);

    // Bulk copies stop at the last whole 16-byte chunk of each array; the few elements
    // past it (they can only matter to the tiles at the very end of the matrix) are
    // fetched with ordinary loads.  All conditions are uniform across the CTA.
    if (row0 + nr + 1 > rp_lim) {
      for (int64_t i = max(row0, (int64_t)rp_lim) + tid; i <= row0 + nr; i += kBlock)
        rp[i - row0] = A.rowptr[i];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }
    {
      const int s0 = rp[0], e0 = rp[nr];
      if (e0 > nnz_v_lim || e0 > nnz_c_lim) {
        double* svw = stage_vals(st);
        int*    scw = stage_cols(st);
        for (int i = max(s0, nnz_v_lim) + tid; i < e0; i += kBlock) svw[i - (s0 & ~1)] = A.vals[i];
        if (C16) {
          short* scw16 = reinterpret_cast<short*>(scw);
          for (int i = max(s0, nnz_c_lim) + tid; i < e0; i += kBlock) scw16[i - (s0 & ~kCA)] = A.colind16[i];
        } else {
          for (int i = max(s0, nnz_c_lim) + tid; i < e0; i += kBlock) scw[i - (s0 & ~kCA)] = A.colind[i];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
      }
    }

    const int s = rp[0];
    const int offv = s & ~1, offc = s & ~kCA;
    const short* sc16 = reinterpret_cast<const short*>(sc);
    // EPI_DOT_PUP: the tile's own rows of p = x + beta * pold are formed once, kept in shared
    // memory and stored; gathers that fall inside the tile's row window (the k, k-1, k+1 entries
    // of a stencil) read them back instead of two global gathers each.
    __shared__ double ptile[EPI == EPI_DOT_PUP ? R : 1];
    constexpr bool kWindow = PSB_PUP_WINDOW != 0;       // serve in-tile gathers from ptile
    const long long win0 = A.row_off + row0;
    if constexpr (EPI == EPI_DOT_PUP) {
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const int lr = tid + j * kBlock;
        if (lr < nr) {
          const int64_t row = win0 + lr;
          const double pn = ld_ca(x + row) + beta * ld_ca(ea.pold + row);   // as K3 would store it
          ptile[lr] = pn;
          ea.pnew[row] = pn;
#pragma unroll
          for (int q = 0; q < 4; ++q)                 // boundary rows also go to the neighbours' halo
            if (q < ea.pp_n && row >= ea.pp_off[q] && row < ea.pp_off[q] + ea.pp_cnt[q])
              ea.pp_remote[q][row - ea.pp_off[q]] = pn;
        }
      }
      if (kWindow) __syncthreads();
    }
    auto gather = [&](int c) -> double {
      if constexpr (EPI == EPI_DOT_PUP) {
        const long long d = (long long)c - win0;
        if (kWindow && d >= 0 && d < nr) return ptile[d];
        return ld_ca(x + c) + beta * ld_ca(ea.pold + c);
      } else {
        return ld_ca(x + c);
      }
    };
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int lr = tid + j * kBlock;
      if (lr < nr) {
        const int a = rp[lr], b = rp[lr + 1];
        double sum = 0.0;
        int k = a;
        const int rabs = (int)(A.row_off + row0 + lr);  // C16: columns are stored relative to the row
        auto col = [&](int kk) -> int { return C16 ? rabs + (int)sc16[kk - offc] : sc[kk - offc]; };
        for (; k + 4 <= b; k += 4) {                   // 4 independent gathers in flight
          const int c0 = col(k), c1 = col(k + 1), c2 = col(k + 2), c3 = col(k + 3);
          const double x0 = gather(c0), x1 = gather(c1), x2 = gather(c2), x3 = gather(c3);
          sum += sv[k - offv] * x0;
          sum += sv[k + 1 - offv] * x1;
          sum += sv[k + 2 - offv] * x2;
          sum += sv[k + 3 - offv] * x3;
        }
        for (; k < b; ++k) {
          sum += sv[k - offv] * gather(col(k));
        }
        if constexpr (EPI == EPI_DOT_PUP) {
          y[win0 + lr] = sum;
          acc += ptile[lr] * sum;                   // own element: written by this very thread
        } else {
          epilogue<EPI>(A.row_off + row0 + lr, sum, x, y, ea, acc);
        }
      }
    }
    __syncthreads();                                   // stage st may be refilled now
  }
  P.phase_bits = phase_bits;
}

}  // namespace psb
