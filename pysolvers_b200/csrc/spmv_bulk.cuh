// Bulk-async (TMA engine) staged STREAM SpMV: the tile pipeline as a device function, shared by
// the stand-alone kernel (spmv.cu) and the persistent PCG kernel (pcg_mega.cu).
#pragma once
#include "spmv.cuh"

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
// epilogue shared by all SpMV kernels.  The operands that belong to the ROW itself (f, dinv,
// x[row], y[row]) are loaded by epi_preload BEFORE the row's gathers are consumed, so that they
// travel together with them instead of costing a second memory round trip per tile.
// ---------------------------------------------------------------------------
struct EpiRow { double a, b, c; };

template <int EPI>
__device__ __forceinline__ EpiRow epi_preload(int64_t row, const double* x, const double* y, const EpiArgs& ea) {
  EpiRow o;
  o.a = 0.0; o.b = 0.0; o.c = 0.0;
  if (EPI == EPI_DOT) {
    o.a = __ldg(x + row);
  } else if (EPI == EPI_RESID) {
    o.a = ea.f[row];
  } else if (EPI == EPI_RESID_NORM) {
    o.a = ea.f[row];
    if (ea.jac_out != nullptr) { o.b = ea.dinv[row]; o.c = __ldg(x + row); }
  } else if (EPI == EPI_ADD) {
    o.a = y[row];
  } else if (EPI == EPI_JACOBI) {
    o.a = ea.f[row]; o.b = ea.dinv[row]; o.c = __ldg(x + row);
  }
  return o;
}

template <int EPI>
__device__ __forceinline__ void epi_finish(int64_t row, double sum, const EpiRow& o, double* y,
                                           const EpiArgs& ea, double& acc) {
  if (EPI == EPI_STORE) {
    y[row] = sum;
  } else if (EPI == EPI_DOT) {
    y[row] = sum;
    acc += o.a * sum;
  } else if (EPI == EPI_RESID) {
    y[row] = o.a - sum;
  } else if (EPI == EPI_RESID_NORM) {
    const double r = o.a - sum;
    y[row] = r;
    acc += r * r;
    if (ea.jac_out != nullptr) {                 // next cycle's first Jacobi sweep, as EPI_JACOBI
      double d = o.b * r;
      if (ea.omega != 1.0) d = ea.omega * d;
      ea.jac_out[row] = o.c + d;
    }
  } else if (EPI == EPI_ADD) {
    y[row] = o.a + sum;
  } else if (EPI == EPI_JACOBI) {
    double r = o.a - sum;
    double d = o.b * r;
    if (ea.omega != 1.0) d = ea.omega * d;
    y[row] = o.c + d;
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue(int64_t row, double sum, const double* x, double* y,
                                         const EpiArgs& ea, double& acc) {
  const EpiRow o = epi_preload<EPI>(row, x, y, ea);
  epi_finish<EPI>(row, sum, o, y, ea, acc);
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
// the part of the pipeline state only the producer (thread 0) touches lives in shared memory:
// persistent kernels are short of registers
struct BulkShared {
  uint64_t full[2];         // two mbarriers
  uint64_t pol;             // L2 evict-first policy
  int s_first, e_first, s_second, e_second;   // nonzero bounds of this CTA's first two tiles
  int s_next, e_next;       // bounds of the tile to be staged next, between prime and pass
};
struct BulkPipe {
  unsigned char* smem;      // dynamic shared memory
  BulkShared* sh;
  uint64_t* full;           // = sh->full
  int cap_v, cap_c;
  uint32_t phase_bits;      // parity of each stage's barrier; carried across passes
  int tile_rows;            // rows per tile at run time (multiple of 4, <= kBlock * RPT); 0 = kBlock * RPT
  // bulk_prime() already staged the first tile of the coming pass (thread 0 keeps the bounds
  // of the second one): lets a persistent kernel put the first copies in flight BEFORE it
  // waits at a grid barrier
  bool primed;
  bool issued;              // ... and a copy is really in flight (the CTA owns at least one tile)
  // the bounds of the first two tiles are iteration-invariant: cached (in sh) by a persistent
  // kernel so that priming costs no global load
  bool bounds_cached;
};

__device__ __forceinline__ void bulk_pipe_init(BulkPipe& P, unsigned char* smem, BulkShared* sh,
                                               int cap_v, int cap_c) {
  P.smem = smem; P.sh = sh; P.full = sh->full; P.cap_v = cap_v; P.cap_c = cap_c; P.phase_bits = 0u;
  P.tile_rows = 0; P.primed = false; P.issued = false;
  P.bounds_cached = false;
  if (threadIdx.x == 0) {
    mbar_init(&sh->full[0], 1);
    mbar_init(&sh->full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sh->pol = policy_evict_first();
    sh->s_first = sh->e_first = sh->s_second = sh->e_second = sh->s_next = sh->e_next = 0;
  }
  __syncthreads();
}

// Tile bookkeeping shared by bulk_prime and bulk_pass: which rows a logical tile covers and how
// its three slices are staged.  All members are cheap to rebuild (a few registers).
template <int RPT, bool C16>
struct BulkTiler {
  static constexpr int kCA = C16 ? 7 : 3;                    // column entries per 16 bytes, minus 1
  const psb_csr& A;
  const EpiArgs& ea;
  const BulkPipe& P;
  int R;                                                     // rows per tile (run time)
  int64_t n_tiles, n_int;
  int nnz_v_lim, nnz_c_lim, rp_lim;
  size_t stage_bytes;

  __device__ __forceinline__ BulkTiler(const psb_csr& A_, const EpiArgs& ea_, const BulkPipe& P_)
      : A(A_), ea(ea_), P(P_) {
    R = P.tile_rows > 0 ? P.tile_rows : kBlock * RPT;
    n_tiles = (A.n_rows + R - 1) / R;
    n_int = ea.rot_t1 - ea.rot_t0;
    nnz_v_lim = (int)(A.nnz & ~(int64_t)1);                  // bulk copies stop at the last
    nnz_c_lim = (int)(A.nnz & ~(int64_t)kCA);                // whole 16-byte chunk of each array
    rp_lim = (int)((A.n_rows + 1) & ~(int64_t)3);
    stage_bytes = (size_t)P.cap_v * 8 + (size_t)P.cap_c * 4 + (size_t)(kBlock * RPT + 4) * 4;
  }
  // logical -> physical tile: interior tiles [rot_t0, rot_t1) first (multi-GPU overlap)
  __device__ __forceinline__ int64_t phys(int64_t t) const {
    return t < n_int ? ea.rot_t0 + t : (t < ea.rot_t1 ? t - n_int : t);
  }
  __device__ __forceinline__ double* stage_vals(int st) const {
    return reinterpret_cast<double*>(P.smem + st * stage_bytes);
  }
  __device__ __forceinline__ int* stage_cols(int st) const {
    return reinterpret_cast<int*>(P.smem + st * stage_bytes + (size_t)P.cap_v * 8);
  }
  __device__ __forceinline__ int* stage_rp(int st) const {
    return reinterpret_cast<int*>(P.smem + st * stage_bytes + (size_t)P.cap_v * 8 + (size_t)P.cap_c * 4);
  }
  __device__ __forceinline__ int cap_c_entries() const { return C16 ? 2 * P.cap_c : P.cap_c; }
  __device__ __forceinline__ void tile_bounds(int64_t lt, int& s, int& e) const {
    const int64_t row0 = phys(lt) * R;
    s = ld_stream_i(A.rowptr + row0);
    e = ld_stream_i(A.rowptr + min(row0 + R, A.n_rows));
  }
  // producer (thread 0): stage tile `lt` whose nonzero range is [s, e).  The byte counts are
  // clamped to the stage capacity: a tile fuller than the capacity the host sized the stages for
  // (it never is when the capacity comes from the statistics of THIS tiling) cannot overrun
  // shared memory -- its surplus entries are fetched by the fix-up loads in bulk_pass.
  __device__ __forceinline__ void issue(int64_t lt, int st, int s, int e) const {
    const int64_t row0 = phys(lt) * R;
    const int nr = (int)min((int64_t)R, A.n_rows - row0);
    const int v0 = s & ~1, c0 = s & ~kCA;
    const int v1 = min(min((e + 1) & ~1, nnz_v_lim), v0 + (P.cap_v & ~1));
    const int c1 = min(min((e + kCA) & ~kCA, nnz_c_lim), c0 + (cap_c_entries() & ~kCA));
    const int r1 = (int)min((int64_t)((nr + 1 + 3) & ~3), (int64_t)rp_lim - row0);
    const uint32_t bv = v1 > v0 ? (uint32_t)(v1 - v0) * 8u : 0u;
    const uint32_t bc = c1 > c0 ? (uint32_t)(c1 - c0) * (C16 ? 2u : 4u) : 0u;
    const uint32_t br = r1 > 0 ? (uint32_t)r1 * 4u : 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&P.full[st], bv + bc + br);
    const uint64_t pol = P.sh->pol;
    if (bv) bulk_g2s(stage_vals(st), A.vals + v0, bv, &P.full[st], pol);
    if (bc) {
      if (C16) bulk_g2s(stage_cols(st), A.colind16 + c0, bc, &P.full[st], pol);
      else     bulk_g2s(stage_cols(st), A.colind + c0, bc, &P.full[st], pol);
    }
    if (br) bulk_g2s(stage_rp(st), A.rowptr + row0, br, &P.full[st], pol);
  }
};

// Put the copies of this CTA's first tile in flight (stage 0 is free between passes).  The
// matrix arrays never change, so a persistent kernel calls this BEFORE waiting at the barrier
// that precedes the pass; bulk_pass then finds the tile staged.
template <int RPT, bool C16 = false>
__device__ __forceinline__ void bulk_prime(const psb_csr& A, const EpiArgs& ea, BulkPipe& P) {
  if (P.primed) return;
  const BulkTiler<RPT, C16> T(A, ea, P);
  const int64_t tile = blockIdx.x;
  if (threadIdx.x == 0 && tile < T.n_tiles) {
    BulkShared* sh = P.sh;
    if (!P.bounds_cached) {
      T.tile_bounds(tile, sh->s_first, sh->e_first);
      if (tile + gridDim.x < T.n_tiles) T.tile_bounds(tile + gridDim.x, sh->s_second, sh->e_second);
    }
    T.issue(tile, 0, sh->s_first, sh->e_first);
    sh->s_next = sh->s_second; sh->e_next = sh->e_second;
  }
  P.primed = true;
  P.issued = tile < T.n_tiles;
}

// Persistent kernels: compute the bounds of the first two tiles once (tile order must not change
// afterwards), so that bulk_prime issues its copies without touching global memory.
template <int RPT, bool C16 = false>
__device__ __forceinline__ void bulk_cache_bounds(const psb_csr& A, const EpiArgs& ea, BulkPipe& P) {
  const BulkTiler<RPT, C16> T(A, ea, P);
  const int64_t tile = blockIdx.x;
  if (threadIdx.x == 0 && tile < T.n_tiles) {
    T.tile_bounds(tile, P.sh->s_first, P.sh->e_first);
    if (tile + gridDim.x < T.n_tiles) T.tile_bounds(tile + gridDim.x, P.sh->s_second, P.sh->e_second);
  }
  P.bounds_cached = true;
}

// A primed pass that will not run (the solver stopped): wait for the staged tile so that no
// bulk copy is in flight when the CTA exits.
__device__ __forceinline__ void bulk_drain(BulkPipe& P) {
  if (P.primed && P.issued) {
    while (!mbar_try_wait(&P.full[0], P.phase_bits & 1u)) {}
    P.phase_bits ^= 1u;
  }
  P.primed = false; P.issued = false;
}

// One pass over all tiles of A assigned to this CTA (grid-stride).  `beta` is used by
// EPI_DOT_PUP only.  x may have been written earlier by this same kernel (persistent PCG):
// it is read through the coherent path.
// G: gathers of a row whose loads are all issued before the first product is formed (rows with
// more entries continue four at a time).
template <int EPI, int RPT, bool C16 = false, int G = 4>
__device__ __forceinline__ void bulk_pass(const psb_csr& A, const double* x, double* y,
                                          const EpiArgs& ea, const double beta, BulkPipe& P,
                                          double& acc) {
  constexpr int kCA = C16 ? 7 : 3;
  bulk_prime<RPT, C16>(A, ea, P);
  const BulkTiler<RPT, C16> T(A, ea, P);
  uint64_t* const full = P.full;
  const int tid = threadIdx.x;
  const int R = T.R;
  const int64_t n_tiles = T.n_tiles;
  const int64_t n_int = T.n_int;
  bool waited = (ea.wait_n == 0);
  const int nnz_v_lim = T.nnz_v_lim, nnz_c_lim = T.nnz_c_lim, rp_lim = T.rp_lim;

  int64_t tile = blockIdx.x;
  int s_next = 0, e_next = 0;                     // bounds of the tile to be staged next (thread 0 only)
  if (tid == 0) { s_next = P.sh->s_next; e_next = P.sh->e_next; }

  uint32_t phase_bits = P.phase_bits;
  int st = 0;
  for (; tile < n_tiles; tile += gridDim.x, st ^= 1) {
    const int64_t nxt = tile + gridDim.x;
    if (tid == 0 && nxt < n_tiles) {
      T.issue(nxt, st ^ 1, s_next, e_next);            // stage st^1 was released by the
      if (nxt + gridDim.x < n_tiles)                   // __syncthreads of the previous trip
        T.tile_bounds(nxt + gridDim.x, s_next, e_next);
    }
    while (!mbar_try_wait(&full[st], (phase_bits >> st) & 1u)) {}
    phase_bits ^= (1u << st);
    if (!waited && tile >= n_int) { wait_for_halo(ea); waited = true; }   // uniform per CTA

    const int64_t row0 = T.phys(tile) * R;
    const int nr = (int)min((int64_t)R, A.n_rows - row0);
    const double* sv = T.stage_vals(st);
    const int*    sc = T.stage_cols(st);
    int*          rp = T.stage_rp(st);

    // Bulk copies stop at the last whole 16-byte chunk of each array (and at the stage
    // capacity); the few elements past it (they can only matter to the tiles at the very end
    // of the matrix) are fetched with ordinary loads.  All conditions are uniform across the CTA.
    if (row0 + nr + 1 > rp_lim) {
      for (int64_t i = max(row0, (int64_t)rp_lim) + tid; i <= row0 + nr; i += kBlock)
        rp[i - row0] = A.rowptr[i];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }
    {
      const int s0 = rp[0], e0 = rp[nr];
      const int v_end = min(nnz_v_lim, (s0 & ~1) + (P.cap_v & ~1));      // first entry NOT bulk-copied
      const int c_end = min(nnz_c_lim, (s0 & ~kCA) + (T.cap_c_entries() & ~kCA));
      if (e0 > v_end || e0 > c_end) {
        double* svw = T.stage_vals(st);
        int*    scw = T.stage_cols(st);
        // entries past the capacity do not fit the stage at all: host-side sizing error
        if (e0 - (s0 & ~1) > P.cap_v || e0 - (s0 & ~kCA) > T.cap_c_entries()) {
          if (ea.error_flag) *ea.error_flag = 2;
        } else {
          for (int i = max(s0, v_end) + tid; i < e0; i += kBlock) svw[i - (s0 & ~1)] = A.vals[i];
          if (C16) {
            short* scw16 = reinterpret_cast<short*>(scw);
            for (int i = max(s0, c_end) + tid; i < e0; i += kBlock) scw16[i - (s0 & ~kCA)] = A.colind16[i];
          } else {
            for (int i = max(s0, c_end) + tid; i < e0; i += kBlock) scw[i - (s0 & ~kCA)] = A.colind[i];
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
      }
    }

    const int s = rp[0];
    const int offv = s & ~1, offc = s & ~kCA;
    const short* sc16 = reinterpret_cast<const short*>(sc);
    const long long win0 = A.row_off + row0;
    // One thread per row.  Order of the memory operations: (1) the row's own operands and the
    // first G gathers are ALL loaded before anything is consumed or stored -- one memory round
    // trip per tile, not one per dependent step (stores in between would also keep the compiler
    // from hoisting the later loads: the vectors may alias for all it knows); (2) products added
    // in stored order from +0 (bit-identical to csr_matvec); (3) stores.
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int lr = tid + j * kBlock;
      if (lr < nr) {
        const int64_t row = win0 + lr;
        const int a = rp[lr], b = rp[lr + 1];
        const int rabs = (int)row;                      // C16: columns are stored relative to the row
        auto col = [&](int kk) -> int { return C16 ? rabs + (int)sc16[kk - offc] : sc[kk - offc]; };
        // (1) own-row operands
        double po = 0.0, rown = 0.0, xo = 0.0;
        EpiRow own;
        if constexpr (EPI == EPI_DOT_PUP) {
          po = ld_ca(ea.pold + row);
          rown = ld_ca(x + row);
          if (ea.xsol != nullptr) xo = ea.xsol[row];
        } else {
          own = epi_preload<EPI>(row, x, y, ea);
        }
        // first G gathers, predicated on the row length
        int cg[G];
        double gx[G], gp[G];
#pragma unroll
        for (int g = 0; g < G; ++g) cg[g] = (a + g < b) ? col(a + g) : -1;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          gx[g] = 0.0; gp[g] = 0.0;
          if (cg[g] >= 0) {
            gx[g] = ld_ca(x + cg[g]);
            if constexpr (EPI == EPI_DOT_PUP) gp[g] = ld_ca(ea.pold + cg[g]);
          }
        }
        // (2) sum in stored order
        double sum = 0.0;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (a + g < b) {
            const double xv = (EPI == EPI_DOT_PUP) ? gx[g] + beta * gp[g] : gx[g];   // p as K3 would store it
            sum += sv[a + g - offv] * xv;
          }
        }
        auto gather = [&](int c) -> double {
          if constexpr (EPI == EPI_DOT_PUP) return ld_ca(x + c) + beta * ld_ca(ea.pold + c);
          else return ld_ca(x + c);
        };
        int k = a + G;
        for (; k + 4 <= b; k += 4) {                   // longer rows: 4 independent gathers in flight
          const int c0 = col(k), c1 = col(k + 1), c2 = col(k + 2), c3 = col(k + 3);
          const double x0 = gather(c0), x1 = gather(c1), x2 = gather(c2), x3 = gather(c3);
          sum += sv[k - offv] * x0;
          sum += sv[k + 1 - offv] * x1;
          sum += sv[k + 2 - offv] * x2;
          sum += sv[k + 3 - offv] * x3;
        }
        for (; k < b; ++k) sum += sv[k - offv] * gather(col(k));
        // (3) stores
        if constexpr (EPI == EPI_DOT_PUP) {
          const double pn = rown + beta * po;                                   // as K3 would store it
          ea.pnew[row] = pn;
          y[row] = sum;
          // deferred solution update of the PREVIOUS iteration, x += alpha_{k-1} p_{k-1}
          // (PCGSolver.py:121): p_{k-1} is in a register here anyway, so the update pass never
          // has to read p -- same operands, same rounding, 8 bytes per row less traffic
          if (ea.xsol != nullptr) ea.xsol[row] = xo + ea.alpha_prev * po;
          if (ea.pp_n > 0 && (rabs < ea.pp_skip_lo || rabs >= ea.pp_skip_hi)) {
#pragma unroll
            for (int q = 0; q < 4; ++q)               // boundary rows also go to the neighbours' halo
              if (q < ea.pp_n && row >= ea.pp_off[q] && row < ea.pp_off[q] + ea.pp_cnt[q])
                ea.pp_remote[q][row - ea.pp_off[q]] = pn;
          }
          acc += pn * sum;
        } else {
          epi_finish<EPI>(row, sum, own, y, ea, acc);
        }
      }
    }
    __syncthreads();                                   // stage st may be refilled now
  }
  P.phase_bits = phase_bits;
  P.primed = false; P.issued = false;
}

}  // namespace psb
