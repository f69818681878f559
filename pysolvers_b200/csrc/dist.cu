// Row-partitioned PCG across the GPUs of one NVSwitch box: one process per GPU.
//
// The reference has no distributed code; this is the
// multi-GPU contract of SURVEY.md section 8e.  Each rank owns a contiguous block of rows
// (local CSR with columns renumbered: owned first, then halo columns in sorted-global
// order) and the matching slices of x, r, p.  Per iteration the only traffic is
//   * the SpMV halo: p's boundary entries to the neighbouring ranks (ncclSend/ncclRecv
//     grouped, on a side stream), overlapped with the SpMV of the interior rows -- rows
//     that touch halo columns run after the halo has landed;
//   * two scalar all-reduces (p.Ap; r.r), issued on the compute stream between the
//     kernels that produce and consume them.
// Vector updates are local.  Every rank takes identical decisions (they all see the same
// all-reduced scalars), and the host loops poll their state snapshots at the same logical
// points, so all ranks enqueue the same sequence of collectives.
//
// NCCL is resolved at run time from the libnccl.so.2 torch has already loaded (dlopen), so
// the single-GPU path has no NCCL dependency.
#include "pcg_mega.cuh"
#include "prec.cuh"
#include "spmv.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <new>
#include <vector>

namespace psb {

// ---------------------------------------------------------------------------
// NCCL entry points, looked up once
// ---------------------------------------------------------------------------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.ok) return PSB_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { set_error("NCCL: cannot load libnccl.so.2 (%s)", dlerror()); return PSB_ERR_NCCL; }
  g_nccl.handle = h;
#define PSB_SYM(field, name)                                                      \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));        \
  if (!g_nccl.field) { set_error("NCCL: missing symbol %s", name); return PSB_ERR_NCCL; }
  PSB_SYM(GetUniqueId, "ncclGetUniqueId");
  PSB_SYM(CommInitRank, "ncclCommInitRank");
  PSB_SYM(CommDestroy, "ncclCommDestroy");
  PSB_SYM(AllReduce, "ncclAllReduce");
  PSB_SYM(Send, "ncclSend");
  PSB_SYM(Recv, "ncclRecv");
  PSB_SYM(GroupStart, "ncclGroupStart");
  PSB_SYM(GroupEnd, "ncclGroupEnd");
  PSB_SYM(GetErrorString, "ncclGetErrorString");
#undef PSB_SYM
  g_nccl.ok = true;
  return PSB_OK;
}

#define PSB_NCCL(expr)                                                           \
  do {                                                                           \
    ncclResult_t _r = (expr);                                                    \
    if (_r != ncclSuccess) {                                                     \
      psb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,               \
                     psb::g_nccl.GetErrorString(_r));                            \
      return PSB_ERR_NCCL;                                                       \
    }                                                                            \
  } while (0)

}  // namespace psb

struct psb_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_halo = nullptr;
};

struct psb_dist {
  psb_comm* comm = nullptr;
  psb_csr* A = nullptr;           // local block: n_loc rows, n_own + n_halo columns
  int64_t n_loc = 0, n_halo = 0;
  int64_t n_own = 0;              // owned entries of the INPUT vector (= n_loc for a square operator;
                                  // restriction / prolongation blocks are rectangular)
  int64_t r0 = 0, r1 = 0;         // rows [r0, r1) touch no halo column
  struct Peer { int rank; int64_t send_off, send_cnt, recv_off, recv_cnt; int32_t* d_send_idx; double* d_send_buf; };
  std::vector<Peer> peers;
  // ---- NVLink peer-memory mode (psb_dist_p2p_*) ------------------------------------------
  bool p2p = false;
  char* shm = nullptr;                 // this rank's exported region (cudaMalloc + IPC handle)
  size_t shm_bytes = 0;
  int64_t pbuf_off[2] = {0, 0};        // byte offsets of the two p buffers (n_loc + n_halo each)
  int64_t rbuf_off = 0;                // ... and of the r buffer (persistent-kernel path)
  std::vector<char*> peer_shm;         // [nranks] mapped base pointers (self = shm)
  unsigned long long** d_slot_ptrs = nullptr;   // device [kRing][nranks]: my slot in rank q, ring e
  int* d_error = nullptr;
  unsigned int red_epoch = 1;          // next reduction epoch (host-assigned, same on all ranks)
  unsigned long long halo_epoch = 1;   // epoch of the next p vector
  struct Push { int64_t send_off, cnt; double* remote[2]; double* remote_r; unsigned long long* remote_flag; };
  std::vector<Push> pushes;            // contiguous halo pushes (one per receiving peer)
};

namespace psb {

struct DistState {
  double udr[2];
  double loc[4];       // this rank's partial sums: [0..2] p.Ap parts (low boundary, interior,
                       // high boundary; an absent part stays 0), [3] r.r / b.b
  double red[4];       // the same, summed over ranks
  double norm_b, norm_r, tau;
  int k, maxiter, done, status, k_final, fail_on_maxiter, n_hist, pad;
};

__device__ __forceinline__ double2 ld2d(const double* p) {
  double2 v;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(kBlock)
dist_gather_kernel(const double* __restrict__ v, const int32_t* __restrict__ idx, int64_t cnt,
                   double* __restrict__ out, const int* d_skip) {
  if (d_skip != nullptr && ld_cg(d_skip) != 0) return;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * kBlock)
    out[i] = v[idx[i]];
}

// r = b, x = 0, p = b (local slices), local b.b -> red[3]
__global__ void __launch_bounds__(kBlock)
dist_init_kernel(DistState* st, int64_t n, const double* __restrict__ b, double* __restrict__ x,
                 double* __restrict__ r, double* __restrict__ p, ReduceBuf rb) {
  __shared__ double scratch[kWarps];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const double v = b[i];
    r[i] = v; x[i] = 0.0; p[i] = v;
    acc += v * v;
  }
  double t = block_sum(acc, scratch);
  if (threadIdx.x == 0) rb.partials[blockIdx.x] = t;
  if (last_block(rb.ticket)) {
    double s = sum_partials(rb.partials, gridDim.x, scratch);
    if (threadIdx.x == 0) st->loc[3] = s;
  }
}

// after the all-reduce of b.b
__global__ void dist_init_finish_kernel(DistState* st) {
  if (threadIdx.x != 0) return;
  const double bb = st->red[3];
  const double nb = sqrt(bb);
  st->norm_b = nb;
  st->udr[0] = bb;
  if (nb == 0.0) { st->done = 1; st->status = PSB_TRIVIAL; st->k_final = 0; st->norm_r = 0.0; }
}

// K2: x += alpha p ; r -= alpha Ap ; local r.r -> red[3].  `it` is the iteration index.
__global__ void __launch_bounds__(kBlock)
dist_update_kernel(DistState* st, int64_t n, int it, double* __restrict__ x, const double* __restrict__ p,
                   double* __restrict__ r, const double* __restrict__ Ap, ReduceBuf rb) {
  __shared__ double scratch[kWarps];
  if (ld_cg(&st->done) != 0) return;
  const double pAp = (ld_cg(&st->red[0]) + ld_cg(&st->red[1])) + ld_cg(&st->red[2]);
  if (pAp == 0.0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { st->status = PSB_BREAKDOWN_PAP; st->k_final = it; st->done = 1; }
    return;
  }
  const double alpha = ld_cg(&st->udr[it & 1]) / pAp;
  double acc = 0.0;
  const int64_t n2 = n >> 1;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kBlock) {
    double2 x0 = ld2d(x + 2 * i), p0 = ld2d(p + 2 * i), r0 = ld2d(r + 2 * i), a0 = ld_stream2(Ap + 2 * i);
    x0.x = x0.x + alpha * p0.x; x0.y = x0.y + alpha * p0.y;
    r0.x = r0.x - alpha * a0.x; r0.y = r0.y - alpha * a0.y;
    st_stream2(x + 2 * i, x0); st_stream2(r + 2 * i, r0);
    acc += r0.x * r0.x; acc += r0.y * r0.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t e = n - 1;
    const double xv = x[e] + alpha * p[e], rv = r[e] - alpha * Ap[e];
    x[e] = xv; r[e] = rv;
    acc += rv * rv;
  }
  double t = block_sum(acc, scratch);
  if (threadIdx.x == 0) rb.partials[blockIdx.x] = t;
  if (last_block(rb.ticket)) {
    double s = sum_partials(rb.partials, gridDim.x, scratch);
    if (threadIdx.x == 0) st->loc[3] = s;
  }
}

// K3 (after the all-reduce of r.r): convergence test, beta, p = r + beta p.  Every CTA
// takes the decision from the same scalars; CTA 0 records it.
__global__ void __launch_bounds__(kBlock)
dist_direction_kernel(DistState* st, int64_t n, int it, const double* __restrict__ r,
                      double* __restrict__ p, double* __restrict__ hist) {
  if (ld_cg(&st->done) != 0) return;
  const double rr = ld_cg(&st->red[3]);
  const double nr = sqrt(rr);
  const double tau = st->tau, nb = st->norm_b;
  const int maxiter = st->maxiter;
  const bool conv = (nr <= tau * nb) || (!st->fail_on_maxiter && it == maxiter - 1);
  const bool last = (it + 1 >= maxiter);
  const double rr_old = ld_cg(&st->udr[it & 1]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->norm_r = nr;
    hist[it] = nr;
    st->n_hist = it + 1;
    if (conv) { st->status = PSB_CONVERGED; st->k_final = it; st->done = 1; }
    else {
      st->udr[(it + 1) & 1] = rr;
      st->k = it + 1;
      if (last) { st->status = PSB_MAXITER; st->k_final = it; st->done = 1; }
    }
  }
  if (conv || last) return;
  const double beta = rr / rr_old;
  const int64_t n2 = n >> 1;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kBlock) {
    double2 z0 = ld_stream2(r + 2 * i), p0 = ld2d(p + 2 * i);
    p0.x = z0.x + beta * p0.x; p0.y = z0.y + beta * p0.y;
    st_stream2(p + 2 * i, p0);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) p[n - 1] = r[n - 1] + beta * p[n - 1];
}

// ---------------------------------------------------------------------------------------
// NVLink peer-memory mode: the collectives are fused into the compute kernels.
//   * scalar all-reduce: the last CTA of the producing kernel stores its partial, epoch-tagged,
//     into its slot in EVERY rank's memory (peer stores); the consuming kernel polls its local
//     slots and adds them in rank order -- same bits on every rank, no collective launch;
//   * halo: the kernel that writes p also stores the boundary slices straight into the
//     neighbours' halo (ping-pong p buffers), then raises their flag; the boundary-row SpMV of
//     the neighbour waits on that flag (spmv.cu wait_for_halo), interior rows never wait.
// Layout of the exported region: [slots: kRing x 32 ranks x 2 words][flags: 32 x u64]
// [p buffer 0][p buffer 1].
// ---------------------------------------------------------------------------------------
constexpr int64_t kSlotsBytes = (int64_t)kRing * kMaxRanks * kSlotWords * 8;    // a 128-byte line per slot
constexpr int64_t kFlagsBytes = 4096;

struct P2PView {
  const unsigned long long* my_slots;          // local: slot(e, q) at ((e % kRing) * kMaxRanks + q) * kSlotWords
  unsigned long long* const* slot_ptrs;        // device [kRing * nranks]: my slot in rank q's memory
  int nranks, my_rank;
  int n_push;
  int64_t push_off[kMaxPush], push_cnt[kMaxPush];
  double* push_remote[kMaxPush];
  unsigned long long* push_flag[kMaxPush];
  int* error;
};

// all threads call; returns sum over ranks of the value pushed for `epoch`
__device__ __forceinline__ double p2p_reduce(const P2PView& c, unsigned int epoch) {
  __shared__ double s_sum;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int q = 0; q < c.nranks; ++q) {
      double v;
      if (!peer_wait(c.my_slots + ((size_t)(epoch % kRing) * kMaxRanks + q) * kSlotWords, epoch, &v)) *c.error = 1;
      s += v;
    }
    s_sum = s;
  }
  __syncthreads();
  return s_sum;
}

__device__ __forceinline__ void p2p_push_scalar(const P2PView& c, unsigned int epoch, double v) {
  for (int q = 0; q < c.nranks; ++q)
    peer_push(c.slot_ptrs[(size_t)(epoch % kRing) * c.nranks + q], v, epoch);
}

// store p[i] into the halo of every neighbour whose slice contains i
__device__ __forceinline__ void p2p_push_halo(const P2PView& c, int64_t i, double v) {
#pragma unroll
  for (int k = 0; k < kMaxPush; ++k)
    if (k < c.n_push && i >= c.push_off[k] && i < c.push_off[k] + c.push_cnt[k])
      c.push_remote[k][i - c.push_off[k]] = v;
}

// after every CTA has issued its halo stores: the last CTA raises the neighbours' flags
__device__ __forceinline__ void p2p_raise_flags(const P2PView& c, unsigned int* ticket,
                                                unsigned long long halo_epoch) {
  __threadfence_system();
  if (last_block(ticket)) {
    if (threadIdx.x < c.n_push) {
      __threadfence_system();
      asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(c.push_flag[threadIdx.x]), "l"(halo_epoch) : "memory");
    }
  }
}

// r = b, x = 0, p = b (+ halo push), local b.b pushed for all-reduce
__global__ void __launch_bounds__(kBlock)
p2p_init_kernel(DistState* st, int64_t n, const double* __restrict__ b, double* __restrict__ x,
                double* __restrict__ r, double* __restrict__ p, ReduceBuf rb, P2PView c,
                unsigned int epoch, unsigned long long halo_epoch, unsigned int* ticket2) {
  __shared__ double scratch[kWarps];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const double v = b[i];
    r[i] = v; x[i] = 0.0; p[i] = v;
    p2p_push_halo(c, i, v);
    acc += v * v;
  }
  double t = block_sum(acc, scratch);
  if (threadIdx.x == 0) rb.partials[blockIdx.x] = t;
  if (last_block(rb.ticket)) {
    double s = sum_partials(rb.partials, gridDim.x, scratch);
    if (threadIdx.x == 0) { st->loc[3] = s; p2p_push_scalar(c, epoch, s); }
  }
  p2p_raise_flags(c, ticket2, halo_epoch);
}

__global__ void p2p_init_finish_kernel(DistState* st, P2PView c, unsigned int epoch) {
  const double bb = p2p_reduce(c, epoch);
  if (threadIdx.x != 0) return;
  const double nb = sqrt(bb);
  st->norm_b = nb;
  st->udr[0] = bb;
  if (nb == 0.0) { st->done = 1; st->status = PSB_TRIVIAL; st->k_final = 0; st->norm_r = 0.0; }
}

// K2: waits for the all-reduced p.Ap (epoch e_pap), updates x and r, pushes local r.r (e_rr)
__global__ void __launch_bounds__(kBlock)
p2p_update_kernel(DistState* st, int64_t n, int it, double* __restrict__ x, const double* __restrict__ p,
                  double* __restrict__ r, const double* __restrict__ Ap, ReduceBuf rb, P2PView c,
                  unsigned int e_pap, unsigned int e_rr) {
  __shared__ double scratch[kWarps];
  if (ld_cg(&st->done) != 0) return;
  const double pAp = p2p_reduce(c, e_pap);
  if (pAp == 0.0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { st->status = PSB_BREAKDOWN_PAP; st->k_final = it; st->done = 1; }
    return;
  }
  const double alpha = ld_cg(&st->udr[it & 1]) / pAp;
  double acc = 0.0;
  const int64_t n2 = n >> 1;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kBlock) {
    double2 x0 = ld2d(x + 2 * i), p0 = ld2d(p + 2 * i), r0 = ld2d(r + 2 * i), a0 = ld_stream2(Ap + 2 * i);
    x0.x = x0.x + alpha * p0.x; x0.y = x0.y + alpha * p0.y;
    r0.x = r0.x - alpha * a0.x; r0.y = r0.y - alpha * a0.y;
    st_stream2(x + 2 * i, x0); st_stream2(r + 2 * i, r0);
    acc += r0.x * r0.x; acc += r0.y * r0.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t e = n - 1;
    const double xv = x[e] + alpha * p[e], rv = r[e] - alpha * Ap[e];
    x[e] = xv; r[e] = rv;
    acc += rv * rv;
  }
  double t = block_sum(acc, scratch);
  if (threadIdx.x == 0) rb.partials[blockIdx.x] = t;
  if (last_block(rb.ticket)) {
    double s = sum_partials(rb.partials, gridDim.x, scratch);
    if (threadIdx.x == 0) { st->loc[3] = s; p2p_push_scalar(c, e_rr, s); }
  }
}

// K3: waits for the all-reduced r.r, convergence test, p_new = r + beta p_old written to the
// other p buffer AND to the neighbours' halos, then raises their flags.
__global__ void __launch_bounds__(kBlock)
p2p_direction_kernel(DistState* st, int64_t n, int it, const double* __restrict__ r,
                     const double* __restrict__ p_old, double* __restrict__ p_new,
                     double* __restrict__ hist, P2PView c, unsigned int e_rr,
                     unsigned long long halo_epoch, unsigned int* ticket2) {
  if (ld_cg(&st->done) != 0) return;
  const double rr = p2p_reduce(c, e_rr);
  const double nr = sqrt(rr);
  const int maxiter = st->maxiter;
  const bool conv = (nr <= st->tau * st->norm_b) || (!st->fail_on_maxiter && it == maxiter - 1);
  const bool last = (it + 1 >= maxiter);
  const double rr_old = ld_cg(&st->udr[it & 1]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->norm_r = nr;
    hist[it] = nr;
    st->n_hist = it + 1;
    if (conv) { st->status = PSB_CONVERGED; st->k_final = it; st->done = 1; }
    else {
      st->udr[(it + 1) & 1] = rr;
      st->k = it + 1;
      if (last) { st->status = PSB_MAXITER; st->k_final = it; st->done = 1; }
    }
  }
  if (conv || last) return;
  const double beta = rr / rr_old;
  const int64_t n2 = n >> 1;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kBlock) {
    const double2 z0 = ld_stream2(r + 2 * i), p0 = ld_stream2(p_old + 2 * i);
    double2 v;
    v.x = z0.x + beta * p0.x; v.y = z0.y + beta * p0.y;
    st_stream2(p_new + 2 * i, v);
    p2p_push_halo(c, 2 * i, v.x);
    p2p_push_halo(c, 2 * i + 1, v.y);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const double v = r[n - 1] + beta * p_old[n - 1];
    p_new[n - 1] = v;
    p2p_push_halo(c, n - 1, v);
  }
  p2p_raise_flags(c, ticket2, halo_epoch);
}

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct DistPoll {
  DistState* pinned = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int init() {
    if (pinned) return PSB_OK;
    PSB_CUDA(cudaHostAlloc((void**)&pinned, 2 * sizeof(DistState), cudaHostAllocDefault));
    PSB_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    PSB_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    return PSB_OK;
  }
};
static thread_local DistPoll t_dpoll;

// halo exchange of vector v (length n_loc + n_halo) on the comm stream
static int halo_exchange(psb_dist* D, double* v, const int* d_skip) {
  psb_comm* c = D->comm;
  for (auto& pr : D->peers) {
    if (pr.d_send_idx != nullptr && pr.send_cnt > 0) {
      int grid = (int)std::min<int64_t>((pr.send_cnt + kBlock - 1) / kBlock, (int64_t)sm_count() * 4);
      dist_gather_kernel<<<std::max(grid, 1), kBlock, 0, c->comm_stream>>>(v, pr.d_send_idx, pr.send_cnt,
                                                                          pr.d_send_buf, d_skip);
      PSB_LAUNCH_CHECK();
    }
  }
  PSB_NCCL(g_nccl.GroupStart());
  for (auto& pr : D->peers) {
    if (pr.send_cnt > 0) {
      const double* src = pr.d_send_idx ? pr.d_send_buf : v + pr.send_off;
      PSB_NCCL(g_nccl.Send(src, (size_t)pr.send_cnt, ncclDouble, pr.rank, c->comm, c->comm_stream));
    }
    if (pr.recv_cnt > 0)
      PSB_NCCL(g_nccl.Recv(v + D->n_own + pr.recv_off, (size_t)pr.recv_cnt, ncclDouble, pr.rank, c->comm,
                           c->comm_stream));
  }
  PSB_NCCL(g_nccl.GroupEnd());
  return PSB_OK;
}

}  // namespace psb

using namespace psb;

extern "C" int psb_nccl_unique_id(void* h_id128) {
  PSB_REQUIRE(h_id128 != nullptr, PSB_ERR_ARG, "psb_nccl_unique_id: NULL buffer");
  int rc = load_nccl();
  if (rc != PSB_OK) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  PSB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(h_id128, &id, sizeof(id));
  return PSB_OK;
}

extern "C" int psb_comm_create(const void* h_id128, int32_t rank, int32_t nranks, psb_comm_t* out) {
  PSB_REQUIRE(h_id128 && out && nranks >= 1 && rank >= 0 && rank < nranks, PSB_ERR_ARG,
              "psb_comm_create: bad argument");
  int rc = load_nccl();
  if (rc != PSB_OK) return rc;
  psb_comm* c = new (std::nothrow) psb_comm();
  PSB_REQUIRE(c != nullptr, PSB_ERR_ARG, "psb_comm_create: out of host memory");
  c->rank = rank; c->nranks = nranks;
  ncclUniqueId id;
  memcpy(&id, h_id128, sizeof(id));
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
  if (r != ncclSuccess) {
    set_error("psb_comm_create: ncclCommInitRank -> %s", g_nccl.GetErrorString(r));
    delete c;
    return PSB_ERR_NCCL;
  }
  PSB_CUDA(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
  PSB_CUDA(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
  PSB_CUDA(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
  *out = c;
  return PSB_OK;
}

extern "C" int psb_comm_destroy(psb_comm_t c) {
  if (!c) return PSB_OK;
  if (c->comm) g_nccl.CommDestroy(c->comm);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  if (c->ev_halo) cudaEventDestroy(c->ev_halo);
  delete c;
  return PSB_OK;
}

// Sum `count` doubles in place over all ranks (stream-ordered).
extern "C" int psb_comm_allreduce_sum(psb_comm_t c, double* d_buf, int64_t count, void* stream) {
  PSB_REQUIRE(c && d_buf && count >= 0, PSB_ERR_ARG, "psb_comm_allreduce_sum: bad argument");
  PSB_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)count, ncclDouble, ncclSum, c->comm, (cudaStream_t)stream));
  return PSB_OK;
}

extern "C" int psb_dist_create(psb_comm_t comm, psb_csr_t A_local, int64_t n_loc, int64_t n_halo,
                               int64_t r0, int64_t r1, int32_t n_peers, const int32_t* h_peer_rank,
                               const int64_t* h_send_off, const int64_t* h_send_cnt,
                               const int32_t* const* d_send_idx, const int64_t* h_recv_off,
                               const int64_t* h_recv_cnt, psb_dist_t* out) {
  PSB_REQUIRE(comm && A_local && out && n_loc >= 0 && n_halo >= 0 && n_peers >= 0, PSB_ERR_ARG,
              "psb_dist_create: bad argument");
  PSB_REQUIRE(A_local->n_rows == n_loc && A_local->n_cols >= n_halo, PSB_ERR_ARG,
              "psb_dist_create: local matrix must be n_loc x (n_own + n_halo)");
  PSB_REQUIRE(0 <= r0 && r0 <= r1 && r1 <= n_loc && (r0 % 4) == 0 && (r1 % 4 == 0 || r1 == n_loc),
              PSB_ERR_ARG, "psb_dist_create: interior range must be 4-aligned and inside [0, n_loc]");
  psb_dist* D = new (std::nothrow) psb_dist();
  PSB_REQUIRE(D != nullptr, PSB_ERR_ARG, "psb_dist_create: out of host memory");
  D->comm = comm; D->A = A_local; D->n_loc = n_loc; D->n_halo = n_halo; D->r0 = r0; D->r1 = r1;
  D->n_own = A_local->n_cols - n_halo;
  for (int i = 0; i < n_peers; ++i) {
    psb_dist::Peer p;
    p.rank = h_peer_rank[i];
    p.send_off = h_send_off[i]; p.send_cnt = h_send_cnt[i];
    p.recv_off = h_recv_off[i]; p.recv_cnt = h_recv_cnt[i];
    p.d_send_idx = nullptr; p.d_send_buf = nullptr;
    if (d_send_idx != nullptr && d_send_idx[i] != nullptr && p.send_cnt > 0) {
      p.d_send_idx = const_cast<int32_t*>(d_send_idx[i]);
      cudaError_t e = cudaMalloc((void**)&p.d_send_buf, (size_t)p.send_cnt * sizeof(double));
      if (e != cudaSuccess) { set_error("psb_dist_create: %s", cudaGetErrorString(e)); delete D; return PSB_ERR_CUDA; }
    }
    D->peers.push_back(p);
  }
  *out = D;
  return PSB_OK;
}

extern "C" int psb_dist_destroy(psb_dist_t D) {
  if (!D) return PSB_OK;
  for (auto& p : D->peers) if (p.d_send_buf) cudaFree(p.d_send_buf);
  for (size_t q = 0; q < D->peer_shm.size(); ++q)
    if (D->peer_shm[q] && D->peer_shm[q] != D->shm) cudaIpcCloseMemHandle(D->peer_shm[q]);
  if (D->shm) cudaFree(D->shm);
  if (D->d_slot_ptrs) cudaFree(D->d_slot_ptrs);
  if (D->d_error) cudaFree(D->d_error);
  delete D;
  return PSB_OK;
}

// ---- NVLink peer-memory mode: export / map the shared regions ---------------------------------
extern "C" int psb_dist_p2p_alloc(psb_dist_t D, void* h_handle64, int64_t layout[6]) {
  PSB_REQUIRE(D && h_handle64 && layout, PSB_ERR_ARG, "psb_dist_p2p_alloc: NULL argument");
  PSB_REQUIRE(D->comm->nranks <= kMaxRanks, PSB_ERR_UNSUPP, "psb_dist_p2p_alloc: too many ranks");
  PSB_REQUIRE(D->n_own == D->n_loc, PSB_ERR_UNSUPP, "psb_dist_p2p_alloc: the peer-memory PCG path needs a square operator");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  const int64_t vec = align_up((D->n_loc + D->n_halo) * 8, 256);
  D->pbuf_off[0] = kSlotsBytes + kFlagsBytes;
  D->pbuf_off[1] = D->pbuf_off[0] + vec;
  D->rbuf_off = D->pbuf_off[1] + vec;
  D->shm_bytes = (size_t)(D->rbuf_off + vec);
  PSB_CUDA(cudaMalloc((void**)&D->shm, D->shm_bytes));
  PSB_CUDA(cudaMemset(D->shm, 0, D->shm_bytes));
  PSB_CUDA(cudaMalloc((void**)&D->d_error, sizeof(int)));
  PSB_CUDA(cudaMemset(D->d_error, 0, sizeof(int)));
  cudaIpcMemHandle_t h;
  PSB_CUDA(cudaIpcGetMemHandle(&h, D->shm));
  memcpy(h_handle64, &h, sizeof(h));
  layout[0] = D->pbuf_off[0]; layout[1] = D->pbuf_off[1]; layout[2] = D->rbuf_off;
  layout[3] = D->n_loc; layout[4] = D->n_halo; layout[5] = 0;
  return PSB_OK;
}

extern "C" int psb_dist_p2p_open(psb_dist_t D, const void* h_handles, int32_t n_push,
                                 const int32_t* h_push_rank, const int64_t* h_send_off,
                                 const int64_t* h_send_cnt, const int64_t* h_remote_off0,
                                 const int64_t* h_remote_off1, const int64_t* h_remote_off_r,
                                 const int32_t* h_remote_flag_index) {
  PSB_REQUIRE(D && h_handles && D->shm, PSB_ERR_ARG, "psb_dist_p2p_open: call psb_dist_p2p_alloc first");
  PSB_REQUIRE(n_push >= 0 && n_push <= kMaxPush, PSB_ERR_UNSUPP, "psb_dist_p2p_open: too many halo targets");
  PSB_REQUIRE(D->A->kind == PSB_SPMV_STREAM, PSB_ERR_UNSUPP,
              "psb_dist_p2p_open: the peer-memory mode needs the STREAM SpMV kernel");
  const int nr = D->comm->nranks, me = D->comm->rank;
  D->peer_shm.assign(nr, nullptr);
  for (int q = 0; q < nr; ++q) {
    if (q == me) { D->peer_shm[q] = D->shm; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)h_handles + (size_t)q * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    PSB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    D->peer_shm[q] = (char*)ptr;
  }
  // my slot (ring e) in rank q's memory
  std::vector<unsigned long long*> sp((size_t)kRing * nr);
  for (int e = 0; e < kRing; ++e)
    for (int q = 0; q < nr; ++q)
      sp[(size_t)e * nr + q] = (unsigned long long*)D->peer_shm[q] + ((size_t)e * kMaxRanks + me) * kSlotWords;
  PSB_CUDA(cudaMalloc((void**)&D->d_slot_ptrs, sp.size() * sizeof(void*)));
  PSB_CUDA(cudaMemcpy(D->d_slot_ptrs, sp.data(), sp.size() * sizeof(void*), cudaMemcpyHostToDevice));
  D->pushes.clear();
  for (int i = 0; i < n_push; ++i) {
    psb_dist::Push P;
    const int q = h_push_rank[i];
    PSB_REQUIRE(q >= 0 && q < nr && q != me, PSB_ERR_ARG, "psb_dist_p2p_open: bad push rank");
    P.send_off = h_send_off[i]; P.cnt = h_send_cnt[i];
    P.remote[0] = (double*)(D->peer_shm[q] + h_remote_off0[i]);
    P.remote[1] = (double*)(D->peer_shm[q] + h_remote_off1[i]);
    P.remote_r = (double*)(D->peer_shm[q] + h_remote_off_r[i]);
    P.remote_flag = (unsigned long long*)(D->peer_shm[q] + kSlotsBytes) + h_remote_flag_index[i];
    D->pushes.push_back(P);
  }
  D->p2p = true;
  return PSB_OK;
}

namespace psb {

// y_loc = A_loc [x_own | halo(x)] with any SpMV epilogue: exchanges the halo of d_x_ext (length
// n_own + n_halo) and multiplies; interior rows overlap the exchange.  Reductions: EPI_DOT writes
// three partial sums to ea.dot[0..2] (boundary-low, interior, boundary-high; absent parts stay
// untouched: zero them first); EPI_RESID_NORM accumulates ONE local sum in *ea.dot.
int dist_spmv_epi(psb_dist* D, Epi epi, double* d_x_ext, double* d_y, const EpiArgs& ea0, const int* d_skip,
                  cudaStream_t st) {
  psb_comm* c = D->comm;
  const bool comm_needed = !D->peers.empty();
  if (comm_needed) {
    PSB_CUDA(cudaEventRecord(c->ev_ready, st));
    PSB_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_ready, 0));
    int rc = halo_exchange(D, d_x_ext, d_skip);
    if (rc != PSB_OK) return rc;
    PSB_CUDA(cudaEventRecord(c->ev_halo, c->comm_stream));
  }
  EpiArgs ea = ea0;
  double* const dot = ea0.dot;
  int launched = 0;
  int rc;
  auto part = [&](int64_t a, int64_t b, int which) -> int {
    psb_csr v = csr_row_view(D->A, a, b);
    if (epi == EPI_DOT && dot) ea.dot = dot + which;
    if (epi == EPI_RESID_NORM) ea.dot_accumulate = launched > 0 ? 1 : 0;
    ++launched;
    return spmv_launch(&v, epi, d_x_ext, d_y, ea, d_skip, st);
  };
  if (D->r1 > D->r0) {                                   // interior rows: no halo column
    rc = part(D->r0, D->r1, 1);
    if (rc != PSB_OK) return rc;
  }
  if (comm_needed) PSB_CUDA(cudaStreamWaitEvent(st, c->ev_halo, 0));
  if (D->r0 > 0) {
    rc = part(0, D->r0, 0);
    if (rc != PSB_OK) return rc;
  }
  if (D->r1 < D->n_loc) {
    rc = part(D->r1, D->n_loc, 2);
    if (rc != PSB_OK) return rc;
  }
  return PSB_OK;
}

static int dist_spmv(psb_dist* D, double* d_x_ext, double* d_y, double* d_dot3, const int* d_skip,
                     cudaStream_t st) {
  EpiArgs ea;
  ea.dot = d_dot3;
  return dist_spmv_epi(D, d_dot3 ? EPI_DOT : EPI_STORE, d_x_ext, d_y, ea, d_skip, st);
}

// Sum `count` doubles in place over the ranks of D's communicator (stream-ordered).
int dist_allreduce(psb_comm* c, double* d_buf, int count, cudaStream_t st) {
  PSB_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)count, ncclDouble, ncclSum, c->comm, st));
  return PSB_OK;
}

// Every rank contributes d_full[starts[rank] .. starts[rank+1]) and receives the other slices
// (grouped send / recv: the slices need not have equal lengths).  Stream-ordered on st.
int dist_allgather_slices(psb_comm* c, double* d_full, const int64_t* starts, cudaStream_t st) {
  const int me = c->rank;
  PSB_NCCL(g_nccl.GroupStart());
  for (int q = 0; q < c->nranks; ++q) {
    if (q == me) continue;
    const int64_t mine = starts[me + 1] - starts[me], theirs = starts[q + 1] - starts[q];
    if (mine > 0) PSB_NCCL(g_nccl.Send(d_full + starts[me], (size_t)mine, ncclDouble, q, c->comm, st));
    if (theirs > 0) PSB_NCCL(g_nccl.Recv(d_full + starts[q], (size_t)theirs, ncclDouble, q, c->comm, st));
  }
  PSB_NCCL(g_nccl.GroupEnd());
  return PSB_OK;
}

int dist_rank(const psb_comm* c) { return c->rank; }
int dist_nranks(const psb_comm* c) { return c->nranks; }
int64_t dist_n_loc(const psb_dist* D) { return D->n_loc; }
int64_t dist_n_own(const psb_dist* D) { return D->n_own; }
int64_t dist_n_halo(const psb_dist* D) { return D->n_halo; }
psb_comm* dist_comm(const psb_dist* D) { return D->comm; }

}  // namespace psb

using namespace psb;

extern "C" int psb_dist_spmv(psb_dist_t D, double* d_x_ext, double* d_y, void* stream) {
  PSB_REQUIRE(D && d_x_ext && d_y, PSB_ERR_ARG, "psb_dist_spmv: NULL argument");
  return dist_spmv(D, d_x_ext, d_y, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int64_t psb_dist_pcg_workspace_bytes(int64_t n_loc, int64_t n_halo) {
  if (n_loc < 0 || n_halo < 0) return PSB_ERR_ARG;
  const int64_t hdr = 4096 + align_up((int64_t)sm_count() * 16 * sizeof(double), 256);
  return hdr + 2 * align_up(n_loc * 8, 256) + align_up((n_loc + n_halo) * 8, 256);
}

static int dist_pcg_p2p(psb_dist_t D, const double* d_b, double* d_x, void* d_work, int32_t maxiter,
                        double tau, int32_t fail_on_maxiter, double* d_hist, psb_solve_result* result,
                        cudaStream_t st);

extern "C" int psb_dist_pcg_solve(psb_dist_t D, const double* d_b, double* d_x, void* d_work,
                                  int64_t work_bytes, int32_t maxiter, double tau,
                                  int32_t fail_on_maxiter, double* d_hist, psb_solve_result* result,
                                  void* stream) {
  PSB_REQUIRE(D && d_b && d_x && d_work && d_hist && result, PSB_ERR_ARG, "psb_dist_pcg_solve: NULL argument");
  if (D->p2p) {
    PSB_REQUIRE(maxiter >= 1, PSB_ERR_ARG, "psb_dist_pcg_solve: maxiter must be >= 1");
    PSB_REQUIRE(work_bytes >= psb_dist_pcg_workspace_bytes(D->n_loc, D->n_halo), PSB_ERR_ARG,
                "psb_dist_pcg_solve: workspace too small");
    PSB_REQUIRE(aligned16(d_b) && aligned16(d_x) && ((uintptr_t)d_work & 255u) == 0, PSB_ERR_ARG,
                "psb_dist_pcg_solve: b, x must be 16-byte and work 256-byte aligned");
    return dist_pcg_p2p(D, d_b, d_x, d_work, maxiter, tau, fail_on_maxiter, d_hist, result,
                        (cudaStream_t)stream);
  }
  PSB_REQUIRE(maxiter >= 1, PSB_ERR_ARG, "psb_dist_pcg_solve: maxiter must be >= 1");
  const int64_t n = D->n_loc;
  PSB_REQUIRE(work_bytes >= psb_dist_pcg_workspace_bytes(n, D->n_halo), PSB_ERR_ARG,
              "psb_dist_pcg_solve: workspace too small");
  PSB_REQUIRE(aligned16(d_b) && aligned16(d_x) && ((uintptr_t)d_work & 255u) == 0, PSB_ERR_ARG,
              "psb_dist_pcg_solve: b, x must be 16-byte and work 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  psb_comm* c = D->comm;
  int rc = t_dpoll.init();
  if (rc != PSB_OK) return rc;

  char* base = (char*)d_work;
  DistState* S = (DistState*)base;
  ReduceBuf rb;
  rb.ticket = (unsigned int*)(base + 1024);
  rb.partials = (double*)(base + 4096);
  rb.max_grid = sm_count() * 16;
  char* v = base + 4096 + align_up((int64_t)sm_count() * 16 * sizeof(double), 256);
  const int64_t vec = align_up(n * 8, 256);
  double* r = (double*)v;
  double* Ap = (double*)(v + vec);
  double* p = (double*)(v + 2 * vec);          // n_loc + n_halo

  PSB_CUDA(cudaMemsetAsync(d_work, 0, 4096, st));
  DistState h0;
  memset(&h0, 0, sizeof(h0));
  h0.tau = tau; h0.maxiter = maxiter; h0.fail_on_maxiter = fail_on_maxiter;
  PSB_CUDA(cudaMemcpyAsync(S, &h0, sizeof(h0), cudaMemcpyHostToDevice, st));
  PSB_CUDA(cudaStreamSynchronize(st));

  const int grid = stream_grid(std::max<int64_t>(n, 1), rb.max_grid);
  dist_init_kernel<<<grid, kBlock, 0, st>>>(S, n, d_b, d_x, r, p, rb);
  PSB_LAUNCH_CHECK();
  PSB_NCCL(g_nccl.AllReduce(&S->loc[3], &S->red[3], 1, ncclDouble, ncclSum, c->comm, st));
  dist_init_finish_kernel<<<1, 32, 0, st>>>(S);
  PSB_LAUNCH_CHECK();

  const int chunk = 16;
  int enq = 0, slot = 0;
  bool pending[2] = {false, false};
  bool finished = false;
  while (!finished) {
    const int todo = std::min(chunk, maxiter - enq);
    for (int i = 0; i < todo; ++i) {
      const int it = enq + i;
      rc = dist_spmv(D, p, Ap, &S->loc[0], &S->done, st);
      if (rc != PSB_OK) return rc;
      PSB_NCCL(g_nccl.AllReduce(&S->loc[0], &S->red[0], 3, ncclDouble, ncclSum, c->comm, st));
      dist_update_kernel<<<grid, kBlock, 0, st>>>(S, n, it, d_x, p, r, Ap, rb);
      PSB_LAUNCH_CHECK();
      PSB_NCCL(g_nccl.AllReduce(&S->loc[3], &S->red[3], 1, ncclDouble, ncclSum, c->comm, st));
      dist_direction_kernel<<<grid, kBlock, 0, st>>>(S, n, it, r, p, d_hist);
      PSB_LAUNCH_CHECK();
    }
    enq += todo;
    PSB_CUDA(cudaMemcpyAsync(&t_dpoll.pinned[slot], S, sizeof(DistState), cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaEventRecord(t_dpoll.ev[slot], st));
    pending[slot] = true;
    const int prev = slot ^ 1;
    if (pending[prev]) {            // same logical point on every rank -> same decision
      PSB_CUDA(cudaEventSynchronize(t_dpoll.ev[prev]));
      pending[prev] = false;
      if (t_dpoll.pinned[prev].done) finished = true;
    }
    if (enq >= maxiter) finished = true;
    slot ^= 1;
  }
  PSB_CUDA(cudaStreamSynchronize(st));
  PSB_CUDA(cudaStreamSynchronize(c->comm_stream));
  DistState hs;
  PSB_CUDA(cudaMemcpy(&hs, S, sizeof(hs), cudaMemcpyDeviceToHost));
  if (!hs.done) {
    set_error("psb_dist_pcg_solve: device loop ended without a terminal state (k=%d)", hs.k);
    return PSB_ERR_CUDA;
  }
  result->status = hs.status; result->k = hs.k_final; result->n_hist = hs.n_hist; result->lucky = 0;
  result->norm_r = hs.norm_r; result->norm_b = hs.norm_b; result->norm_r_rec = hs.norm_r;
  return PSB_OK;
}


static int dist_pcg_p2p(psb_dist_t D, const double* d_b, double* d_x, void* d_work, int32_t maxiter,
                        double tau, int32_t fail_on_maxiter, double* d_hist, psb_solve_result* result,
                        cudaStream_t st) {
  const int64_t n = D->n_loc;
  int rc = t_dpoll.init();
  if (rc != PSB_OK) return rc;
  char* base = (char*)d_work;
  DistState* S = (DistState*)base;
  ReduceBuf rb;
  rb.ticket = (unsigned int*)(base + 1024);
  unsigned int* ticket2 = (unsigned int*)(base + 1024 + 64);
  rb.partials = (double*)(base + 4096);
  rb.max_grid = sm_count() * 16;
  char* v = base + 4096 + align_up((int64_t)sm_count() * 16 * sizeof(double), 256);
  const int64_t vec = align_up(n * 8, 256);
  double* r = (double*)v;
  double* Ap = (double*)(v + vec);
  double* pbuf[2] = {(double*)(D->shm + D->pbuf_off[0]), (double*)(D->shm + D->pbuf_off[1])};

  PSB_CUDA(cudaMemsetAsync(d_work, 0, 4096, st));
  DistState h0;
  memset(&h0, 0, sizeof(h0));
  h0.tau = tau; h0.maxiter = maxiter; h0.fail_on_maxiter = fail_on_maxiter;
  PSB_CUDA(cudaMemcpyAsync(S, &h0, sizeof(h0), cudaMemcpyHostToDevice, st));
  PSB_CUDA(cudaStreamSynchronize(st));

  P2PView c;
  memset(&c, 0, sizeof(c));
  c.my_slots = (const unsigned long long*)D->shm;
  c.slot_ptrs = D->d_slot_ptrs;
  c.nranks = D->comm->nranks; c.my_rank = D->comm->rank;
  c.n_push = (int)D->pushes.size();
  c.error = D->d_error;
  for (int k = 0; k < c.n_push; ++k) {
    c.push_off[k] = D->pushes[k].send_off; c.push_cnt[k] = D->pushes[k].cnt;
    c.push_flag[k] = D->pushes[k].remote_flag;
  }
  auto view_for = [&](unsigned long long h) {
    P2PView w = c;
    for (int k = 0; k < w.n_push; ++k) w.push_remote[k] = D->pushes[k].remote[h & 1];
    return w;
  };
  unsigned int e = D->red_epoch;
  const unsigned long long h0e = D->halo_epoch;
  const unsigned long long* my_flags = (const unsigned long long*)(D->shm + kSlotsBytes);
  int n_wait = 0;
  {
    std::vector<int> owners;
    for (auto& pr : D->peers) if (pr.recv_cnt > 0) owners.push_back(pr.rank);
    n_wait = (int)owners.size();
  }

  const int grid = stream_grid(std::max<int64_t>(n, 1), rb.max_grid);
  {
    const char* env = getenv("PSB_DIST_MEGA");
    const bool mega_ok = D->A->kind == PSB_SPMV_STREAM && D->A->rpt == 1 && n > 0 && !(env && env[0] == '0');
    if (mega_ok) {
      // the whole solve as ONE persistent kernel per GPU; halo + all-reduces over NVLink peer memory
      MegaParams P;
      memset(&P, 0, sizeof(P));
      P.n = n; P.n_halo = D->n_halo;
      P.b = d_b; P.x = d_x; P.Ap = Ap;
      P.r = (double*)(D->shm + D->rbuf_off);
      P.pbuf[0] = pbuf[0]; P.pbuf[1] = pbuf[1];
      // ping-pong parity: iteration `it` writes pbuf[it & 1]; halo flags count from h0e
      P.hist = d_hist;
      P.st = (MegaState*)(base + 2048);
      P.ticket = rb.ticket; P.partials = rb.partials;
      P.my_slots = (const unsigned long long*)D->shm; P.slot_ptrs = D->d_slot_ptrs;
      P.nranks = c.nranks; P.epoch0 = e; P.ring_words = kMaxRanks * kSlotWords;
      P.n_push = c.n_push;
      for (int k = 0; k < c.n_push; ++k) {
        P.push_off[k] = D->pushes[k].send_off; P.push_cnt[k] = D->pushes[k].cnt;
        P.push_r[k] = D->pushes[k].remote_r;
        P.push_p[0][k] = D->pushes[k].remote[0]; P.push_p[1][k] = D->pushes[k].remote[1];
        P.push_flag[k] = D->pushes[k].remote_flag;
      }
      P.n_wait = n_wait; P.my_flags = my_flags; P.halo_epoch0 = h0e;
      P.int_r0 = D->r0; P.int_r1 = D->r1;
      P.maxiter = maxiter; P.tau = tau; P.fail_on_maxiter = fail_on_maxiter;
      P.error = D->d_error;
      rc = pcg_mega_launch(P, D->A, st);
      if (rc != PSB_OK) return rc;
      PSB_CUDA(cudaStreamSynchronize(st));
      MegaState ms;
      PSB_CUDA(cudaMemcpy(&ms, P.st, sizeof(ms), cudaMemcpyDeviceToHost));
      int err = 0;
      PSB_CUDA(cudaMemcpy(&err, D->d_error, sizeof(int), cudaMemcpyDeviceToHost));
      D->red_epoch = e + ms.epochs_used;
      D->halo_epoch = h0e + ms.halo_epochs_used;
      if (err) { set_error("psb_dist_pcg_solve: timed out waiting for a peer GPU (rank %d)", D->comm->rank); return PSB_ERR_NCCL; }
      if (!ms.done) { set_error("psb_dist_pcg_solve: persistent kernel ended without a terminal state"); return PSB_ERR_CUDA; }
      result->status = ms.status; result->k = ms.k_final; result->n_hist = ms.n_hist; result->lucky = 0;
      result->norm_r = ms.norm_r; result->norm_b = ms.norm_b; result->norm_r_rec = ms.norm_r;
      return PSB_OK;
    }
  }

  {
    const unsigned int e_bb = e++;
    p2p_init_kernel<<<grid, kBlock, 0, st>>>(S, n, d_b, d_x, r, pbuf[h0e & 1], rb, view_for(h0e), e_bb, h0e, ticket2);
    PSB_LAUNCH_CHECK();
    p2p_init_finish_kernel<<<1, 32, 0, st>>>(S, c, e_bb);
    PSB_LAUNCH_CHECK();
  }

  // the three row ranges of the SpMV; the last one launched carries the peer push
  struct Part { int64_t a, b; bool waits; };
  std::vector<Part> parts;
  if (D->r1 > D->r0) parts.push_back({D->r0, D->r1, false});
  if (D->r0 > 0) parts.push_back({0, D->r0, true});
  if (D->r1 < n) parts.push_back({D->r1, n, true});

  const bool one_launch = D->A->rpt == 1 && (D->r0 % 256) == 0 && ((D->r1 % 256) == 0 || D->r1 == n);

  const int chunk = 16;
  int enq = 0, slot = 0;
  bool pending[2] = {false, false};
  bool finished = false;
  while (!finished) {
    const int todo = std::min(chunk, maxiter - enq);
    for (int i = 0; i < todo; ++i) {
      const int it = enq + i;
      const unsigned long long h = h0e + (unsigned long long)it;
      double* pcur = pbuf[h & 1];
      double* pnext = pbuf[(h + 1) & 1];
      const unsigned int e_pap = e++, e_rr = e++;
      if (one_launch) {
        // one SpMV launch: interior tiles first, a CTA waits for the halo flags only when it
        // reaches its first boundary tile; the last CTA pushes p.Ap to every rank
        EpiArgs ea;
        ea.dot = &S->loc[0];
        ea.error_flag = D->d_error;
        if (n_wait > 0) { ea.wait_flags = my_flags; ea.wait_n = n_wait; ea.wait_value = h; }
        ea.rot_t0 = D->r0 / 256; ea.rot_t1 = (D->r1 + 255) / 256;
        ea.push_slots = D->d_slot_ptrs + (size_t)(e_pap % kRing) * c.nranks;
        ea.push_n = c.nranks;
        ea.push_epoch = e_pap;
        rc = spmv_launch(D->A, EPI_DOT, pcur, Ap, ea, &S->done, st);
        if (rc != PSB_OK) return rc;
      } else {
        for (size_t k = 0; k < parts.size(); ++k) {
          psb_csr vw = csr_row_view(D->A, parts[k].a, parts[k].b);
          EpiArgs ea;
          ea.dot = &S->loc[0];
          ea.dot_accumulate = k > 0 ? 1 : 0;
          ea.error_flag = D->d_error;
          if (parts[k].waits && n_wait > 0) { ea.wait_flags = my_flags; ea.wait_n = n_wait; ea.wait_value = h; }
          if (k + 1 == parts.size()) {
            ea.push_slots = D->d_slot_ptrs + (size_t)(e_pap % kRing) * c.nranks;
            ea.push_n = c.nranks;
            ea.push_epoch = e_pap;
          }
          rc = spmv_launch(&vw, EPI_DOT, pcur, Ap, ea, &S->done, st);
          if (rc != PSB_OK) return rc;
        }
      }
      p2p_update_kernel<<<grid, kBlock, 0, st>>>(S, n, it, d_x, pcur, r, Ap, rb, c, e_pap, e_rr);
      PSB_LAUNCH_CHECK();
      p2p_direction_kernel<<<grid, kBlock, 0, st>>>(S, n, it, r, pcur, pnext, d_hist, view_for(h + 1), e_rr,
                                                    h + 1, ticket2);
      PSB_LAUNCH_CHECK();
    }
    enq += todo;
    PSB_CUDA(cudaMemcpyAsync(&t_dpoll.pinned[slot], S, sizeof(DistState), cudaMemcpyDeviceToHost, st));
    PSB_CUDA(cudaEventRecord(t_dpoll.ev[slot], st));
    pending[slot] = true;
    const int prev = slot ^ 1;
    if (pending[prev]) {
      PSB_CUDA(cudaEventSynchronize(t_dpoll.ev[prev]));
      pending[prev] = false;
      if (t_dpoll.pinned[prev].done) finished = true;
    }
    if (enq >= maxiter) finished = true;
    slot ^= 1;
  }
  D->red_epoch = e;
  D->halo_epoch = h0e + (unsigned long long)enq + 1;
  PSB_CUDA(cudaStreamSynchronize(st));
  DistState hs;
  PSB_CUDA(cudaMemcpy(&hs, S, sizeof(hs), cudaMemcpyDeviceToHost));
  int err = 0;
  PSB_CUDA(cudaMemcpy(&err, D->d_error, sizeof(int), cudaMemcpyDeviceToHost));
  if (err) {
    set_error("psb_dist_pcg_solve: timed out waiting for a peer GPU (rank %d)", D->comm->rank);
    return PSB_ERR_NCCL;
  }
  if (!hs.done) {
    set_error("psb_dist_pcg_solve: device loop ended without a terminal state (k=%d)", hs.k);
    return PSB_ERR_CUDA;
  }
  result->status = hs.status; result->k = hs.k_final; result->n_hist = hs.n_hist; result->lucky = 0;
  result->norm_r = hs.norm_r; result->norm_b = hs.norm_b; result->norm_r_rec = hs.norm_r;
  return PSB_OK;
}
