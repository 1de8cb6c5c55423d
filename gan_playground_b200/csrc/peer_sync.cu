// One-shot all-reduce of small fp32 vectors over NVLink peer memory, fused with the BatchNorm statistics finalize
// (SyncBN: SURVEY.md §8e — 15 forward + 12 backward reductions of <= 8 KB per DCGAN-64 step, each far below NCCL's
// launch + protocol latency).
//
// Every rank owns one symmetric buffer (mapped into all peers) of 2 parities x 8 source ranks x (flag line + 4096
// floats). A call with epoch e (a device-resident counter that advances identically on every rank, so launches can
// sit in CUDA graphs) uses parity e & 1 and sequence number e + 1:
//   push : each rank stores its n floats into slot [parity][own rank] of EVERY peer's buffer (plain stores over NVLink),
//          fences (system scope), then writes the sequence number into the slot's flag with a release store;
//   wait : spins (acquire loads) on the world flags of its OWN buffer until they show the sequence number;
//   sum  : adds the world slots in rank order — the same order on every rank, so all ranks get bit-identical totals.
// Two parities are enough: a rank can be at most one call ahead of the slowest one (it cannot finish call k+1 before
// every rank has pushed k+1, i.e. finished reading call k).
#include "common.h"

namespace gp {

constexpr int kPeerMaxWorld = 8;
constexpr int kPeerMaxFloats = 4096;
constexpr int kPeerFlagFloats = 32;  // one 128-byte line per flag
constexpr int kPeerSlotFloats = kPeerFlagFloats + kPeerMaxFloats;

struct PeerCtx {
  float* bufs[kPeerMaxWorld];
  int world, rank;
  unsigned* epoch;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long peer_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Block-wide: src (n floats, global or shared) summed over ranks into s_tot (shared, n floats). One block per call.
__device__ void peer_exchange_sum(const PeerCtx& px, const float* src, int n, float* s_tot) {
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0) s_epoch = *px.epoch;
  __syncthreads();
  const unsigned e = s_epoch, seq = e + 1u;
  const long long slot_mine = (long long)((e & 1u) * kPeerMaxWorld + px.rank) * kPeerSlotFloats;
  for (int r = 0; r < px.world; ++r) {
    float* dst = px.bufs[r] + slot_mine + kPeerFlagFloats;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < px.world)
    st_release_sys(reinterpret_cast<unsigned*>(px.bufs[threadIdx.x] + slot_mine), seq);
  if ((int)threadIdx.x < px.world) {
    const unsigned* flag = reinterpret_cast<const unsigned*>(
        px.bufs[px.rank] + (long long)((e & 1u) * kPeerMaxWorld + threadIdx.x) * kPeerSlotFloats);
    if (ld_acquire_sys(flag) != seq) {
      const unsigned long long t0 = peer_timer_ns();
      unsigned spins = 0;
      while (ld_acquire_sys(flag) != seq) {
        // a missing peer must not hang the GPU: give up after 20 s and fail the launch
        if ((++spins & 255u) == 0 && peer_timer_ns() - t0 > 20000000000ull) asm volatile("trap;");
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < px.world; ++s)
      acc += __ldcv(px.bufs[px.rank] + (long long)((e & 1u) * kPeerMaxWorld + s) * kPeerSlotFloats + kPeerFlagFloats + i);
    s_tot[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) *px.epoch = e + 1u;
}

__global__ void __launch_bounds__(512) peer_allreduce_kernel(const PeerCtx px, float* data, int n) {
  gp::pdl_sync();
  __shared__ float s_tot[kPeerMaxFloats];
  peer_exchange_sum(px, data, n, s_tot);
  for (int i = threadIdx.x; i < n; i += blockDim.x) data[i] = s_tot[i];
}

// exchange of (sum, sum of squares) + the finalize of elementwise.cu's bn_finalize_kernel, in one launch
__global__ void __launch_bounds__(512) bn_finalize_peer_kernel(const PeerCtx px, float* st, double count, int C, float eps,
                                                               float momentum, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float* __restrict__ mean,
                                                               float* __restrict__ rstd, float* __restrict__ scale,
                                                               float* __restrict__ shift, float* running_mean,
                                                               float* running_var, long long* num_batches_tracked) {
  gp::pdl_sync();
  __shared__ float s_tot[kPeerMaxFloats];
  peer_exchange_sum(px, st, 2 * C, s_tot);
  if (threadIdx.x == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    st[c] = s_tot[c];  // the caller's buffer ends up holding the global sums, as after an in-place all-reduce
    st[C + c] = s_tot[C + c];
    const double m = (double)s_tot[c] / count;
    double var = (double)s_tot[C + c] / count - m * m;
    if (var < 0) var = 0;
    const float r = rsqrtf((float)var + eps);
    mean[c] = (float)m;
    rstd[c] = r;
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    scale[c] = g * r;
    shift[c] = b - (float)m * g * r;
    if (running_mean != nullptr) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      const double unbiased = count > 1 ? var * count / (count - 1) : var;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
}

static int make_ctx(const gp_peer_t* p, PeerCtx* px) {
  if (p == nullptr || p->world < 1 || p->world > kPeerMaxWorld || p->rank < 0 || p->rank >= p->world || p->epoch == nullptr)
    return set_error(GP_ERR_INVALID, "peer sync: bad context (world must be 1..%d)", kPeerMaxWorld);
  for (int r = 0; r < kPeerMaxWorld; ++r) {
    px->bufs[r] = r < p->world ? static_cast<float*>(p->bufs[r]) : nullptr;
    if (r < p->world && px->bufs[r] == nullptr) return set_error(GP_ERR_INVALID, "peer sync: null peer buffer %d", r);
  }
  px->world = p->world;
  px->rank = p->rank;
  px->epoch = static_cast<unsigned*>(p->epoch);
  return GP_OK;
}

}  // namespace gp

using namespace gp;

extern "C" {

long long gp_peer_buffer_bytes(void) { return 2LL * kPeerMaxWorld * kPeerSlotFloats * (long long)sizeof(float); }

int gp_peer_allreduce_sum(const gp_peer_t* peer, float* data, int n, void* stream) {
  GP_REQUIRE(data != nullptr && n > 0 && n <= kPeerMaxFloats, "gp_peer_allreduce_sum: n must be 1..%d", kPeerMaxFloats);
  PeerCtx px;
  int rc = make_ctx(peer, &px);
  if (rc) return rc;
  gp::launch_plain(peer_allreduce_kernel, 1, 512, 0, as_stream(stream), px, data, n);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_finalize_peer(const gp_peer_t* peer, float* st, double count, int C, float eps, float momentum,
                        const float* gamma, const float* beta, float* mean, float* rstd, float* scale, float* shift,
                        float* running_mean, float* running_var, long long* num_batches_tracked, void* stream) {
  GP_REQUIRE(st && mean && rstd && scale && shift && C > 0 && 2 * C <= kPeerMaxFloats && count > 0,
             "gp_bn_finalize_peer: bad arguments (C <= %d)", kPeerMaxFloats / 2);
  PeerCtx px;
  int rc = make_ctx(peer, &px);
  if (rc) return rc;
  gp::launch_plain(bn_finalize_peer_kernel, 1, 512, 0, as_stream(stream), px, st, count, C, eps, momentum, gamma, beta, mean, rstd, scale,
                                                           shift, running_mean, running_var, num_batches_tracked);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

}  // extern "C"
