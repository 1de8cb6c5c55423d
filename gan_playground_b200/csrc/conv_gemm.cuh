// Implicit-GEMM convolution on tcgen05 / TMEM, fed by TMA-staged NHWC bf16 tiles (sm_100a only).
//
// One persistent, warp-specialised kernel template covers every tensor-core contraction of the
// DCGAN-family step (reference call sites: models/dcgan.py:36,106; models/sngan_projection.py:30-44):
//
//   MODE_FWD   out[pix, n] = act( sum_{tap, c} In[gather(pix, tap), c] * Wp[n, tap.koff + c] + bias[n] (+ residual) )
//              - Conv2d k4s2p1 fprop  / ConvTranspose2d k4s2p1 dgrad : 16 taps, stride-2 gather expressed as
//                four parity tensor maps (no elementStrides), zero padding = TMA out-of-bounds fill;
//              - ConvTranspose2d k4s2p1 fprop / Conv2d k4s2p1 dgrad  : 4 output-parity phases x 4 taps, dense
//                2x2 stride-1 gather, output scattered with stride 2 (no zero multiplies);
//              - k3s1p1 / k1s1 convs and plain Linear layers (1 tap).
//   MODE_WGRAD dW[m, tap.koff + n] += sum_{pix} Dense[pix, m] * Gath[gather(pix, tap), n]
//              both operands MN-major (channels contiguous in NHWC), split-K over pixels, fp32 atomics.
//
// Tile = (MT * 128) x BN: MT = 2 issues two M=128 MMAs per K step that share the B tile, which doubles the
// arithmetic intensity against L2 / shared memory (256-row tiles). Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM
// allocator + single-thread tcgen05.mma issuer, warps 2..5 = epilogue (tcgen05.ld -> bias / residual / activation ->
// global, optional per-channel BatchNorm statistics). Two TMEM accumulator buffers: the epilogue of tile i overlaps
// the MMAs of tile i+1.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.h"
#include "sm100_ptx.cuh"

namespace gp {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // one 128-byte swizzle row of bf16
constexpr int kUmmaK = 16;
constexpr int kMaxTaps = 48;
constexpr int kMaxMaps = 8;
constexpr int kNumThreads = 192;
constexpr int kMaxStatCols = 2048;

enum { MODE_FWD = 0, MODE_WGRAD = 1 };
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_TANH = 3 };

struct Tap {
  int8_t map;  // which gathered-operand tensor map (parity / hi-lo variant)
  int8_t dh;   // pixel offset on the small grid
  int8_t dw;
  int8_t pad;
  int32_t koff;  // FWD: K offset of this tap inside a packed weight row; WGRAD: column offset inside a dW row
};

struct alignas(64) ConvGemmParams {
  CUtensorMap map_g[kMaxMaps];  // gathered operand (FWD: A; WGRAD: B)
  CUtensorMap map_d;            // WGRAD: dense operand (A)
  CUtensorMap map_w;            // FWD: packed weights [N][Ktot] (B)
  Tap taps[kMaxTaps];
  int n_phases, taps_per_phase;
  int NB, Hs, Ws;  // small pixel grid
  int Nt, Ht, Wt;  // pixel tile factors (product = MT*128 for FWD, 64 for WGRAD)
  int C;           // FWD: contraction channels per tap
  int N;           // FWD: output channels; WGRAD: channels of the gathered operand (dW columns per tap)
  int M;           // WGRAD: channels of the dense operand (dW rows)
  long long out_sN, out_sH, out_sW;  // FWD: output strides in elements for pixel (n, h, w) of the small grid
  long long phase_off[4];            // FWD: output element offset of each phase
  __nv_bfloat16* out;
  __nv_bfloat16* out_lo;          // optional: low half of a hi/lo bf16 pair (out = hi, out_lo = bf16(v - hi))
  float* out_f32;                 // optional: store fp32 instead of bf16 (pre-BatchNorm outputs of the bf16x3 mode)
  const __nv_bfloat16* residual;  // optional, same indexing as out
  const float* bias;
  float act_slope;  // activation as max(v,0) + slope*min(v,0): 1 = identity, 0 = ReLU, 0.2 = LeakyReLU
  float* dw;        // WGRAD: fp32 [M][ldw]
  int ldw;
  int splits;
  int kblocks_total;  // WGRAD: number of 64-pixel blocks
  int lo_koff;        // FWD, bf16x3 kernels: K offset of the lo block inside a packed weight row [hi taps | lo taps]
  int wg_flat;        // WGRAD: dW columns are tiled over the flattened (tap, channel) index (needs N % 64 == 0), so one
                      // tile can span several taps of a narrow gathered operand; 0 = one tap per tile
  float* col_sum;     // optional fused per-channel statistics of the (pre-activation) output
  float* col_sumsq;
  // FWD run as the DATA GRADIENT of the next layer (its output is dA of the producing layer): backward work of the
  // producing layer done in this epilogue.  bwd_mode 1: out = v * act'(a), a = bwd_src (bf16 activation, layout of
  // out; ReLU / LeakyReLU by sign) — replaces the producer's act_bwd pass.  bwd_mode 2 / 3: BatchNorm-backward
  // reduction of the producer, col_sum += sum dz, col_sumsq += sum dz * xhat with dz = bf16(v) * act'(scale*y + shift),
  // xhat = (y - mean) * rstd, y = bwd_src (fp32 for mode 2, bf16 for mode 3), bwd_fin = [mean | rstd | scale | shift][N];
  // out = v unchanged — replaces the producer's bn_bwd_reduce pass (a second read of dA and y).
  const void* bwd_src;
  const float* bwd_fin;
  float bwd_slope;
  int bwd_mode;
  int fmt_flags;      // FWD: kFmtInF16 = both operands are fp16 (tcgen05 kind::f16 takes fp16 or bf16: only the instruction
                      // descriptor changes, tiles are 2-byte elements either way); kFmtLoF16 = out_lo receives fp16(v), the
                      // single-MMA operand copy of the "fp16" forward mode, instead of the bf16 rounding residual
};
constexpr int kFmtInF16 = 1, kFmtLoF16 = 2;
constexpr int kFmtResF16 = 4;  // FWD: `residual` holds fp16 (the companion of the shortcut activation in the fp16 mode)
constexpr int kFmtOutF16 = 8;  // FWD: `out` receives fp16(v) instead of bf16(v) (pre-BatchNorm output of the fp16 mode)

// X3 = bf16x3 forward (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo): one pipeline stage holds the hi AND lo tiles of both
// operands — 4 tile loads feed 3 MMA blocks, i.e. 1/3 fewer operand bytes from L2 per MMA than issuing the three
// products as separate K steps (the 128x256 mainloop is bound by L2->SM traffic, not by the tensor pipe).
// C2 = CTA pair (tcgen05 cta_group::2, cluster of two CTAs): one MMA of M = 256 per K step; each CTA stages its own 128
// rows of A and HALF of the B tile, so per FLOP it moves 1.5x fewer operand bytes into shared memory than a lone 128 x 256
// CTA (what bounds that mainloop) and, unlike the single-CTA 256 x 256 tile, keeps two accumulator buffers.
template <int BN, int MT, bool X3 = false, bool C2 = false>
struct GemmCfg {
  static_assert(!C2 || ((BN == 256 || BN == 128) && MT == 1), "the CTA-pair variants are the 2 x (128 x {128, 256}) tiles");
  static constexpr int kABytes = MT * kBlockM * kBlockK * 2;   // one A tile (hi or lo)
  static constexpr int kBBytes = (C2 ? BN / 2 : BN) * kBlockK * 2;  // one B tile (hi or lo); a pair CTA holds half of it
  static constexpr int kStageBytes = (X3 ? 2 : 1) * (kABytes + kBBytes);
  static constexpr int kStatBytes = 2 * kMaxStatCols * 4;
  static constexpr int kBiasBytes = 2 * 256 * 4;  // double-buffered bias slice of the current / next tile
  static constexpr int kBarBytes = 256;
  static constexpr int kStoreStageBytes = 4 * 2048;  // epilogue store staging: 2 KB per epilogue warp (coalesced stores)
  static constexpr int kBudget = 196 * 1024;  // operand ring; + statistics + bias + barriers + staging + slack < 227 KB
  static constexpr int kStages = kBudget / kStageBytes > 8 ? 8 : kBudget / kStageBytes;
  static constexpr int kAccCols = MT * BN;      // TMEM columns of one accumulator buffer
  // two accumulator buffers (epilogue of tile i overlaps the MMAs of tile i+1) when they fit the 512 TMEM columns;
  // the 256x256 tile uses all of TMEM for one buffer and trades that overlap for 1.5x less operand traffic per FLOP
  static constexpr int kAccBufs = 2 * kAccCols <= 512 ? 2 : 1;
  static constexpr int kTmemCols = kAccBufs * kAccCols;  // 128 / 256 / 512: powers of two >= 32
  static constexpr int kChunks = MT * (BN / 32);  // 32-column TMEM loads per accumulator buffer and epilogue warp
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kStatBytes + kBiasBytes + kBarBytes + kStoreStageBytes + 1024;  // +1024: alignment slack
  static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM allocation must be a power of two <= 512");
  static_assert(kStages >= (X3 ? 2 : 3), "pipeline too shallow");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
};

// Column sums over the 32 lanes of a warp for 32 per-lane values: butterfly transpose-reduce, 31 shuffles.
// On return lane l holds in v[0] the sum over lanes of the original v[l].
__device__ __forceinline__ void warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = upper ? v[i] : v[i + n / 2];
      const float keep = upper ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

// Coalesced epilogue stores. After tcgen05.ld every lane owns one ROW of a 32-row x 64-byte chunk; storing it directly
// makes each 16-byte store instruction touch 32 different rows (ncu on the K=64 image-side GEMM: 16 of 32 bytes per
// sector used, the L1/L2 store path is the limiter). Instead the warp transposes the chunk through 2 KB of shared memory:
// lane l then stores segment (l & 3) of rows (l >> 2) + 8 i, i.e. every instruction writes 8 rows x 64 contiguous bytes
// (full sectors). Row addresses / validity of the other lanes' rows come by shuffle.
//   stage: this warp's 2 KB buffer (shared address); seg[4]: the lane's own row as four 16-byte vectors; base +
//   row_byte_off: global address of the lane's own row; row_ok: that row is inside the tensor; col_lim: number of valid
//   16-byte segments (partial last column tile).
// RED: the 16-byte vectors are fp32 quadruples ADDED to global memory (red.global.add.v4.f32, the weight-gradient
// epilogue): the same transposition makes every reduction instruction cover full 32-byte sectors of 8 rows instead of
// half a sector of 32 rows - half as many L2 atomic transactions per tile.
template <bool RED = false>
__device__ __forceinline__ void store_rows_coalesced(uint32_t stage, const uint4 (&seg)[4], uint8_t* __restrict__ base,
                                                     long long row_byte_off, bool row_ok, int col_lim, int lane) {
  // swizzle: 16-byte slot (s ^ ((row >> 1) & 3)) of a 64-byte row -> conflict-free for both the row-wise writes and the
  // (8 rows x 4 segments) reads
#pragma unroll
  for (int sg = 0; sg < 4; ++sg) {
    const uint32_t a = stage + lane * 64 + ((sg ^ ((lane >> 1) & 3)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(seg[sg].x), "r"(seg[sg].y), "r"(seg[sg].z),
                 "r"(seg[sg].w)
                 : "memory");
  }
  __syncwarp();
  const uint32_t okmask = __ballot_sync(0xffffffffu, row_ok);
  const int sg = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (lane >> 2) + 8 * i;
    const long long off = __shfl_sync(0xffffffffu, row_byte_off, r);
    uint4 v;
    const uint32_t a = stage + r * 64 + ((sg ^ ((r >> 1) & 3)) << 4);
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    if (((okmask >> r) & 1u) && sg < col_lim) {
      if constexpr (RED) red_add_v4(reinterpret_cast<float*>(base + off + sg * 16), v.x, v.y, v.z, v.w);
      else *reinterpret_cast<uint4*>(base + off + sg * 16) = v;
    }
  }
  __syncwarp();  // the buffer is rewritten by the next call
}

// One unit of work of a persistent CTA: an output tile and (WGRAD) the range of 64-pixel K blocks it accumulates.
struct WorkItem {
  int tile;
  int kb0, kb1;
};

// Work distribution.
//   FWD  : tiles are dealt round-robin (tile = blockIdx.x + it * gridDim.x).
//   WGRAD: T = splits * base_tiles tiles ordered split-major (so concurrently running CTAs sweep the same pixel range
//          and share operand tiles in L2). The first floor(T / G) * G tiles are dealt round-robin as whole tiles; the
//          K ranges of the remaining R < G tiles are cut into G equal pieces (each CTA gets <= 2 segments), so every
//          SM executes the same number of MMA K steps — no partial last wave ("hybrid stream-K"); the fp32 atomic
//          epilogue makes partial tiles free of extra bookkeeping.
struct WorkPlan {
  int num_items;
  int nfull, rem_tiles, base_tiles, per, kblocks_total;
  int num_tiles;

  int first, stride;  // index of this CTA (or CTA pair) among the work units and their number

  template <int MODE>
  __device__ __forceinline__ WorkItem item(int it) const {
    WorkItem w;
    const int G = stride;  // number of work units running side by side: CTAs, or CTA pairs
    if constexpr (MODE == MODE_FWD) {
      w.tile = first + it * stride;
      w.kb0 = 0;
      w.kb1 = 1;
    } else {
      if (it < nfull) {
        w.tile = first + it * G;
        const int sp = w.tile / base_tiles;
        w.kb0 = sp * per;
        w.kb1 = min(w.kb0 + per, kblocks_total);
      } else {
        const int j = it - nfull;
        const long long tot = (long long)rem_tiles * per;
        const long long r0 = tot * first / G, r1 = tot * (first + 1) / G;
        const int t = (int)(r0 / per) + j;                       // remainder tile touched by segment j
        const long long lo = j == 0 ? r0 : (long long)t * per;
        const long long hi = min(r1, (long long)(t + 1) * per);
        w.tile = nfull * G + t;
        const int sp = w.tile / base_tiles;
        w.kb0 = sp * per + (int)(lo - (long long)t * per);
        w.kb1 = hi > lo ? min(sp * per + (int)(hi - (long long)t * per), kblocks_total) : w.kb0;
        if (w.kb1 < w.kb0) w.kb1 = w.kb0;
      }
    }
    return w;
  }
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int MODE, int BN, int MT, bool X3 = false, bool C2 = false>
__global__ void __launch_bounds__(kNumThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  static_assert(!(X3 && MODE == MODE_WGRAD), "bf16x3 applies to the forward GEMMs only");
  using Cfg = GemmCfg<BN, MT, X3, C2>;
  gp::pdl_launch_dependents();  // the next launch may become resident as this grid's CTAs retire (common.h)
  const uint32_t pair_rank = C2 ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs), 1 = peer
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  float* s_stat = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes);
  float* s_bias = s_stat + 2 * kMaxStatCols;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kStatBytes + Cfg::kBiasBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* s_store = smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kStatBytes + Cfg::kBiasBytes + Cfg::kBarBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool do_stats = (MODE == MODE_FWD) && p.col_sum != nullptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_g[0]);
    tma_prefetch_desc(MODE == MODE_FWD ? &p.map_w : &p.map_d);
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], C2 ? 8 : 4);  // one arrival per epilogue warp (pair: of BOTH CTAs, on the leader's barrier)
    }
    fence_mbar_init();
  }
  if constexpr (C2) cluster_sync_all();  // the peer's barriers exist before anything can arrive on them
  // Everything above touched shared memory and kernel parameters only. From here on the grids this one depends on must have
  // completed and flushed. The wait also comes BEFORE the TMEM allocation: a CTA parked here while its predecessor still
  // runs must not hold (or queue for) tensor memory that a co-resident CTA of the predecessor has yet to allocate.
  gp::pdl_wait();
  if (warp == 1) {
    if constexpr (C2) {
      tmem_alloc_2cta(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  if (do_stats) {
    for (int i = threadIdx.x; i < 2 * p.N; i += blockDim.x) s_stat[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- tile bookkeeping shared by all roles
  const int wtiles = p.Ws / p.Wt, htiles = p.Hs / p.Ht;
  // WGRAD: dW columns seen by the tiler (all taps side by side when wg_flat) and tile slots per column tile
  const int ncols = (MODE == MODE_WGRAD && p.wg_flat) ? p.ldw : p.N;
  const int tap_slots = (MODE == MODE_WGRAD && p.wg_flat) ? 1 : p.taps_per_phase;
  const int ntiles_n = (ncols + BN - 1) / BN;
  int ksteps_fwd = 0, cchunks = 0, mtiles = 0;
  WorkPlan plan;
  plan.first = C2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  plan.stride = C2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  if constexpr (MODE == MODE_FWD) {
    mtiles = ((p.NB + p.Nt - 1) / p.Nt) * htiles * wtiles;
    if constexpr (C2) mtiles = (mtiles + 1) / 2;  // pair tiles: CTA r of the pair owns 128-pixel tile 2 * mt + r
    cchunks = (p.C + kBlockK - 1) / kBlockK;
    ksteps_fwd = p.taps_per_phase * cchunks;
    plan.num_tiles = p.n_phases * mtiles * ntiles_n;
    plan.num_items = plan.num_tiles > plan.first ? (plan.num_tiles - plan.first + plan.stride - 1) / plan.stride : 0;
  } else {
    // pair: a work unit covers 256 dW rows, CTA r of the pair owns rows [256 mt + 128 r, + 128)
    mtiles = C2 ? (p.M + 2 * kBlockM - 1) / (2 * kBlockM) : (p.M + MT * kBlockM - 1) / (MT * kBlockM);
    plan.base_tiles = mtiles * tap_slots * ntiles_n;
    plan.kblocks_total = p.kblocks_total;
    plan.per = (p.kblocks_total + p.splits - 1) / p.splits;
    plan.num_tiles = p.splits * plan.base_tiles;
    plan.nfull = plan.num_tiles / plan.stride;
    plan.rem_tiles = plan.num_tiles - plan.nfull * plan.stride;
    plan.num_items = plan.nfull + (plan.rem_tiles > 0 ? 2 : 0);
  }

  if (warp == 0 && lane == 0) {
    // =========================== TMA producer ===========================
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < plan.num_items; ++it) {
      const WorkItem wk = plan.template item<MODE>(it);
      const int tile = wk.tile;
      if constexpr (MODE == MODE_FWD) {
        const int nt = tile % ntiles_n;
        const int mt = C2 ? 2 * ((tile / ntiles_n) % mtiles) + (int)pair_rank : (tile / ntiles_n) % mtiles;
        const int ph = tile / (ntiles_n * mtiles);
        const int w0 = (mt % wtiles) * p.Wt;
        const int h0 = ((mt / wtiles) % htiles) * p.Ht;
        const int n0 = (mt / (wtiles * htiles)) * p.Nt;  // a pair tile past the end (odd tile count): all out of bounds
        for (int ks = 0; ks < ksteps_fwd; ++ks) {
          const Tap t = p.taps[ph * p.taps_per_phase + ks / cchunks];
          const int c0 = (ks % cchunks) * kBlockK;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          if constexpr (C2) {
            // both CTAs' bytes are counted on the leader's barrier, which the leader arms for the pair; this CTA loads
            // its own rows of A and its half of the B columns
            if (pair_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            const int brow = nt * BN + (int)pair_rank * (BN / 2);
            if constexpr (X3) {
              tma_load_4d_2cta(sa, &p.map_g[t.map], &full_bar[stage], c0, w0 + t.dw, h0 + t.dh, n0);
              tma_load_4d_2cta(sa + Cfg::kABytes, &p.map_g[t.map + 4], &full_bar[stage], c0, w0 + t.dw, h0 + t.dh, n0);
              tma_load_2d_2cta(sa + 2 * Cfg::kABytes, &p.map_w, &full_bar[stage], t.koff + c0, brow);
              tma_load_2d_2cta(sa + 2 * Cfg::kABytes + Cfg::kBBytes, &p.map_w, &full_bar[stage], p.lo_koff + t.koff + c0, brow);
            } else {
              tma_load_4d_2cta(sa, &p.map_g[t.map], &full_bar[stage], c0, w0 + t.dw, h0 + t.dh, n0);
              tma_load_2d_2cta(sa + Cfg::kABytes, &p.map_w, &full_bar[stage], t.koff + c0, brow);
            }
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if constexpr (X3) {  // stage = [A_hi | A_lo | B_hi | B_lo]; the lo activation maps are map_g[4..7]
            tma_load_4d(sa, &p.map_g[t.map], &full_bar[stage], c0, w0 + t.dw, h0 + t.dh, n0);
            tma_load_4d(sa + Cfg::kABytes, &p.map_g[t.map + 4], &full_bar[stage], c0, w0 + t.dw, h0 + t.dh, n0);
            tma_load_2d(sa + 2 * Cfg::kABytes, &p.map_w, &full_bar[stage], t.koff + c0, nt * BN);
            tma_load_2d(sa + 2 * Cfg::kABytes + Cfg::kBBytes, &p.map_w, &full_bar[stage], p.lo_koff + t.koff + c0, nt * BN);
          } else {
            tma_load_4d(sa, &p.map_g[t.map], &full_bar[stage], c0, w0 + t.dw, h0 + t.dh, n0);
            tma_load_2d(sa + Cfg::kABytes, &p.map_w, &full_bar[stage], t.koff + c0, nt * BN);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      } else {
        const int nt = tile % ntiles_n;
        const int tp = (tile / ntiles_n) % tap_slots;
        const int mt = (tile / (ntiles_n * tap_slots)) % mtiles;
        // 64-column chunks of the gathered operand: (tap, first channel) of each; columns past the end read channel
        // N, which TMA zero-fills
        constexpr int kBChunks = C2 ? BN / 128 : BN / 64;  // 64-column chunks this CTA stages (a pair CTA: half of them)
        Tap tj[kBChunks];
        int chj[kBChunks];
#pragma unroll
        for (int j = 0; j < kBChunks; ++j) {
          const int col = nt * BN + ((C2 ? (int)pair_rank * kBChunks : 0) + j) * 64;
          if (p.wg_flat) {
            const int tap = col / p.N;
            tj[j] = p.taps[tap < p.taps_per_phase ? tap : 0];
            chj[j] = tap < p.taps_per_phase ? col % p.N : p.N;
          } else {
            tj[j] = p.taps[tp];
            chj[j] = col;
          }
        }
        for (int kb = wk.kb0; kb < wk.kb1; ++kb) {
          const int w0 = (kb % wtiles) * p.Wt;
          const int h0 = ((kb / wtiles) % htiles) * p.Ht;
          const int n0 = (kb / (wtiles * htiles)) * p.Nt;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          if constexpr (C2) {
            if (pair_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            const int m0 = (2 * mt + (int)pair_rank) * kBlockM;
#pragma unroll
            for (int i = 0; i < kBlockM / 64; ++i)
              tma_load_4d_2cta(sa + i * 8192, &p.map_d, &full_bar[stage], m0 + i * 64, w0, h0, n0);
#pragma unroll
            for (int j = 0; j < kBChunks; ++j)
              tma_load_4d_2cta(sa + Cfg::kABytes + j * 8192, &p.map_g[tj[j].map], &full_bar[stage], chj[j], w0 + tj[j].dw,
                               h0 + tj[j].dh, n0);
          } else {
            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
#pragma unroll
            for (int i = 0; i < MT * kBlockM / 64; ++i)
              tma_load_4d(sa + i * 8192, &p.map_d, &full_bar[stage], mt * MT * kBlockM + i * 64, w0, h0, n0);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_4d(sa + Cfg::kABytes + j * 8192, &p.map_g[tj[j].map], &full_bar[stage], chj[j], w0 + tj[j].dw,
                          h0 + tj[j].dh, n0);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0 && pair_rank == 0) {
    // =========================== MMA issuer (one thread; in a CTA pair: of the leader) ===========================
    // a_format / b_format (bits [7,10) / [10,13)): 1 = BF16, 0 = F16
    const uint32_t idesc = make_idesc_bf16(C2 ? 2 * kBlockM : kBlockM, BN, MODE == MODE_WGRAD, MODE == MODE_WGRAD) &
                           ~((MODE == MODE_FWD && (p.fmt_flags & kFmtInF16)) ? ((1u << 7) | (1u << 10)) : 0u);
    // K-major: SBO = 8 rows * 128 B. MN-major: SBO = 8 K-rows * 128 B, LBO = 64 K-rows * 128 B (next 64-channel chunk).
    constexpr uint64_t dbase = (MODE == MODE_FWD) ? make_smem_desc_base(0, 1024) : make_smem_desc_base(8192, 1024);
    constexpr uint32_t kadv = (MODE == MODE_FWD) ? (kUmmaK * 2) : (kUmmaK * 128);  // bytes per UMMA_K step
    constexpr uint32_t a_sub = kBlockM * kBlockK * 2;  // bytes between the MT sub-tiles of A (both modes: 16 KB)
    int stage = 0;
    uint32_t phase = 0;
    int na = 0;  // accumulator buffers handed to the epilogue so far
    for (int it = 0; it < plan.num_items; ++it) {
      const WorkItem wk = plan.template item<MODE>(it);
      const int ksteps = (MODE == MODE_FWD) ? ksteps_fwd : wk.kb1 - wk.kb0;
      if (ksteps <= 0) continue;
      const int acc = na % Cfg::kAccBufs;
      const uint32_t acc_phase = (na / Cfg::kAccBufs) & 1;
      ++na;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * Cfg::kAccCols;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
        if constexpr (X3) {
          const uint32_t a_lo = a_addr + Cfg::kABytes, b_hi = a_addr + 2 * Cfg::kABytes, b_lo = b_hi + Cfg::kBBytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t bh = smem_desc(dbase, b_hi + k * kadv), bl = smem_desc(dbase, b_lo + k * kadv);
#pragma unroll
            for (int mi = 0; mi < MT; ++mi) {
              const uint64_t ah = smem_desc(dbase, a_addr + mi * a_sub + k * kadv);
              const uint64_t al = smem_desc(dbase, a_lo + mi * a_sub + k * kadv);
              if constexpr (C2) {
                umma_bf16_2cta(d_tmem, ah, bh, idesc, (ks | k) != 0);
                umma_bf16_2cta(d_tmem, al, bh, idesc, 1u);
                umma_bf16_2cta(d_tmem, ah, bl, idesc, 1u);
              } else {
                umma_bf16(d_tmem + mi * BN, ah, bh, idesc, (ks | k) != 0);  // x_hi * w_hi
                umma_bf16(d_tmem + mi * BN, al, bh, idesc, 1u);             // x_lo * w_hi
                umma_bf16(d_tmem + mi * BN, ah, bl, idesc, 1u);             // x_hi * w_lo
              }
            }
          }
        } else {
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t bdesc = smem_desc(dbase, b_addr + k * kadv);
#pragma unroll
            for (int mi = 0; mi < MT; ++mi) {
              if constexpr (C2)
                umma_bf16_2cta(d_tmem, smem_desc(dbase, a_addr + k * kadv), bdesc, idesc, (ks | k) != 0);
              else
                umma_bf16(d_tmem + mi * BN, smem_desc(dbase, a_addr + mi * a_sub + k * kadv), bdesc, idesc, (ks | k) != 0);
            }
          }
        }
        // frees the smem slot (pair: in both CTAs) once these MMAs retire
        if constexpr (C2) umma_commit_2cta(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      // accumulator ready for the epilogue (pair: of both CTAs)
      if constexpr (C2) umma_commit_2cta(&tfull_bar[acc]); else umma_commit(&tfull_bar[acc]);
    }
  } else if (warp >= 2) {
    // =========================== epilogue (4 warps, one TMEM lane quarter each) ===========================
    const int q = warp & 3;  // warps 2,3,4,5 -> quarters 2,3,0,1
    const int et = threadIdx.x - 64;  // 0..127
    const bool has_bias = (MODE == MODE_FWD) && p.bias != nullptr;
    constexpr int kBiasPerThread = (BN + 127) / 128;
    // bias slice of a tile -> registers (issued one tile ahead) -> shared memory (double buffered)
    auto fetch_bias = [&](int tile, float (&b)[kBiasPerThread]) {
      const int nt = tile % ntiles_n;
#pragma unroll
      for (int k = 0; k < kBiasPerThread; ++k) {
        const int idx = k * 128 + et, col = nt * BN + idx;
        b[k] = (idx < BN && col < p.N) ? __ldg(p.bias + col) : 0.f;
      }
    };
    auto put_bias = [&](int buf, const float (&b)[kBiasPerThread]) {
#pragma unroll
      for (int k = 0; k < kBiasPerThread; ++k) {
        const int idx = k * 128 + et;
        if (idx < BN) s_bias[buf * 256 + idx] = b[k];
      }
    };
    if (has_bias && plan.num_items > 0) {
      float b0[kBiasPerThread];
      fetch_bias(plan.first, b0);
      put_bias(0, b0);
      epi_bar_sync();
    }
    int na = 0;
    for (int it = 0; it < plan.num_items; ++it) {
      const WorkItem wk = plan.template item<MODE>(it);
      if (MODE == MODE_WGRAD && wk.kb1 <= wk.kb0) continue;
      const int tile = wk.tile;
      const int acc = na % Cfg::kAccBufs;
      const uint32_t acc_phase = (na / Cfg::kAccBufs) & 1;
      ++na;
      float bnext[kBiasPerThread];
      const bool more = it + 1 < plan.num_items;
      if (has_bias && more) fetch_bias(tile + plan.stride, bnext);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + acc * Cfg::kAccCols + (static_cast<uint32_t>(q * 32) << 16);
      // 32-column chunks in the order (column chunk c, sub-tile mi). The loop stays rolled (a fully unrolled epilogue
      // overflows the instruction cache); the TMEM load of the next chunk is issued as soon as the current one has
      // been copied out of its registers, so it is in flight while the chunk is processed.
      uint32_t r[32];
      tmem_ld_32x32(tbase, r);
      if constexpr (MODE == MODE_FWD) {
        const int nt = tile % ntiles_n;
        const int mt = C2 ? 2 * ((tile / ntiles_n) % mtiles) + (int)pair_rank : (tile / ntiles_n) % mtiles;
        const int ph = tile / (ntiles_n * mtiles);
        const int w_t0 = (mt % wtiles) * p.Wt, h_t0 = ((mt / wtiles) % htiles) * p.Ht;
        const int n_t0 = (mt / (wtiles * htiles)) * p.Nt;
        const float slope = p.act_slope;
        const uint32_t sb = smem_u32(s_bias) + (it & 1) * 1024;
        long long off[MT];
        bool row_ok[MT];
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
          const int row = mi * kBlockM + q * 32 + lane;
          const int w = w_t0 + row % p.Wt;
          const int h = h_t0 + (row / p.Wt) % p.Ht;
          const int n = n_t0 + row / (p.Wt * p.Ht);
          row_ok[mi] = n < p.NB;
          off[mi] = p.phase_off[ph] + n * p.out_sN + h * p.out_sH + w * p.out_sW;
        }
        const int nchunks = min(BN / 32, (p.N - nt * BN + 31) / 32);  // partial last column tile
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
          const int col0 = nt * BN + c * 32;
          float s1[32], s2[32];
#pragma unroll
          for (int mi = 0; mi < MT; ++mi) {
            float v[32];
            tmem_ld_wait_regs(r);
            if (has_bias) {
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 b4 = lds_f32x4(sb + (c * 32 + 4 * g) * 4);  // same address for every lane: broadcast
                v[4 * g] = b4.x + __uint_as_float(r[4 * g]);
                v[4 * g + 1] = b4.y + __uint_as_float(r[4 * g + 1]);
                v[4 * g + 2] = b4.z + __uint_as_float(r[4 * g + 2]);
                v[4 * g + 3] = b4.w + __uint_as_float(r[4 * g + 3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            }
            // next chunk: (c, mi + 1) or (c + 1, 0)
            if (mi + 1 < MT) tmem_ld_32x32(tbase + (mi + 1) * BN + c * 32, r);
            else if (c + 1 < nchunks) tmem_ld_32x32(tbase + (c + 1) * 32, r);
            if (p.bwd_mode == 1 && row_ok[mi]) {
              // activation derivative of the producing layer from the sign of its (bf16) output
              const __nv_bfloat16* arow = static_cast<const __nv_bfloat16*>(p.bwd_src) + off[mi];
              const float sl = p.bwd_slope;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (col0 + g * 8 < p.N) {
                  const uint4 av = *reinterpret_cast<const uint4*>(arow + col0 + g * 8);
                  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&av);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(h2[e]);
                    v[g * 8 + 2 * e] *= f.x > 0.f ? 1.f : sl;
                    v[g * 8 + 2 * e + 1] *= f.y > 0.f ? 1.f : sl;
                  }
                }
              }
            }
            if (do_stats) {
              // per-channel sum / sum of squares of the fp32 pre-activation output: the MT rows of a thread are added
              // first, then the 32 rows of the warp via a shuffle butterfly, then shared-memory accumulators per CTA
              // (flushed once at kernel end)
              if (p.bwd_mode >= 2) {
                // BatchNorm-backward sums of the producing layer instead (see ConvGemmParams::bwd_mode)
                const float* fin = p.bwd_fin;
                const float sl = p.bwd_slope;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                  float y4[4] = {0.f, 0.f, 0.f, 0.f};
                  float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rs = mu, sc = mu, sh = mu;
                  const bool on = row_ok[mi] && col0 + g * 4 < p.N;
                  if (on) {
                    if (p.bwd_mode == 2) {
                      const float4 t = *reinterpret_cast<const float4*>(static_cast<const float*>(p.bwd_src) + off[mi] + col0 + g * 4);
                      y4[0] = t.x, y4[1] = t.y, y4[2] = t.z, y4[3] = t.w;
                    } else {
                      const uint2 t = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(p.bwd_src) + off[mi] + col0 + g * 4);
                      const float2 lo2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
                      const float2 hi2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
                      y4[0] = lo2.x, y4[1] = lo2.y, y4[2] = hi2.x, y4[3] = hi2.y;
                    }
                    mu = __ldg(reinterpret_cast<const float4*>(fin + col0 + g * 4));             // same address in every lane
                    rs = __ldg(reinterpret_cast<const float4*>(fin + p.N + col0 + g * 4));
                    sc = __ldg(reinterpret_cast<const float4*>(fin + 2 * p.N + col0 + g * 4));
                    sh = __ldg(reinterpret_cast<const float4*>(fin + 3 * p.N + col0 + g * 4));
                  }
                  const float m4[4] = {mu.x, mu.y, mu.z, mu.w}, r4[4] = {rs.x, rs.y, rs.z, rs.w};
                  const float c4[4] = {sc.x, sc.y, sc.z, sc.w}, h4[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const int j = g * 4 + e;
                    // the value the apply pass will read back: dA rounded to bf16
                    const float dav = on ? __bfloat162float(__float2bfloat16_rn(v[j])) : 0.f;
                    const float dz = dav * (y4[e] * c4[e] + h4[e] > 0.f ? 1.f : sl);
                    s1[j] = (mi == 0 ? 0.f : s1[j]) + dz;
                    s2[j] = (mi == 0 ? 0.f : s2[j]) + dz * (y4[e] - m4[e]) * r4[e];
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float x = row_ok[mi] ? v[j] : 0.f;
                  s1[j] = (mi == 0 ? 0.f : s1[j]) + x;
                  s2[j] = (mi == 0 ? 0.f : s2[j]) + x * x;
                }
              }
              if (mi == MT - 1) {
                warp_transpose_sum32(s1, lane);
                warp_transpose_sum32(s2, lane);
                if (col0 + lane < p.N) {
                  atomicAdd(&s_stat[col0 + lane], s1[0]);
                  atomicAdd(&s_stat[p.N + col0 + lane], s2[0]);
                }
              }
            }
            if (p.residual != nullptr && row_ok[mi]) {
              const __nv_bfloat16* rrow = p.residual + off[mi];
              const bool res_f16 = (p.fmt_flags & kFmtResF16) != 0;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (col0 + g * 8 < p.N) {
                  const uint4 rv = *reinterpret_cast<const uint4*>(rrow + col0 + g * 8);
                  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
                  const __half2* q2 = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = res_f16 ? __half22float2(q2[e]) : __bfloat1622float2(h2[e]);
                    v[g * 8 + 2 * e] += f.x;
                    v[g * 8 + 2 * e + 1] += f.y;
                  }
                }
              }
            }
            // activation as max(v, slope * v): slope 1 = identity, 0 = ReLU, 0.2 = LeakyReLU
            if (slope != 1.f) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], slope * v[j]);
            }
            // ---- stores, transposed through this warp's staging buffer so that each instruction writes whole sectors
            const uint32_t stage = smem_u32(s_store) + (warp - 2) * 2048;
            if (p.out_f32 != nullptr) {
              uint8_t* base = reinterpret_cast<uint8_t*>(p.out_f32 + col0);
#pragma unroll
              for (int hlf = 0; hlf < 2; ++hlf) {  // 32 fp32 columns = two 64-byte halves
                uint4 seg[4];
#pragma unroll
                for (int g = 0; g < 4; ++g)
                  seg[g] = make_uint4(__float_as_uint(v[hlf * 16 + 4 * g]), __float_as_uint(v[hlf * 16 + 4 * g + 1]),
                                      __float_as_uint(v[hlf * 16 + 4 * g + 2]), __float_as_uint(v[hlf * 16 + 4 * g + 3]));
                const int lim = (p.N - col0 - hlf * 16 + 3) / 4;  // valid 4-float segments of this half
                store_rows_coalesced(stage, seg, base + hlf * 64, off[mi] * 4, row_ok[mi], lim, lane);
              }
            } else {
              uint32_t w32[16];
              if (p.fmt_flags & kFmtOutF16) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const __half2 h2 = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                  w32[e] = *reinterpret_cast<const uint32_t*>(&h2);
                }
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
                  w32[e] = *reinterpret_cast<const uint32_t*>(&b2);
                }
              }
              const int lim = (p.N - col0 + 7) / 8;  // valid 8-element segments
              {
                uint4 seg[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) seg[g] = make_uint4(w32[4 * g], w32[4 * g + 1], w32[4 * g + 2], w32[4 * g + 3]);
                store_rows_coalesced(stage, seg, reinterpret_cast<uint8_t*>(p.out + col0), off[mi] * 2, row_ok[mi], lim, lane);
              }
              if (p.out_lo != nullptr) {
                if (p.fmt_flags & kFmtLoF16) {
#pragma unroll
                  for (int e = 0; e < 16; ++e) {
                    const __half2 h2 = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                    w32[e] = *reinterpret_cast<const uint32_t*>(&h2);
                  }
                } else {
#pragma unroll
                  for (int e = 0; e < 16; ++e) {
                    const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w32[e]));
                    const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
                    w32[e] = *reinterpret_cast<const uint32_t*>(&l2);
                  }
                }
                uint4 seg[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) seg[g] = make_uint4(w32[4 * g], w32[4 * g + 1], w32[4 * g + 2], w32[4 * g + 3]);
                store_rows_coalesced(stage, seg, reinterpret_cast<uint8_t*>(p.out_lo + col0), off[mi] * 2, row_ok[mi], lim, lane);
              }
            }
          }
        }
      } else {
        const int nt = tile % ntiles_n;
        const int tp = (tile / ntiles_n) % tap_slots;
        const int mt = (tile / (ntiles_n * tap_slots)) % mtiles;
        const int nchunks = min(BN / 32, (ncols - nt * BN + 31) / 32);
        const int col_base = p.wg_flat ? 0 : p.taps[tp].koff;  // dW row = [tap][channel]: flat columns are contiguous
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
          const int col0 = nt * BN + c * 32;
#pragma unroll
          for (int mi = 0; mi < MT; ++mi) {
            uint32_t v[32];
            tmem_ld_wait_regs(r);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = r[j];
            if (mi + 1 < MT) tmem_ld_32x32(tbase + (mi + 1) * BN + c * 32, r);
            else if (c + 1 < nchunks) tmem_ld_32x32(tbase + (c + 1) * 32, r);
            const int m = (C2 ? 2 * mt + (int)pair_rank : mt * MT + mi) * kBlockM + q * 32 + lane;
            // 16-byte vector reductions (red.global.add.v4.f32), transposed through the warp's staging buffer so that each
            // instruction adds 8 rows x 64 contiguous bytes (two 64-byte halves of the lane's 32 columns)
            const uint32_t stage = smem_u32(s_store) + (warp - 2) * 2048;
            uint8_t* base = reinterpret_cast<uint8_t*>(p.dw + col_base + col0);
            const long long row_off = static_cast<long long>(m) * p.ldw * 4;
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
              uint4 seg[4];
#pragma unroll
              for (int g = 0; g < 4; ++g)
                seg[g] = make_uint4(v[hlf * 16 + 4 * g], v[hlf * 16 + 4 * g + 1], v[hlf * 16 + 4 * g + 2], v[hlf * 16 + 4 * g + 3]);
              const int lim = (ncols - col0 - hlf * 16 + 3) / 4;  // valid 4-float segments of this half (<= 0: none)
              store_rows_coalesced<true>(stage, seg, base + hlf * 64, row_off, m < p.M, lim, lane);
            }
          }
        }
      }
      if (has_bias && more) put_bias((it + 1) & 1, bnext);
      tc_fence_before();
      __syncwarp();  // every lane's TMEM loads of this buffer have completed: one arrival per warp releases it
      if (lane == 0) {
        if constexpr (C2) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]);
      }
      if (has_bias) epi_bar_sync();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (do_stats) {
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) {
      atomicAdd(p.col_sum + i, s_stat[i]);
      atomicAdd(p.col_sumsq + i, s_stat[p.N + i]);
    }
  }
  if constexpr (C2) cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other can still signal it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (C2) tmem_dealloc_2cta(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace gp
