// Implicit-GEMM convolution on tcgen05 / TMEM, fed by TMA-staged NHWC bf16 tiles (sm_100a only).
//
// One persistent, warp-specialised kernel template covers every tensor-core contraction of the
// DCGAN-family step (reference call sites: models/dcgan.py:36,106; models/sngan_projection.py:30-44):
//
//   MODE_FWD   out[pix, n] = act( sum_{tap, c} In[gather(pix, tap), c] * Wp[n, tap.koff + c] + bias[n] )
//              - Conv2d k4s2p1 fprop  / ConvTranspose2d k4s2p1 dgrad : 16 taps, stride-2 gather expressed as
//                four parity tensor maps (no elementStrides), zero padding = TMA out-of-bounds fill;
//              - ConvTranspose2d k4s2p1 fprop / Conv2d k4s2p1 dgrad  : 4 output-parity phases x 4 taps, dense
//                2x2 stride-1 gather, output scattered with stride 2 (no zero multiplies);
//              - k3s1p1 / k1s1 convs and plain Linear layers (1 tap).
//   MODE_WGRAD dW[m, tap.koff + n] += sum_{pix} Dense[pix, m] * Gath[gather(pix, tap), n]
//              both operands MN-major (channels contiguous in NHWC), split-K over pixels, fp32 atomics.
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer,
// warps 2..5 = epilogue (tcgen05.ld -> bias/activation -> global). Two TMEM accumulator buffers so the
// epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include <cuda_bf16.h>
#include "sm100_ptx.cuh"

namespace gp {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // one 128-byte swizzle row of bf16
constexpr int kUmmaK = 16;
constexpr int kMaxTaps = 16;
constexpr int kNumThreads = 192;

enum { MODE_FWD = 0, MODE_WGRAD = 1 };
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_TANH = 3 };

struct Tap {
  int8_t map;  // which gathered-operand tensor map (parity variant)
  int8_t dh;   // pixel offset on the small grid
  int8_t dw;
  int8_t pad;
  int32_t koff;  // FWD: K offset of this tap inside a packed weight row; WGRAD: column offset inside a dW row
};

struct alignas(64) ConvGemmParams {
  CUtensorMap map_g[4];  // gathered operand (FWD: A; WGRAD: B)
  CUtensorMap map_d;     // WGRAD: dense operand (A)
  CUtensorMap map_w;     // FWD: packed weights [N][Ktot] (B)
  Tap taps[kMaxTaps];
  int n_phases, taps_per_phase;
  int NB, Hs, Ws;  // small pixel grid
  int Nt, Ht, Wt;  // pixel tile factors (product = 128 for FWD, 64 for WGRAD)
  int C;           // FWD: contraction channels per tap
  int N;           // FWD: output channels; WGRAD: channels of the gathered operand (dW columns per tap)
  int M;           // WGRAD: channels of the dense operand (dW rows)
  long long out_sN, out_sH, out_sW;  // FWD: output strides in elements for pixel (n, h, w) of the small grid
  long long phase_off[4];            // FWD: output element offset of each phase
  __nv_bfloat16* out;
  const float* bias;
  int act;
  float* dw;  // WGRAD: fp32 [M][ldw]
  int ldw;
  int splits;
  int kblocks_total;  // WGRAD: number of 64-pixel blocks
  float* col_sum;     // optional fused per-channel statistics of the (pre-activation) output
  float* col_sumsq;
};

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : ((BN == 128) ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN;  // 128 / 256 / 512: all powers of two >= 32
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // +1024: manual alignment slack
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_LRELU) return v > 0.f ? v : 0.2f * v;
  if (act == ACT_TANH) return tanhf(v);
  return v;
}

template <int MODE, int BN>
__global__ void __launch_bounds__(kNumThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.map_g[i]);
    tma_prefetch_desc(MODE == MODE_FWD ? &p.map_w : &p.map_d);
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- tile bookkeeping shared by all roles
  const int wtiles = p.Ws / p.Wt, htiles = p.Hs / p.Ht;
  const int ntiles_n = (p.N + BN - 1) / BN;
  int num_tiles, ksteps_fwd = 0, cchunks = 0, mtiles = 0, kb_per_split = 0;
  if constexpr (MODE == MODE_FWD) {
    mtiles = ((p.NB + p.Nt - 1) / p.Nt) * htiles * wtiles;
    cchunks = (p.C + kBlockK - 1) / kBlockK;
    ksteps_fwd = p.taps_per_phase * cchunks;
    num_tiles = p.n_phases * mtiles * ntiles_n;
  } else {
    mtiles = (p.M + kBlockM - 1) / kBlockM;
    kb_per_split = (p.kblocks_total + p.splits - 1) / p.splits;
    num_tiles = p.splits * mtiles * p.taps_per_phase * ntiles_n;
  }

  if (warp == 0 && lane == 0) {
    // =========================== TMA producer ===========================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      if constexpr (MODE == MODE_FWD) {
        const int nt = tile % ntiles_n;
        const int mt = (tile / ntiles_n) % mtiles;
        const int ph = tile / (ntiles_n * mtiles);
        const int w0 = (mt % wtiles) * p.Wt;
        const int h0 = ((mt / wtiles) % htiles) * p.Ht;
        const int n0 = (mt / (wtiles * htiles)) * p.Nt;
        for (int ks = 0; ks < ksteps_fwd; ++ks) {
          const Tap t = p.taps[ph * p.taps_per_phase + ks / cchunks];
          const int c0 = (ks % cchunks) * kBlockK;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_4d(sa, &p.map_g[t.map], &full_bar[stage], c0, w0 + t.dw, h0 + t.dh, n0);
          tma_load_2d(sa + Cfg::kABytes, &p.map_w, &full_bar[stage], t.koff + c0, nt * BN);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      } else {
        const int nt = tile % ntiles_n;
        const int tp = (tile / ntiles_n) % p.taps_per_phase;
        const int mt = (tile / (ntiles_n * p.taps_per_phase)) % mtiles;
        const int sp = tile / (ntiles_n * p.taps_per_phase * mtiles);
        const Tap t = p.taps[tp];
        const int kb0 = sp * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, p.kblocks_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int w0 = (kb % wtiles) * p.Wt;
          const int h0 = ((kb / wtiles) % htiles) * p.Ht;
          const int n0 = (kb / (wtiles * htiles)) * p.Nt;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
#pragma unroll
          for (int i = 0; i < kBlockM / 64; ++i)
            tma_load_4d(sa + i * 8192, &p.map_d, &full_bar[stage], mt * kBlockM + i * 64, w0, h0, n0);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_4d(sa + Cfg::kABytes + j * 8192, &p.map_g[t.map], &full_bar[stage], nt * BN + j * 64,
                        w0 + t.dw, h0 + t.dh, n0);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // =========================== MMA issuer (one thread) ===========================
    constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN, MODE == MODE_WGRAD, MODE == MODE_WGRAD);
    // K-major: SBO = 8 rows * 128 B. MN-major: SBO = 8 K-rows * 128 B, LBO = 64 K-rows * 128 B (next 64-channel chunk).
    constexpr uint64_t dbase = (MODE == MODE_FWD) ? make_smem_desc_base(0, 1024) : make_smem_desc_base(8192, 1024);
    constexpr uint32_t kadv = (MODE == MODE_FWD) ? (kUmmaK * 2) : (kUmmaK * 128);  // bytes per UMMA_K step
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int ksteps;
      if constexpr (MODE == MODE_FWD) {
        ksteps = ksteps_fwd;
      } else {
        const int sp = tile / (ntiles_n * p.taps_per_phase * mtiles);
        const int kb0 = sp * kb_per_split;
        ksteps = min(kb0 + kb_per_split, p.kblocks_total) - kb0;
      }
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          umma_bf16(d_tmem, smem_desc(dbase, a_addr + k * kadv), smem_desc(dbase, b_addr + k * kadv), idesc,
                    (ks | k) != 0);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tfull_bar[acc]);  // accumulator ready for the epilogue
    }
  } else if (warp >= 2) {
    // =========================== epilogue (4 warps, one TMEM lane quarter each) ===========================
    const int q = warp & 3;  // warps 2,3,4,5 -> quarters 2,3,0,1
    const int row = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
      if constexpr (MODE == MODE_FWD) {
        const int nt = tile % ntiles_n;
        const int mt = (tile / ntiles_n) % mtiles;
        const int ph = tile / (ntiles_n * mtiles);
        const int w = (mt % wtiles) * p.Wt + row % p.Wt;
        const int h = ((mt / wtiles) % htiles) * p.Ht + (row / p.Wt) % p.Ht;
        const int n = (mt / (wtiles * htiles)) * p.Nt + row / (p.Wt * p.Ht);
        const bool row_ok = n < p.NB;
        __nv_bfloat16* orow = p.out + p.phase_off[ph] + n * p.out_sN + h * p.out_sH + w * p.out_sW;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = nt * BN + c * 32;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            float b = (p.bias != nullptr && col < p.N) ? __ldg(p.bias + col) : 0.f;
            v[j] = __uint_as_float(r[j]) + b;
          }
          if (p.col_sum != nullptr) {
            // per-channel sum / sum of squares of the fp32 pre-activation output over this warp's 32 rows
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float s = row_ok ? v[j] : 0.f;
              float s2 = s * s;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
              }
              if (lane == j && col0 + j < p.N) {
                atomicAdd(p.col_sum + col0 + j, s);
                atomicAdd(p.col_sumsq + col0 + j, s2);
              }
            }
          }
          if (row_ok) {
            if (col0 + 32 <= p.N) {
              uint4* dst = reinterpret_cast<uint4*>(orow + col0);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint32_t w32[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  __nv_bfloat162 b2 = __floats2bfloat162_rn(apply_act(v[g * 8 + e * 2], p.act),
                                                            apply_act(v[g * 8 + e * 2 + 1], p.act));
                  w32[e] = *reinterpret_cast<uint32_t*>(&b2);
                }
                dst[g] = make_uint4(w32[0], w32[1], w32[2], w32[3]);
              }
            } else {
              for (int j = 0; j < 32 && col0 + j < p.N; ++j) orow[col0 + j] = __float2bfloat16(apply_act(v[j], p.act));
            }
          }
        }
      } else {
        const int nt = tile % ntiles_n;
        const int tp = (tile / ntiles_n) % p.taps_per_phase;
        const int mt = (tile / (ntiles_n * p.taps_per_phase)) % mtiles;
        const int m = mt * kBlockM + row;
        const bool row_ok = m < p.M;
        float* drow = p.dw + static_cast<long long>(m) * p.ldw + p.taps[tp].koff;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = nt * BN + c * 32;
          if (col0 >= p.N) break;
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) atomicAdd(drow + col0 + j, __uint_as_float(r[j]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace gp
