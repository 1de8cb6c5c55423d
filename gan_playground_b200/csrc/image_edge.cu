// Fused image-edge layers: the 4x4 stride-2 convolution between the fp32 NCHW image and the first / last 64-channel NHWC
// activation (D's first Conv2d, models/dcgan.py:106-109; the gradient side of G's last ConvTranspose2d + Tanh,
// models/dcgan.py:41-44) WITHOUT a column buffer in HBM.
//
// Both kernels build the im2col tile of 128 output pixels in shared memory straight from the image rows — fp32 -> bf16
// (or a hi/lo bf16 pair, or fp16), written in the 128-byte-swizzled UMMA layout by plain stores — and feed it to
// tcgen05.mma with the accumulator in TMEM:
//
//   gp_image_conv_k4s2_fwd  : out[px][n] = act(bias[n] + sum_j col[px][j] * w[n][j])      col tile = K-major A operand
//   gp_image_conv_k4s2_wgrad: dw[m][j] += sum_px dense[px][m] * col[px][j]                col tile = MN-major B operand
//                             (the SAME bytes in shared memory: a 128-byte row per pixel is the contiguous dimension
//                              of both views)
//   col[px = (n, oh, ow)][j = (c*4 + kh)*4 + kw] = img[n, c, 2oh-1+kh, 2ow-1+kw] * (mul ? 1 - mul[same]^2 : 1)
//
// They are HBM-bound (K = 48): per 128-pixel tile the forward reads 7.7 KB of image and writes 16 KB (32 KB with a
// companion tensor); the round trip of the 128-byte-per-pixel column buffer (written by im2col, re-read by the K = 64
// GEMM and again by wgrad) is gone. Several small CTAs (128 threads, 36-60 KB of shared memory, 64 TMEM columns) share
// an SM so that one CTA's image loads overlap another's MMA and stores; within a CTA the next tile's image rows are
// prefetched into registers while the current tile is processed.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.h"
#include "act_io.cuh"
#include "bn_stream.cuh"
#include "conv_gemm.cuh"

namespace gp {

constexpr int kEdgeThreads = 128;
constexpr int kEdgePatchFloats = 3072;  // ch * (2R + 2) * Wi <= 3 * (512 + 4 * Wo), Wo <= 128
constexpr int kEdgePatchVec = kEdgePatchFloats / 4 / kEdgeThreads;  // float4 per thread
constexpr int kEdgeMaxDenseVec = 16;  // 128 pixels x (M <= 128 channels) bf16 = 2048 16-byte chunks / 128 threads

struct ImageEdgeParams {
  const float* img;
  const float* mul;
  // forward
  const float* w;     // fp32 [N][ch*16] (torch layout of Conv2d / ConvTranspose2d weights with the image on the ch side)
  const float* bias;  // fp32 [N] or null
  __nv_bfloat16* out;
  void* out_comp;
  int comp_fmt;
  float slope;
  // weight gradient
  const __nv_bfloat16* dense;
  float* dw;     // fp32 [M][ch*16], accumulated
  float* dbias;  // fp32 [M] or null, accumulated (column sums of dense)
  int NB, Hi, Wi, N;  // N: forward output channels / wgrad dense channels (M)
  int tiles;
  int tmem_cols;
};

// geometry of a 128-pixel tile: R = 128 / Wo whole output rows of one image
struct EdgeGeom {
  int Ho, Wo, R, rows_in, w4;
};
__device__ __forceinline__ EdgeGeom edge_geom(const ImageEdgeParams& p) {
  EdgeGeom g;
  g.Ho = p.Hi / 2;
  g.Wo = p.Wi / 2;
  g.R = 128 / g.Wo;
  g.rows_in = 2 * g.R + 2;
  g.w4 = p.Wi / 4;
  return g;
}

// the tile's 2R + 2 input rows of every channel -> registers (zero above / below the image), tanh' fused when mul != null.
// Element i of the flattened [ch][rows_in][Wi / 4] patch goes to thread i % 128.
template <int CH>
__device__ __forceinline__ void patch_fetch(const ImageEdgeParams& p, const EdgeGeom& g, int tile, float4 (&pre)[kEdgePatchVec]) {
  const long long pix0 = (long long)tile * 128;
  const int n = (int)(pix0 / (g.Ho * g.Wo));
  const int oh0 = (int)(pix0 % (g.Ho * g.Wo)) / g.Wo;
  const int total = CH * g.rows_in * g.w4;
#pragma unroll
  for (int k = 0; k < kEdgePatchVec; ++k) {
    const int i = k * kEdgeThreads + threadIdx.x;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < total) {
      const int q = i % g.w4, rr = (i / g.w4) % g.rows_in, c = i / (g.w4 * g.rows_in);
      const int ih = 2 * oh0 - 1 + rr;
      if (ih >= 0 && ih < p.Hi) {
        const long long off = (((long long)n * CH + c) * p.Hi + ih) * p.Wi + 4 * q;
        v = __ldg(reinterpret_cast<const float4*>(p.img + off));
        if (p.mul != nullptr) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(p.mul + off));
          v.x *= 1.f - t.x * t.x, v.y *= 1.f - t.y * t.y, v.z *= 1.f - t.z * t.z, v.w *= 1.f - t.w * t.w;
        }
      }
    }
    pre[k] = v;
  }
}
template <int CH>
__device__ __forceinline__ void patch_store(const EdgeGeom& g, float* s_patch, const float4 (&pre)[kEdgePatchVec]) {
  const int total = CH * g.rows_in * g.w4;
#pragma unroll
  for (int k = 0; k < kEdgePatchVec; ++k) {
    const int i = k * kEdgeThreads + threadIdx.x;
    if (i < total) *reinterpret_cast<float4*>(s_patch + 4 * i) = pre[k];
  }
}

// Row r (one output pixel) of the column tile: 2 * CH 16-byte groups of 8 columns, group gi at byte
// r * 128 + ((gi ^ (r & 7)) << 4) — the SWIZZLE_128B pattern TMA would produce for a {64 elements, 128 rows} box.
// FMT: GP_COMP_NONE -> bf16 into t_hi; GP_COMP_LO -> bf16 hi into t_hi and bf16(v - hi) into t_lo; GP_COMP_F16 -> fp16 into t_hi.
template <int CH, int FMT>
__device__ __forceinline__ void build_col_row(const EdgeGeom& g, int Wi, const float* s_patch, int r, uint32_t t_hi, uint32_t t_lo) {
  const int ol = r / g.Wo, ow = r % g.Wo;
#pragma unroll
  for (int gi = 0; gi < 2 * CH; ++gi) {
    const int c = gi >> 1;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kh = (gi & 1) * 2 + (j >> 2), kw = j & 3;
      const int iw = 2 * ow - 1 + kw;
      f[j] = (iw >= 0 && iw < Wi) ? s_patch[(c * g.rows_in + 2 * ol + kh) * Wi + iw] : 0.f;
    }
    const Packed8c pk = pack8c(FMT, f);
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((gi ^ (r & 7)) << 4);
    if (FMT == GP_COMP_F16) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t_hi + off), "r"(pk.comp.x), "r"(pk.comp.y), "r"(pk.comp.z), "r"(pk.comp.w) : "memory");
    } else {
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t_hi + off), "r"(pk.hi.x), "r"(pk.hi.y), "r"(pk.hi.z), "r"(pk.hi.w) : "memory");
      if (FMT == GP_COMP_LO)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t_lo + off), "r"(pk.comp.x), "r"(pk.comp.y), "r"(pk.comp.z), "r"(pk.comp.w) : "memory");
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward
// shared memory: [A hi 16 KB | A lo 16 KB (bf16x3) | B hi Npad*128 | B lo (bf16x3) | patch 12 KB | bias | store staging 8 KB | barrier]
template <int CH, int FMT>
__global__ void __launch_bounds__(kEdgeThreads, 4) image_conv_fwd_kernel(const __grid_constant__ ImageEdgeParams p) {
  constexpr bool X3 = FMT == GP_COMP_LO;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int Npad = (p.N + 15) & ~15;
  const int bbytes = Npad * 128;
  uint8_t* sA = smem;
  uint8_t* sB = sA + (X3 ? 2 : 1) * 16384;
  float* s_patch = reinterpret_cast<float*>(sB + (X3 ? 2 : 1) * bbytes);
  float* s_bias = s_patch + kEdgePatchFloats;
  uint8_t* s_store = reinterpret_cast<uint8_t*>(s_bias + 160);
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_store + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const EdgeGeom g = edge_geom(p);

  float4 pre[kEdgePatchVec];
  int tile = blockIdx.x;
  if (tile < p.tiles) patch_fetch<CH>(p, g, tile, pre);

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  // weights -> B tile(s): row n = 128 bytes (K = CH*16 live elements), swizzled like the A rows; rows >= N are zero
  for (int n = tid; n < Npad; n += kEdgeThreads) {
#pragma unroll
    for (int gi = 0; gi < 2 * CH; ++gi) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = n < p.N ? __ldg(p.w + (long long)n * (CH * 16) + gi * 8 + j) : 0.f;
      const Packed8c pk = pack8c(FMT, f);
      uint8_t* d = sB + n * 128 + ((gi ^ (n & 7)) << 4);
      *reinterpret_cast<uint4*>(d) = FMT == GP_COMP_F16 ? pk.comp : pk.hi;
      if (X3) *reinterpret_cast<uint4*>(d + bbytes) = pk.comp;
    }
  }
  for (int i = tid; i < 160; i += kEdgeThreads) s_bias[i] = (p.bias != nullptr && i < p.N) ? __ldg(p.bias + i) : 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t idesc = make_idesc_bf16(kBlockM, Npad, 0, 0) & ~(FMT == GP_COMP_F16 ? ((1u << 7) | (1u << 10)) : 0u);
  constexpr uint64_t dbase = make_smem_desc_base(0, 1024);
  const uint32_t a_hi = smem_u32(sA), a_lo = a_hi + 16384, b_hi = smem_u32(sB), b_lo = b_hi + (uint32_t)bbytes;
  const float slope = p.slope;
  const int nch = (p.N + 31) / 32;
  uint32_t phase = 0;

  for (; tile < p.tiles; tile += gridDim.x) {
    patch_store<CH>(g, s_patch, pre);
    __syncthreads();
    const int next = tile + gridDim.x;
    if (next < p.tiles) patch_fetch<CH>(p, g, next, pre);  // in flight during the MMA and the epilogue
    build_col_row<CH, FMT>(g, p.Wi, s_patch, tid, a_hi, a_lo);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < CH; ++k) {  // K = CH * 16
        const uint64_t ah = smem_desc(dbase, a_hi + k * 32), bh = smem_desc(dbase, b_hi + k * 32);
        umma_bf16(tmem_base, ah, bh, idesc, k != 0);
        if (X3) {
          umma_bf16(tmem_base, smem_desc(dbase, a_lo + k * 32), bh, idesc, 1u);
          umma_bf16(tmem_base, ah, smem_desc(dbase, b_lo + k * 32), idesc, 1u);
        }
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // epilogue: thread = pixel row; 32-column chunks
    const long long row_off = ((long long)tile * 128 + tid) * p.N;
    const uint32_t stage = smem_u32(s_store) + warp * 2048;
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, r);
      tmem_ld_wait_regs(r);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = __uint_as_float(r[j]) + s_bias[c * 32 + j];
        v[j] = fmaxf(x, slope * x);
      }
      const int col0 = c * 32;
      const int lim = (p.N - col0 + 7) / 8;
      uint32_t w32[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        w32[e] = *reinterpret_cast<const uint32_t*>(&b2);
      }
      {
        uint4 seg[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) seg[q] = make_uint4(w32[4 * q], w32[4 * q + 1], w32[4 * q + 2], w32[4 * q + 3]);
        store_rows_coalesced(stage, seg, reinterpret_cast<uint8_t*>(p.out + col0), row_off * 2, true, lim, lane);
      }
      if (p.out_comp != nullptr) {
        if (p.comp_fmt == GP_COMP_F16) {
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
        for (int q = 0; q < 4; ++q) seg[q] = make_uint4(w32[4 * q], w32[4 * q + 1], w32[4 * q + 2], w32[4 * q + 3]);
        store_rows_coalesced(stage, seg, reinterpret_cast<uint8_t*>(static_cast<uint16_t*>(p.out_comp) + col0), row_off * 2, true,
                             lim, lane);
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM, the A tile and the patch are free for the next tile
  }

  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// shared memory: [A = dense tile, MN-major: 2 x (128 pixel rows x 128 B) = 32 KB | B = column tile 16 KB | patch 12 KB | barrier]
// Every CTA accumulates its share of the pixel tiles into ONE 128 x 64 fp32 accumulator in TMEM and adds it to dw once.
// Column CH*16 of the tile is the constant 1 (when a bias gradient is wanted), so accumulator column CH*16 = sum_px dense.
template <int CH>
__global__ void __launch_bounds__(kEdgeThreads, 3) image_conv_wgrad_kernel(const __grid_constant__ ImageEdgeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + 32768;
  float* s_patch = reinterpret_cast<float*>(sB + 16384);
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_patch + kEdgePatchFloats);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const EdgeGeom g = edge_geom(p);
  const int M = p.N;
  const int cpr = M / 8;  // 16-byte chunks per dense row
  const bool want_db = p.dbias != nullptr && CH < 4;
  const int ncols = CH * 16 + (want_db ? 16 : 0);

  // contiguous range of tiles for this CTA
  const int per = (p.tiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per, t1 = min(t0 + per, p.tiles);

  float4 pre[kEdgePatchVec];
  uint4 dpre[kEdgeMaxDenseVec];
  auto dense_fetch = [&](int tile) {
    const uint4* src = reinterpret_cast<const uint4*>(p.dense + (long long)tile * 128 * M);
#pragma unroll
    for (int k = 0; k < kEdgeMaxDenseVec; ++k) {
      const int i = k * kEdgeThreads + tid;
      if (i < 128 * cpr) dpre[k] = __ldg(src + i);
    }
  };
  auto dense_store = [&]() {
#pragma unroll
    for (int k = 0; k < kEdgeMaxDenseVec; ++k) {
      const int i = k * kEdgeThreads + tid;
      if (i < 128 * cpr) {
        const int px = i / cpr, j = i % cpr;
        *reinterpret_cast<uint4*>(sA + (j >> 3) * 16384 + px * 128 + (((j & 7) ^ (px & 7)) << 4)) = dpre[k];
      }
    }
  };
  if (t0 < t1) {
    patch_fetch<CH>(p, g, t0, pre);
    dense_fetch(t0);
  }

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  // zero the operand tiles once: dense channels beyond M and the padding columns of the column tile stay zero;
  // the constant-1 column of the bias gradient is written once as well (the tile builder never touches those groups)
  for (int i = tid; i < (32768 + 16384) / 16; i += kEdgeThreads) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  if (want_db) {
    const int r = tid;  // pixel row; group 2*CH holds columns CH*16 .. CH*16+7: (1, 0, 0, ...) in bf16
    *reinterpret_cast<uint4*>(sB + r * 128 + (((2 * CH) ^ (r & 7)) << 4)) = make_uint4(0x00003F80u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t idesc = make_idesc_bf16(kBlockM, ncols, 1, 1);
  constexpr uint64_t dbase = make_smem_desc_base(16384, 1024);  // MN-major: next 64-channel chunk 128 rows * 128 B further
  const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
  uint32_t phase = 0;

  for (int tile = t0; tile < t1; ++tile) {
    patch_store<CH>(g, s_patch, pre);
    dense_store();
    __syncthreads();
    if (tile + 1 < t1) {
      patch_fetch<CH>(p, g, tile + 1, pre);
      dense_fetch(tile + 1);
    }
    build_col_row<CH, GP_COMP_NONE>(g, p.Wi, s_patch, tid, b_addr, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 128 / kUmmaK; ++k)  // K = the tile's 128 pixels
        umma_bf16(tmem_base, smem_desc(dbase, a_addr + k * (kUmmaK * 128)), smem_desc(dbase, b_addr + k * (kUmmaK * 128)), idesc,
                  (tile != t0 || k != 0) ? 1u : 0u);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);  // the MMAs have read both tiles: they may be overwritten
    phase ^= 1;
  }

  if (t0 < t1) {
    tc_fence_after();
    const int m = tid;  // accumulator row = dense channel
#pragma unroll 1
    for (int c = 0; c < (ncols + 31) / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, r);
      tmem_ld_wait_regs(r);
      if (m < M) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = c * 32 + j;
          if (col < CH * 16) atomicAdd(p.dw + (long long)m * (CH * 16) + col, __uint_as_float(r[j]));
          else if (want_db && col == CH * 16) atomicAdd(p.dbias + m, __uint_as_float(r[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// ------------------------------------------------------------------------------------------------ transposed direction
// img[n, c, oh, ow] = act(bias[c] + sum_{(ih,kh): 2ih-1+kh = oh} sum_{(iw,kw): 2iw-1+kw = ow} col[(n,ih,iw)][(c*4+kh)*4+kw]),
// col[px][j] = sum_ci x[px][ci] * w[ci][j]           (G's last ConvTranspose2d + Tanh; the image gradient of D's first conv)
// One CTA tile = 128 pixels of the small grid (R rows) plus one halo row above and below: (R + 2) * Ws <= 256 rows =
// two M = 128 MMAs per K step into two 64-column TMEM accumulators; the fp32 column values never leave the SM — they
// are exchanged through shared memory one image channel at a time (16 columns per pixel) and summed into the 2R output
// rows this tile owns. x: K-major A tile copied with cp.async (16-byte chunks, SWIZZLE_128B by address arithmetic);
// w: its torch layout [C][ch*16] IS the MN-major B tile (one 128-byte row per input channel).
struct ImageConvTParams {
  const uint16_t* x;     // 2-byte elements: bf16 (hi) or fp16
  const uint16_t* x_lo;  // bf16x3: low halves
  const float* w;        // fp32 [C][48]
  const float* bias;     // fp32 [3] or null
  float* img;            // fp32 NCHW (NB, 3, 2Hs, 2Ws)
  int act;
  int NB, Hs, Ws, C;
  int tiles;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// 32 lanes x 16 columns of fp32
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

constexpr int kColPitch = 17;  // fp32 words per pixel in the exchange buffer (16 columns of one channel + 1: no bank conflicts)

// shared memory: [A hi 2 x 16 KB | A lo 2 x 16 KB (bf16x3) | B hi 8 KB | B lo 8 KB (bf16x3) | exchange 256 x 17 fp32 | barrier]
template <int FMT>
__global__ void __launch_bounds__(kEdgeThreads, 3) image_convt_fwd_kernel(const __grid_constant__ ImageConvTParams p) {
  constexpr bool X3 = FMT == GP_COMP_LO;
  constexpr int CH = 3;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + (X3 ? 2 : 1) * 32768;
  float* s_col = reinterpret_cast<float*>(sB + (X3 ? 2 : 1) * 8192);
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_col + 256 * kColPitch);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int Hs = p.Hs, Ws = p.Ws, C = p.C;
  const int R = 128 / Ws;
  const int live = (R + 2) * Ws;  // pixel rows of the A tile in use
  const int cpr = C / 8;
  const int Ho = 2 * Hs, Wo = 2 * Ws;

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  // zero the A tiles once (rows past `live` feed accumulator rows nobody reads, but keep them finite)
  for (int i = tid; i < (X3 ? 2 : 1) * 32768 / 16; i += kEdgeThreads) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0u, 0u, 0u, 0u);
  // weights: row ci = 48 columns = 6 groups, swizzled by the row index (MN-major B: K rows of 128 bytes)
  for (int ci = tid; ci < C; ci += kEdgeThreads) {
#pragma unroll
    for (int gi = 0; gi < 2 * CH; ++gi) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __ldg(p.w + (long long)ci * (CH * 16) + gi * 8 + j);
      const Packed8c pk = pack8c(FMT, f);
      uint8_t* d = sB + ci * 128 + ((gi ^ (ci & 7)) << 4);
      *reinterpret_cast<uint4*>(d) = FMT == GP_COMP_F16 ? pk.comp : pk.hi;
      if (X3) *reinterpret_cast<uint4*>(d + 8192) = pk.comp;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t idesc = make_idesc_bf16(kBlockM, CH * 16, 0, 1) & ~(FMT == GP_COMP_F16 ? ((1u << 7) | (1u << 10)) : 0u);
  constexpr uint64_t da_base = make_smem_desc_base(0, 1024);     // K-major A
  constexpr uint64_t db_base = make_smem_desc_base(8192, 1024);  // MN-major B (one 64-column chunk: LBO unused)
  const uint32_t a_hi = smem_u32(sA), a_lo = a_hi + 32768, b_hi = smem_u32(sB), b_lo = b_hi + 8192;
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    const long long pix0 = (long long)tile * 128;
    const int n = (int)(pix0 / (Hs * Ws));
    const int ih0 = (int)(pix0 % (Hs * Ws)) / Ws;
    // ---- A tile: rows ih0 - 1 .. ih0 + R of image n (zero outside the image)
    for (int i = tid; i < live * cpr; i += kEdgeThreads) {
      const int lp = i / cpr, j = i % cpr;
      const int ih = ih0 - 1 + lp / Ws, iw = lp % Ws;
      const uint32_t off = (uint32_t)(lp >> 7) * 16384u + (uint32_t)(lp & 127) * 128u + (uint32_t)((j ^ (lp & 7)) << 4);
      if (ih >= 0 && ih < Hs) {
        const long long src = (((long long)n * Hs + ih) * Ws + iw) * C + j * 8;
        cp_async16(a_hi + off, p.x + src);
        if (X3) cp_async16(a_lo + off, p.x_lo + src);
      } else {
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a_hi + off), "r"(0u) : "memory");
        if (X3) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a_lo + off), "r"(0u) : "memory");
      }
    }
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      for (int mi = 0; mi < 2; ++mi) {
        for (int k = 0; k < C / kUmmaK; ++k) {
          const uint64_t ah = smem_desc(da_base, a_hi + mi * 16384 + k * 32);
          const uint64_t bh = smem_desc(db_base, b_hi + k * (kUmmaK * 128));
          umma_bf16(tmem_base + mi * 64, ah, bh, idesc, k != 0);
          if (X3) {
            umma_bf16(tmem_base + mi * 64, smem_desc(da_base, a_lo + mi * 16384 + k * 32), bh, idesc, 1u);
            umma_bf16(tmem_base + mi * 64, ah, smem_desc(db_base, b_lo + k * (kUmmaK * 128)), idesc, 1u);
          }
        }
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- col2im, one image channel at a time through the exchange buffer
#pragma unroll 1
    for (int c = 0; c < CH; ++c) {
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        uint32_t r[16];
        tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + mi * 64 + c * 16, r);
        const int lp = mi * 128 + tid;
        if (lp < live) {
#pragma unroll
          for (int j = 0; j < 16; ++j) s_col[lp * kColPitch + j] = __uint_as_float(r[j]);
        }
      }
      if (c == CH - 1) tc_fence_before();
      __syncthreads();
      const float b0 = p.bias != nullptr ? __ldg(p.bias + c) : 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // 2R x 2Ws = 512 outputs per channel
        const int idx = k * kEdgeThreads + tid;
        const int ol = idx / Wo, ow = idx % Wo;
        const int oh = 2 * ih0 + ol;
        const int kh0 = (oh + 1) & 1, kw0 = (ow + 1) & 1;
        float acc = b0;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int kh = kh0 + 2 * a;
          const int lr = (oh + 1 - kh) / 2 - (ih0 - 1);  // local row 0 .. R + 1 (rows outside the image hold zeros)
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const int kw = kw0 + 2 * b;
            const int iw2 = ow + 1 - kw;  // = 2 * iw
            if (iw2 >= 0 && iw2 < Wo) acc += s_col[(lr * Ws + (iw2 >> 1)) * kColPitch + kh * 4 + kw];
          }
        }
        p.img[(((long long)n * CH + c) * Ho + oh) * Wo + ow] = act_fwd(acc, p.act);
      }
      __syncthreads();  // the exchange buffer is rewritten for the next channel / TMEM and A for the next tile
    }
  }

  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

static bool edge_geometry_ok(int ch, int Hi, int Wi, int C) {
  if (ch != 3 || Hi <= 0 || Wi <= 0 || (Hi & 1) || (Wi & 3)) return false;
  const int Ho = Hi / 2, Wo = Wi / 2;
  if (Wo > 128 || 128 % Wo != 0 || (Ho * Wo) % 128 != 0) return false;
  return C > 0 && C % 8 == 0 && C <= 128;
}

static int edge_grid(int tiles, int ctas_per_sm) {
  const long long cap = (long long)num_sms() * ctas_per_sm;
  return (int)(tiles < cap ? tiles : cap);
}

}  // namespace gp

using namespace gp;

extern "C" int gp_image_conv_k4s2_fwd(const float* img, const float* mul, const float* w, const float* bias, void* out,
                                      void* out_comp, int comp_fmt, int NB, int ch, int Hi, int Wi, int Cout, int act,
                                      void* stream) {
  GP_REQUIRE(img && w && out && NB > 0, "gp_image_conv_k4s2_fwd: null pointer / empty batch");
  GP_REQUIRE(edge_geometry_ok(ch, Hi, Wi, Cout),
             "gp_image_conv_k4s2_fwd: unsupported geometry ch=%d %dx%d Cout=%d (ch == 3, Wo | 128, Ho*Wo %% 128 == 0, Cout %% 8 == 0 <= 128)",
             ch, Hi, Wi, Cout);
  GP_REQUIRE(comp_fmt >= GP_COMP_NONE && comp_fmt <= GP_COMP_F16, "gp_image_conv_k4s2_fwd: unknown companion format %d", comp_fmt);
  GP_REQUIRE(comp_fmt == GP_COMP_NONE || out_comp != nullptr, "gp_image_conv_k4s2_fwd: companion format %d without out_comp", comp_fmt);
  GP_REQUIRE(act == GP_ACT_NONE || act == GP_ACT_RELU || act == GP_ACT_LRELU, "gp_image_conv_k4s2_fwd: activation %d not supported", act);
  ImageEdgeParams p;
  memset(&p, 0, sizeof(p));
  p.img = img, p.mul = mul, p.w = w, p.bias = bias;
  p.out = static_cast<__nv_bfloat16*>(out), p.out_comp = out_comp, p.comp_fmt = comp_fmt;
  p.slope = act == GP_ACT_NONE ? 1.f : (act == GP_ACT_RELU ? 0.f : 0.2f);
  p.NB = NB, p.Hi = Hi, p.Wi = Wi, p.N = Cout;
  p.tiles = (int)((long long)NB * (Hi / 2) * (Wi / 2) / 128);
  const int Npad = (Cout + 15) & ~15;
  p.tmem_cols = Npad <= 32 ? 32 : (Npad <= 64 ? 64 : 128);
  const int halves = comp_fmt == GP_COMP_LO ? 2 : 1;
  const int smem = 1024 + halves * (16384 + Npad * 128) + kEdgePatchFloats * 4 + 160 * 4 + 8192 + 64;
  const int per_sm = (220 * 1024) / smem < 512 / p.tmem_cols ? (220 * 1024) / smem : 512 / p.tmem_cols;
  const int grid = edge_grid(p.tiles, per_sm > 6 ? 6 : per_sm);
  auto launch = [&](auto kfn) -> int {
    GP_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kfn<<<grid, kEdgeThreads, smem, as_stream(stream)>>>(p);
    GP_CHECK_LAUNCH();
    return 0;
  };
  switch (comp_fmt) {
    case GP_COMP_LO: return launch(image_conv_fwd_kernel<3, GP_COMP_LO>);
    case GP_COMP_F16: return launch(image_conv_fwd_kernel<3, GP_COMP_F16>);
    default: return launch(image_conv_fwd_kernel<3, GP_COMP_NONE>);
  }
}

extern "C" int gp_image_conv_k4s2_wgrad(const void* dense, const float* img, const float* mul, float* dw, float* dbias,
                                        int NB, int ch, int Hi, int Wi, int M, void* stream) {
  GP_REQUIRE(dense && img && dw && NB > 0, "gp_image_conv_k4s2_wgrad: null pointer / empty batch");
  GP_REQUIRE(edge_geometry_ok(ch, Hi, Wi, M),
             "gp_image_conv_k4s2_wgrad: unsupported geometry ch=%d %dx%d M=%d (ch == 3, Wo | 128, Ho*Wo %% 128 == 0, M %% 8 == 0 <= 128)",
             ch, Hi, Wi, M);
  ImageEdgeParams p;
  memset(&p, 0, sizeof(p));
  p.img = img, p.mul = mul, p.dense = static_cast<const __nv_bfloat16*>(dense), p.dw = dw, p.dbias = dbias;
  p.NB = NB, p.Hi = Hi, p.Wi = Wi, p.N = M;
  p.tiles = (int)((long long)NB * (Hi / 2) * (Wi / 2) / 128);
  p.tmem_cols = 64;
  const int smem = 1024 + 32768 + 16384 + kEdgePatchFloats * 4 + 64;
  const int grid = edge_grid(p.tiles, 3);
  auto kfn = image_conv_wgrad_kernel<3>;
  GP_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kfn<<<grid, kEdgeThreads, smem, as_stream(stream)>>>(p);
  GP_CHECK_LAUNCH();
  return 0;
}

extern "C" int gp_image_convt_k4s2_fwd(const void* x, const void* x_lo, int fmt, const float* w, const float* bias, float* img,
                                       int NB, int Hs, int Ws, int C, int ch, int act, void* stream) {
  GP_REQUIRE(x && w && img && NB > 0, "gp_image_convt_k4s2_fwd: null pointer / empty batch");
  GP_REQUIRE(fmt >= GP_COMP_NONE && fmt <= GP_COMP_F16, "gp_image_convt_k4s2_fwd: unknown operand format %d", fmt);
  GP_REQUIRE(fmt != GP_COMP_LO || x_lo != nullptr, "gp_image_convt_k4s2_fwd: bf16x3 operands need x_lo");
  GP_REQUIRE(ch == 3 && Hs > 0 && Ws > 0 && Ws <= 64 && 128 % Ws == 0 && (Hs * Ws) % 128 == 0 && C % 16 == 0 && C > 0 && C <= 64,
             "gp_image_convt_k4s2_fwd: unsupported geometry ch=%d %dx%d C=%d (ch == 3, Ws | 128 <= 64, Hs*Ws %% 128 == 0, C %% 16 == 0 <= 64)",
             ch, Hs, Ws, C);
  ImageConvTParams p;
  memset(&p, 0, sizeof(p));
  p.x = static_cast<const uint16_t*>(x), p.x_lo = static_cast<const uint16_t*>(x_lo), p.w = w, p.bias = bias, p.img = img;
  p.act = act, p.NB = NB, p.Hs = Hs, p.Ws = Ws, p.C = C;
  p.tiles = (int)((long long)NB * Hs * Ws / 128);
  const int halves = fmt == GP_COMP_LO ? 2 : 1;
  const int smem = 1024 + halves * (32768 + 8192) + 256 * kColPitch * 4 + 64;
  const int per_sm = (220 * 1024) / smem < 3 ? (220 * 1024) / smem : 3;
  const int grid = edge_grid(p.tiles, per_sm);
  auto launch = [&](auto kfn) -> int {
    GP_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kfn<<<grid, kEdgeThreads, smem, as_stream(stream)>>>(p);
    GP_CHECK_LAUNCH();
    return 0;
  };
  switch (fmt) {
    case GP_COMP_LO: return launch(image_convt_fwd_kernel<GP_COMP_LO>);
    case GP_COMP_F16: return launch(image_convt_fwd_kernel<GP_COMP_F16>);
    default: return launch(image_convt_fwd_kernel<GP_COMP_NONE>);
  }
}
