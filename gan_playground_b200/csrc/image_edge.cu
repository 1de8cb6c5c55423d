// Fused image-edge layers: the 4x4 stride-2 convolution between the fp32 NCHW image and the first / last NHWC
// activation (D's first Conv2d, models/dcgan.py:106-109; G's last ConvTranspose2d + Tanh, models/dcgan.py:41-44, and the
// gradient side of both) WITHOUT a column buffer in HBM.
//
//   gp_image_conv_k4s2_fwd  : out[px][n] = act(bias[n] + sum_j col[px][j] * w[n][j])      col tile = K-major A operand
//   gp_image_conv_k4s2_wgrad: dw[m][j] += sum_px dense[px][m] * col[px][j]                col tile = MN-major B operand
//   gp_image_convt_k4s2_fwd : img = act(bias[c] + col2im(x * w))                          col values stay in TMEM / registers
//   col[px = (n, oh, ow)][j = (c*4 + kh)*4 + kw] = img[n, c, 2oh-1+kh, 2ow-1+kw] * (mul ? 1 - mul[same]^2 : 1)
//
// All three are HBM-bound (K = 48); what they must avoid is instruction work per byte, so every bulk transfer is a TMA
// copy and the threads only convert:
//   * the image rows of a 128-pixel tile arrive as ONE cp.async.bulk.tensor.3d box (Wi + 8 columns from column -4, 2R + 2
//     rows from row 2*oh0 - 1, 3 channels): padding = TMA out-of-bounds zero fill, so the tile builder has no bounds checks —
//     per pixel 36 loads, fp32 -> bf16 (or hi/lo pair, or fp16), 6 x STS.128 into the 128-byte-swizzled UMMA layout;
//   * forward: the bias rides in the MMA (column 48/49 of the K = 64 tile are the constant 1, the weight tile holds the bias
//     split over them), the epilogue is tcgen05.ld -> LeakyReLU -> pack -> swizzled shared memory -> per-warp TMA STORE
//     (cp.async.bulk.tensor.2d.global.shared::cta, a ring of two 4 KB slots per warp);
//   * wgrad: dy tiles arrive by TMA straight in the MN-major operand layout, 3-4 stages deep, one persistent CTA per SM
//     accumulating dW (and, through a constant-1 column, the bias gradient) in TMEM, flushed once with vector reductions;
//   * transposed direction: x tiles (128 pixels + one halo row above and below, in the second M = 128 half) arrive by TMA,
//     the 48 column values of a pixel are read from TMEM by the thread that owns the pixel, horizontal neighbours come by
//     warp shuffle, vertical neighbours through 9 KB of shared memory, and every thread writes its 2x2x3 output patch.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "act_io.cuh"
#include "bn_stream.cuh"
#include "conv_gemm.cuh"

namespace gp {

constexpr int kEdgeThreads = 128;

// ------------------------------------------------------------------------------------------------ tensor maps (host)
static PFN_cuTensorMapEncodeTiled_v12000 edge_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// fp32 NCHW image seen as (Wi, Hi, ch * NB); box (Wi + 8, rows, ch), no swizzle. Loaded from (-4, 2*oh0 - 1, n*ch): the
// innermost start coordinate of a tiled TMA load must be a multiple of 16 bytes (measured on the B200 with
// tools/probe/tma_probe.cu: start -1 is an illegal instruction, -4 loads with the left columns zero-filled).
static int make_map_image(CUtensorMap* m, const float* base, int NB, int ch, int Hi, int Wi, int rows) {
  auto fn = edge_encode_fn();
  if (fn == nullptr) return set_error(GP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(GP_ERR_INVALID, "image base not 16-byte aligned");
  cuuint64_t dims[3] = {(cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)ch * NB};
  cuuint64_t strides[2] = {(cuuint64_t)Wi * 4, (cuuint64_t)Hi * Wi * 4};
  cuuint32_t box[3] = {(cuuint32_t)(Wi + 8), (cuuint32_t)rows, (cuuint32_t)ch};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(GP_ERR_CUDA, "cuTensorMapEncodeTiled(image %dx%dx%dx%d, %d rows) failed: %d", NB, ch, Hi, Wi, rows, (int)r);
  return GP_OK;
}

// 2-byte row-major matrix [P][C]; box (64 columns, box_rows), 128-byte swizzle (loads and stores)
static int make_map_rows(CUtensorMap* m, const void* base, long long P, int C, int box_rows) {
  auto fn = edge_encode_fn();
  if (fn == nullptr) return set_error(GP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(GP_ERR_INVALID, "activation base not 16-byte aligned");
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)P};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(GP_ERR_CUDA, "cuTensorMapEncodeTiled(rows P=%lld C=%d) failed: %d", P, C, (int)r);
  return GP_OK;
}

// 2-byte NHWC (C, W, H, N); box (64 channels, W, bh rows, 1 image), 128-byte swizzle; rows outside the image are zero
static int make_map_nhwc_rows(CUtensorMap* m, const void* base, int C, int W, int H, int N, int bh) {
  auto fn = edge_encode_fn();
  if (fn == nullptr) return set_error(GP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(GP_ERR_INVALID, "activation base not 16-byte aligned");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)W, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(GP_ERR_CUDA, "cuTensorMapEncodeTiled(nhwc C=%d W=%d H=%d N=%d rows=%d) failed: %d", C, W, H, N, bh, (int)r);
  return GP_OK;
}

// ------------------------------------------------------------------------------------------------ shared device pieces
// geometry of a 128-pixel tile: R = 128 / Wo whole output rows of one image; its image patch is [3][2R + 2][Wi + 8] fp32,
// patch column q = image column q - 4
struct EdgeGeom {
  int Wi, Wo, R, tiles, tiles_per_img;
  int pitch_b;       // bytes of one patch row: (Wi + 8) * 4
  int chan_b;        // bytes of one channel of the patch: (2R + 2) * pitch_b
  int patch_bytes;   // 3 * chan_b: the TMA transaction size
  int patch_stride;  // patch_bytes rounded up to 128
};

static EdgeGeom make_geom(int NB, int Hi, int Wi) {
  EdgeGeom g;
  g.Wi = Wi, g.Wo = Wi / 2;
  g.R = 128 / g.Wo;
  g.tiles_per_img = (Hi / 2) * g.Wo / 128;
  g.tiles = NB * g.tiles_per_img;
  g.pitch_b = (Wi + 8) * 4;
  g.chan_b = (2 * g.R + 2) * g.pitch_b;
  g.patch_bytes = 3 * g.chan_b;
  g.patch_stride = (g.patch_bytes + 127) & ~127;
  return g;
}

__device__ __forceinline__ float2 lds_f32x2(uint32_t saddr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_u32x4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts_u32x4(uint32_t saddr, const uint4& v) { sts_u32x4(saddr, v.x, v.y, v.z, v.w); }

// Bounded barrier wait that names itself: with GP_EDGE_DEBUG=1 the library maps a small pinned host buffer into every
// launch and a wait that times out (1 s) records (code, block, thread, parity, aux) there before trapping.
__device__ __noinline__ void edge_wait_timeout(unsigned int* dbg, uint32_t code, uint32_t parity, uint32_t aux) {
  if (dbg != nullptr) {
    dbg[1] = blockIdx.x, dbg[2] = threadIdx.x, dbg[3] = parity, dbg[4] = aux;
    __threadfence_system();
    dbg[0] = code;
    __threadfence_system();
  }
  asm volatile("trap;");
}
__device__ __forceinline__ void edge_wait(uint64_t* bar, uint32_t parity, unsigned int* dbg, uint32_t code, uint32_t aux = 0) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 1000000000ull) edge_wait_timeout(dbg, code, parity, aux);
  }
}

// issue the patch load(s) of one tile; the caller has armed `bar` with the transaction bytes
__device__ __forceinline__ void patch_load(const EdgeGeom& g, const CUtensorMap* map, uint32_t dst, uint64_t* bar, int tile) {
  const int n = tile / g.tiles_per_img;
  const int oh0 = (tile - n * g.tiles_per_img) * g.R;
  tma_load_3d(dst, map, bar, -4, 2 * oh0 - 1, n * 3);
}

// Row r (one output pixel) of the column tile: six 16-byte groups of 8 columns, group gi at byte
// r * 128 + ((gi ^ (r & 7)) << 4) — the SWIZZLE_128B pattern of a {64 elements, 128 rows} box. Groups 6 and 7 (columns
// 48..63) are written once by the kernels (constants) and never touched here.
// FMT: GP_COMP_NONE -> bf16 into t_hi; GP_COMP_LO -> bf16 hi into t_hi and bf16(v - hi) into t_lo; GP_COMP_F16 -> fp16 into t_hi.
// row_off: byte offset of patch element (row 2*ol, column 2*ow + 3) = image (2*(oh0+ol) - 1, 2*ow - 1); the four taps of a
// row are read as 4 + 8 + 4 bytes (the middle pair is 8-byte aligned).
template <int FMT, bool MUL>
__device__ __forceinline__ void build_col_row(const EdgeGeom& g, uint32_t patch, uint32_t mpatch, uint32_t row_off, int r,
                                              uint32_t t_hi, uint32_t t_lo) {
#pragma unroll
  for (int gi = 0; gi < 6; ++gi) {
    const uint32_t rel = (uint32_t)(gi >> 1) * g.chan_b + (uint32_t)((gi & 1) * 2) * g.pitch_b + row_off;
    float f[8];
    {
      const uint32_t a = patch + rel, a2 = a + g.pitch_b;
      const float2 m0 = lds_f32x2(a + 4), m1 = lds_f32x2(a2 + 4);
      f[0] = lds_f32(a), f[1] = m0.x, f[2] = m0.y, f[3] = lds_f32(a + 12);
      f[4] = lds_f32(a2), f[5] = m1.x, f[6] = m1.y, f[7] = lds_f32(a2 + 12);
    }
    if (MUL) {
      const uint32_t a = mpatch + rel, a2 = a + g.pitch_b;
      const float2 m0 = lds_f32x2(a + 4), m1 = lds_f32x2(a2 + 4);
      const float t[8] = {lds_f32(a), m0.x, m0.y, lds_f32(a + 12), lds_f32(a2), m1.x, m1.y, lds_f32(a2 + 12)};
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= fmaf(-t[j], t[j], 1.f);
    }
    const Packed8c pk = pack8c(FMT, f);
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((gi ^ (r & 7)) << 4);
    if (FMT == GP_COMP_F16) {
      sts_u32x4(t_hi + off, pk.comp);
    } else {
      sts_u32x4(t_hi + off, pk.hi);
      if (FMT == GP_COMP_LO) sts_u32x4(t_lo + off, pk.comp);
    }
  }
}

__device__ __forceinline__ uint32_t bf16_bits(float v) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float bf16_val(uint32_t b) { return __uint_as_float(b << 16); }

// ------------------------------------------------------------------------------------------------ forward
struct EdgeFwdParams {
  CUtensorMap map_img, map_mul, map_out, map_comp;
  const float* w;     // fp32 [N][48] (torch layout of Conv2d / ConvTranspose2d weights with the image on the 3-channel side)
  const float* bias;  // fp32 [N] or null
  float slope;
  int N;  // 64 or 128
  EdgeGeom g;
  unsigned int* dbg;
};

// shared memory: [A hi 16 KB | A lo 16 KB (bf16x3) | B hi N*128 | B lo (bf16x3) | store ring 4 warps x 2 x 4 KB |
//                 patch (| mul patch) | 2 barriers | TMEM slot]
template <int FMT, bool MUL>
__global__ void __launch_bounds__(kEdgeThreads) image_conv_fwd_kernel(const __grid_constant__ EdgeFwdParams p) {
  gp::pdl_sync();
  constexpr bool X3 = FMT == GP_COMP_LO;
  constexpr bool COMP = FMT != GP_COMP_NONE;
  constexpr int HALVES = X3 ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const EdgeGeom g = p.g;
  const int N = p.N;
  const int bbytes = N * 128;
  uint8_t* sA = smem;
  uint8_t* sB = sA + HALVES * 16384;
  uint8_t* sStage = sB + HALVES * bbytes;
  uint8_t* sPatch = sStage + 32768;
  uint64_t* bar_patch = reinterpret_cast<uint64_t*>(sPatch + (MUL ? 2 : 1) * g.patch_stride);
  uint64_t* bar_mma = bar_patch + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_hi = smem_u32(sA), a_lo = a_hi + 16384, b_hi = smem_u32(sB), b_lo = b_hi + (uint32_t)bbytes;
  const uint32_t patch = smem_u32(sPatch), mpatch = patch + (uint32_t)g.patch_stride;
  const uint32_t tx_bytes = (uint32_t)g.patch_bytes * (MUL ? 2u : 1u);

  if (tid == 0) {
    tma_prefetch_desc(&p.map_img);
    tma_prefetch_desc(&p.map_out);
    if (MUL) tma_prefetch_desc(&p.map_mul);
    if (COMP) tma_prefetch_desc(&p.map_comp);
    mbar_init(bar_patch, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, (uint32_t)N);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  int tile = blockIdx.x;
  if (tid == 0 && tile < g.tiles) {
    mbar_expect_tx(bar_patch, tx_bytes);
    patch_load(g, &p.map_img, patch, bar_patch, tile);
    if (MUL) patch_load(g, &p.map_mul, mpatch, bar_patch, tile);
  }

  // constant columns 48..63 of the A tile: (1, 1, 0, ...) in the hi half — they multiply the bias columns of B
  {
    const uint32_t one2 = FMT == GP_COMP_F16 ? 0x3C003C00u : 0x3F803F80u;
    const uint32_t o6 = (uint32_t)tid * 128u + (uint32_t)((6 ^ (tid & 7)) << 4), o7 = (uint32_t)tid * 128u + (uint32_t)((7 ^ (tid & 7)) << 4);
    sts_u32x4(a_hi + o6, one2, 0u, 0u, 0u);
    sts_u32x4(a_hi + o7, 0u, 0u, 0u, 0u);
    if (X3) {
      sts_u32x4(a_lo + o6, 0u, 0u, 0u, 0u);
      sts_u32x4(a_lo + o7, 0u, 0u, 0u, 0u);
    }
  }
  // weights -> B tile(s): row n = 128 bytes, swizzled like the A rows; columns 48 / 49 carry the bias split into two
  // (bf16x3: three) terms, so that 1 * b1 + 1 * b2 (+ 1 * b3) reproduces it to 16 (24) bits in the fp32 accumulator
  for (int n = tid; n < N; n += kEdgeThreads) {
    const uint32_t rowb = (uint32_t)n * 128u;
#pragma unroll
    for (int gi = 0; gi < 6; ++gi) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w + (long long)n * 48 + gi * 8));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w + (long long)n * 48 + gi * 8 + 4));
      const float f[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const Packed8c pk = pack8c(FMT, f);
      const uint32_t off = rowb + (uint32_t)((gi ^ (n & 7)) << 4);
      sts_u32x4(b_hi + off, FMT == GP_COMP_F16 ? pk.comp : pk.hi);
      if (X3) sts_u32x4(b_lo + off, pk.comp);
    }
    const float b = p.bias != nullptr ? __ldg(p.bias + n) : 0.f;
    uint32_t w_hi, w_lo = 0u;
    if (FMT == GP_COMP_F16) {
      const __half h1 = __float2half_rn(b);
      const __half h2 = __float2half_rn(b - __half2float(h1));
      w_hi = (uint32_t)__half_as_ushort(h1) | ((uint32_t)__half_as_ushort(h2) << 16);
    } else {
      const uint32_t b1 = bf16_bits(b);
      const uint32_t b2 = bf16_bits(b - bf16_val(b1));
      w_hi = b1 | (b2 << 16);
      if (X3) w_lo = bf16_bits(b - bf16_val(b1) - bf16_val(b2));
    }
    const uint32_t o6 = rowb + (uint32_t)((6 ^ (n & 7)) << 4), o7 = rowb + (uint32_t)((7 ^ (n & 7)) << 4);
    sts_u32x4(b_hi + o6, w_hi, 0u, 0u, 0u);
    sts_u32x4(b_hi + o7, 0u, 0u, 0u, 0u);
    if (X3) {
      sts_u32x4(b_lo + o6, w_lo, 0u, 0u, 0u);
      sts_u32x4(b_lo + o7, 0u, 0u, 0u, 0u);
    }
  }

  const uint32_t idesc = make_idesc_bf16(kBlockM, N, 0, 0) & ~(FMT == GP_COMP_F16 ? ((1u << 7) | (1u << 10)) : 0u);
  constexpr uint64_t dbase = make_smem_desc_base(0, 1024);
  const float slope = p.slope;
  const int nh = N / 64;
  const uint32_t row_off = (uint32_t)(2 * (tid / g.Wo)) * g.pitch_b + (uint32_t)(2 * (tid % g.Wo) + 3) * 4u;
  const uint32_t stage_w = smem_u32(sStage) + (uint32_t)warp * 8192u;
  const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t srow = (uint32_t)lane * 128u;
  const int sw = lane & 7;
  uint32_t ph_patch = 0, ph_mma = 0, sidx = 0;

  for (; tile < g.tiles; tile += gridDim.x) {
    edge_wait(bar_patch, ph_patch, p.dbg, 1, (uint32_t)tile);
    ph_patch ^= 1;
    build_col_row<FMT, MUL>(g, patch, mpatch, row_off, tid, a_hi, a_lo);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();  // the A tile (and, the first time, B) is complete; the patch is consumed; TMEM has been drained
    if (tid == 0) {
      const int next = tile + gridDim.x;
      if (next < g.tiles) {  // lands during the MMA and the epilogue
        mbar_expect_tx(bar_patch, tx_bytes);
        patch_load(g, &p.map_img, patch, bar_patch, next);
        if (MUL) patch_load(g, &p.map_mul, mpatch, bar_patch, next);
      }
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // K = 64: 48 image taps + the bias columns
        const uint64_t ah = smem_desc(dbase, a_hi + k * 32), bh = smem_desc(dbase, b_hi + k * 32);
        umma_bf16(tmem_base, ah, bh, idesc, k != 0);
        if (X3) {
          umma_bf16(tmem_base, smem_desc(dbase, a_lo + k * 32), bh, idesc, 1u);
          umma_bf16(tmem_base, ah, smem_desc(dbase, b_lo + k * 32), idesc, 1u);
        }
      }
      umma_commit(bar_mma);
    }
    edge_wait(bar_mma, ph_mma, p.dbg, 2, (uint32_t)tile);
    ph_mma ^= 1;
    tc_fence_after();
    // epilogue: thread = pixel row; 64 columns at a time -> this warp's 32-row box in shared memory -> TMA store
    const int row0 = tile * 128 + warp * 32;
#pragma unroll 1
    for (int h = 0; h < nh; ++h) {
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(trow + h * 64, ra);
      tmem_ld_32x32(trow + h * 64 + 32, rb);
      tmem_ld_wait_regs(ra);
      tmem_ld_wait_regs(rb);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float xa = __uint_as_float(ra[j]), xb = __uint_as_float(rb[j]);
        ra[j] = __float_as_uint(fmaxf(xa, slope * xa));
        rb[j] = __float_as_uint(fmaxf(xb, slope * xb));
      }
      uint32_t hw[32];  // bf16 pairs of the 64 columns
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const __nv_bfloat162 pa = __floats2bfloat162_rn(__uint_as_float(ra[2 * e]), __uint_as_float(ra[2 * e + 1]));
        const __nv_bfloat162 pb = __floats2bfloat162_rn(__uint_as_float(rb[2 * e]), __uint_as_float(rb[2 * e + 1]));
        hw[e] = *reinterpret_cast<const uint32_t*>(&pa);
        hw[16 + e] = *reinterpret_cast<const uint32_t*>(&pb);
      }
      {
        const uint32_t slot = stage_w + (sidx & 1u) * 4096u;
        ++sidx;
        if (lane == 0) bulk_wait_read<1>();  // the store that used this slot two groups ago has read it
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q) sts_u32x4(slot + srow + (uint32_t)((q ^ sw) << 4), hw[4 * q], hw[4 * q + 1], hw[4 * q + 2], hw[4 * q + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&p.map_out, slot, h * 64, row0);
          bulk_commit();
        }
      }
      if (COMP) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          if (FMT == GP_COMP_F16) {
            const __half2 pa = __floats2half2_rn(__uint_as_float(ra[2 * e]), __uint_as_float(ra[2 * e + 1]));
            const __half2 pb = __floats2half2_rn(__uint_as_float(rb[2 * e]), __uint_as_float(rb[2 * e + 1]));
            hw[e] = *reinterpret_cast<const uint32_t*>(&pa);
            hw[16 + e] = *reinterpret_cast<const uint32_t*>(&pb);
          } else {
            const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[e]));
            const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[16 + e]));
            const __nv_bfloat162 pa = __floats2bfloat162_rn(__uint_as_float(ra[2 * e]) - fa.x, __uint_as_float(ra[2 * e + 1]) - fa.y);
            const __nv_bfloat162 pb = __floats2bfloat162_rn(__uint_as_float(rb[2 * e]) - fb.x, __uint_as_float(rb[2 * e + 1]) - fb.y);
            hw[e] = *reinterpret_cast<const uint32_t*>(&pa);
            hw[16 + e] = *reinterpret_cast<const uint32_t*>(&pb);
          }
        }
        const uint32_t slot = stage_w + (sidx & 1u) * 4096u;
        ++sidx;
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q) sts_u32x4(slot + srow + (uint32_t)((q ^ sw) << 4), hw[4 * q], hw[4 * q + 1], hw[4 * q + 2], hw[4 * q + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&p.map_comp, slot, h * 64, row0);
          bulk_commit();
        }
      }
    }
  }

  if (lane == 0) bulk_wait_read<0>();  // the ring must outlive the stores that read it
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)N);
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
struct EdgeWgradParams {
  CUtensorMap map_img, map_mul, map_dense;
  float* dw;     // fp32 [M][48], accumulated
  float* dbias;  // fp32 [M] or null, accumulated (column sums of dense)
  int M;         // 64 or 128
  EdgeGeom g;
  unsigned int* dbg;
};

// shared memory: [STAGES x dense tile (MN-major A: 2 x 128 pixel rows x 128 B = 32 KB) | 2 x column tile (B, 16 KB) |
//                 STAGES x patch (| mul patch) | barriers | TMEM slot]
// One persistent CTA per SM walks a contiguous range of tiles: TMA keeps STAGES - 1 tiles in flight, the 128 threads only
// build the column tile, one thread issues the 8 MMAs (K = the tile's 128 pixels) into ONE 128 x 64 fp32 accumulator in
// TMEM, which is added to dw once at the end. Column 48 of the column tile is the constant 1 when a bias gradient is
// wanted, so accumulator column 48 = sum_px dense.
template <bool MUL, int STAGES>
__global__ void __launch_bounds__(kEdgeThreads, 1) image_conv_wgrad_kernel(const __grid_constant__ EdgeWgradParams p) {
  gp::pdl_sync();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const EdgeGeom g = p.g;
  constexpr int NP = MUL ? 2 : 1;
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * 32768;
  uint8_t* sPatch = sB + 2 * 16384;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sPatch + STAGES * NP * g.patch_stride);
  uint64_t* bar_done = bar_full + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int M = p.M;
  const int nchunk = M / 64;
  const bool want_db = p.dbias != nullptr;

  // contiguous range of tiles for this CTA
  const int per = (g.tiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per;
  const int n = max(0, min(t0 + per, g.tiles) - t0);

  if (tid == 0) {
    tma_prefetch_desc(&p.map_img);
    tma_prefetch_desc(&p.map_dense);
    if (MUL) tma_prefetch_desc(&p.map_mul);
    for (int s = 0; s < STAGES; ++s) mbar_init(bar_full + s, 1);
    mbar_init(bar_done, 1);
    mbar_init(bar_done + 1, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  // dense channels beyond M stay zero; columns 48..63 of both column tiles are constants
  if (nchunk < 2)
    for (int s = 0; s < STAGES; ++s)
      for (int i = tid; i < 16384 / 16; i += kEdgeThreads) reinterpret_cast<uint4*>(sA + s * 32768 + 16384)[i] = make_uint4(0u, 0u, 0u, 0u);
  const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB), patch0 = smem_u32(sPatch);
  {
    const uint32_t o6 = (uint32_t)tid * 128u + (uint32_t)((6 ^ (tid & 7)) << 4), o7 = (uint32_t)tid * 128u + (uint32_t)((7 ^ (tid & 7)) << 4);
    for (int b = 0; b < 2; ++b) {
      sts_u32x4(b_addr + b * 16384 + o6, want_db ? 0x00003F80u : 0u, 0u, 0u, 0u);
      sts_u32x4(b_addr + b * 16384 + o7, 0u, 0u, 0u, 0u);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t tx_bytes = (uint32_t)nchunk * 16384u + (uint32_t)g.patch_bytes * NP;
  auto issue_loads = [&](int i) {  // local tile i -> stage i % STAGES
    const int s = i % STAGES, tile = t0 + i;
    mbar_expect_tx(bar_full + s, tx_bytes);
    for (int c = 0; c < nchunk; ++c) tma_load_2d(sA + s * 32768 + c * 16384, &p.map_dense, bar_full + s, c * 64, tile * 128);
    const uint32_t pd = patch0 + (uint32_t)(s * NP) * g.patch_stride;
    patch_load(g, &p.map_img, pd, bar_full + s, tile);
    if (MUL) patch_load(g, &p.map_mul, pd + g.patch_stride, bar_full + s, tile);
  };
  if (tid == 0)
    for (int i = 0; i < STAGES - 1 && i < n; ++i) issue_loads(i);

  const uint32_t idesc = make_idesc_bf16(kBlockM, 64, 1, 1);
  constexpr uint64_t dbase = make_smem_desc_base(16384, 1024);  // MN-major: next 64-channel chunk 128 rows * 128 B further
  const uint32_t row_off = (uint32_t)(2 * (tid / g.Wo)) * g.pitch_b + (uint32_t)(2 * (tid % g.Wo) + 3) * 4u;

  for (int i = 0; i < n; ++i) {
    const int s = i % STAGES, b = i & 1;
    edge_wait(bar_full + s, (uint32_t)(i / STAGES) & 1u, p.dbg, 3, (uint32_t)i);
    if (i >= 2) edge_wait(bar_done + b, (uint32_t)((i - 2) >> 1) & 1u, p.dbg, 4, (uint32_t)i);  // the MMAs of tile i - 2 have read this column tile
    const uint32_t pd = patch0 + (uint32_t)(s * NP) * g.patch_stride;
    build_col_row<GP_COMP_NONE, MUL>(g, pd, pd + g.patch_stride, row_off, tid, b_addr + b * 16384, 0u);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t sa = a_addr + s * 32768, sb = b_addr + b * 16384;
#pragma unroll
      for (int k = 0; k < 128 / kUmmaK; ++k)
        umma_bf16(tmem_base, smem_desc(dbase, sa + k * (kUmmaK * 128)), smem_desc(dbase, sb + k * (kUmmaK * 128)), idesc,
                  (i != 0 || k != 0) ? 1u : 0u);
      umma_commit(bar_done + b);
      const int nxt = i + STAGES - 1;  // its stage was last read by the MMAs of tile i - 1
      if (nxt < n) {
        if (i >= 1) edge_wait(bar_done + (b ^ 1), (uint32_t)((i - 1) >> 1) & 1u, p.dbg, 5, (uint32_t)i);
        issue_loads(nxt);
      }
    }
  }

  if (n > 0) {
    edge_wait(bar_done + ((n - 1) & 1), (uint32_t)((n - 1) >> 1) & 1u, p.dbg, 6, (uint32_t)n);
    tc_fence_after();
    uint32_t ra[32], rb[32];
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    tmem_ld_32x32(trow, ra);
    tmem_ld_32x32(trow + 32, rb);
    tmem_ld_wait_regs(ra);
    tmem_ld_wait_regs(rb);
    if (tid < M) {  // accumulator row = dense channel
      float* drow = p.dw + (long long)tid * 48;
#pragma unroll
      for (int q = 0; q < 8; ++q) red_add_v4(drow + 4 * q, ra[4 * q], ra[4 * q + 1], ra[4 * q + 2], ra[4 * q + 3]);
#pragma unroll
      for (int q = 0; q < 4; ++q) red_add_v4(drow + 32 + 4 * q, rb[4 * q], rb[4 * q + 1], rb[4 * q + 2], rb[4 * q + 3]);
      if (want_db) atomicAdd(p.dbias + tid, __uint_as_float(rb[16]));
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
// One tile = 128 pixels of the small grid (R whole rows, accumulator half 0) plus the row above and the row below
// (2 * Ws rows of accumulator half 1). Thread t owns pixel t: it reads the pixel's 48 column values from TMEM, forms the
// horizontal pair sums with its lane neighbours (shuffles), publishes the kh = 0 / kh = 3 partial sums its vertical
// neighbours need, and after one barrier writes the 2 x 2 x 3 output values of its pixel.
struct EdgeConvTParams {
  CUtensorMap map_x, map_x_halo, map_lo, map_lo_halo;
  const float* w;     // fp32 [C][48]
  const float* bias;  // fp32 [3] or null
  float* img;         // fp32 NCHW (NB, 3, 2Hs, 2Ws)
  int act;
  int NB, Hs, Ws, C;
  int tiles, tiles_per_img, R;
  unsigned int* dbg;
};

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

constexpr int kExRows = 192;  // exchange rows: 128 owned pixels + 2 * Ws halo pixels, Ws <= 32

// shared memory: [A: C/64 main chunks (128 pixel rows x 128 B = 16 KB each), then C/64 halo chunks (2 * Ws rows each) |
//                 B hi C*128 | B lo (bf16x3) | exchange 12 x 192 fp32 | barriers | TMEM slot]
// The halo MMAs (M = 128) read 16 KB from the start of a 2*Ws-row halo chunk: the rows past it belong to whatever follows
// in shared memory and only feed accumulator rows nobody reads. 74 KB for C = 128 (bf16 / fp16): three CTAs share an SM,
// so one CTA's loads and MMAs run under another's epilogue; within a CTA the next tile's loads fly during the epilogue.
// bf16x3 streams the hi halves and then the lo halves of a tile through the SAME operand tile (hi * w_hi + hi * w_lo, then
// lo * w_hi): 90 KB, two CTAs per SM.
template <int FMT>
__global__ void __launch_bounds__(kEdgeThreads) image_convt_fwd_kernel(const __grid_constant__ EdgeConvTParams p) {
  gp::pdl_sync();
  constexpr bool X3 = FMT == GP_COMP_LO;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int C = p.C, Ws = p.Ws, Hs = p.Hs, R = p.R;
  const int kch = C / 64;
  const uint32_t halo_bytes = (uint32_t)(2 * Ws) * 128u;                  // one halo chunk: the row above, then the row below
  const uint32_t half_bytes = (uint32_t)kch * (16384u + halo_bytes);     // one operand tile (hi or lo)
  uint8_t* sA = smem;
  uint8_t* sB = sA + half_bytes;
  float* s_ex = reinterpret_cast<float*>(sB + (X3 ? 2 : 1) * C * 128);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(s_ex + 12 * kExRows);
  uint64_t* bar_acc = bar_full + 1;
  uint64_t* bar_mid = bar_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t a_addr = smem_u32(sA), b_hi = smem_u32(sB), b_lo = b_hi + (uint32_t)C * 128u;

  if (tid == 0) {
    tma_prefetch_desc(&p.map_x);
    tma_prefetch_desc(&p.map_x_halo);
    if (X3) {
      tma_prefetch_desc(&p.map_lo);
      tma_prefetch_desc(&p.map_lo_halo);
    }
    mbar_init(bar_full, 1);
    mbar_init(bar_acc, 1);
    mbar_init(bar_mid, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  // weights: row ci = 48 columns = 6 groups, swizzled by the row index (MN-major B: K rows of 128 bytes)
  for (int ci = tid; ci < C; ci += kEdgeThreads) {
    const uint32_t rowb = (uint32_t)ci * 128u;
#pragma unroll
    for (int gi = 0; gi < 8; ++gi) {
      const uint32_t off = rowb + (uint32_t)((gi ^ (ci & 7)) << 4);
      if (gi < 6) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.w + (long long)ci * 48 + gi * 8));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.w + (long long)ci * 48 + gi * 8 + 4));
        const float f[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const Packed8c pk = pack8c(FMT, f);
        sts_u32x4(b_hi + off, FMT == GP_COMP_F16 ? pk.comp : pk.hi);
        if (X3) sts_u32x4(b_lo + off, pk.comp);
      } else {
        sts_u32x4(b_hi + off, 0u, 0u, 0u, 0u);
        if (X3) sts_u32x4(b_lo + off, 0u, 0u, 0u, 0u);
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // this CTA's tiles: blockIdx.x, blockIdx.x + gridDim.x, ...; local index i
  const int n_local = p.tiles > (int)blockIdx.x ? (p.tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto issue_loads = [&](int i, int part) {  // part 0: the bf16 / fp16 operand (bf16x3: hi halves); part 1: the lo halves
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int nimg = tile / p.tiles_per_img, ih0 = (tile - nimg * p.tiles_per_img) * R;
    const CUtensorMap* mm = part == 0 ? &p.map_x : &p.map_lo;
    const CUtensorMap* mh = part == 0 ? &p.map_x_halo : &p.map_lo_halo;
    mbar_expect_tx(bar_full, half_bytes);
    for (int kc = 0; kc < kch; ++kc) {
      uint8_t* dm = sA + kc * 16384;                       // 128 owned pixels
      uint8_t* dh = sA + kch * 16384 + kc * halo_bytes;    // the rows above / below the tile (zero outside the image)
      tma_load_4d(dm, mm, bar_full, kc * 64, 0, ih0, nimg);
      tma_load_4d(dh, mh, bar_full, kc * 64, 0, ih0 - 1, nimg);
      tma_load_4d(dh + Ws * 128, mh, bar_full, kc * 64, 0, ih0 + R, nimg);
    }
  };
  const uint32_t idesc = make_idesc_bf16(kBlockM, 48, 0, 1) & ~(FMT == GP_COMP_F16 ? ((1u << 7) | (1u << 10)) : 0u);
  constexpr uint64_t da_base = make_smem_desc_base(0, 1024);     // K-major A
  constexpr uint64_t db_base = make_smem_desc_base(8192, 1024);  // MN-major B (one 64-column chunk: LBO unused)
  // accumulator columns 0..47: the owned pixels, 64..111: the halo rows. part 0: a * w_hi (+ a * w_lo for bf16x3), zeroing
  // the accumulator; part 1 (bf16x3): a_lo * w_hi on top
  auto mma_block = [&](int part) {
    for (int mh = 0; mh < 2; ++mh)
      for (int kc = 0; kc < kch; ++kc) {
        const uint32_t sa = a_addr + (mh == 0 ? (uint32_t)kc * 16384u : (uint32_t)kch * 16384u + (uint32_t)kc * halo_bytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = smem_desc(da_base, sa + k * 32);
          umma_bf16(tmem_base + mh * 64, ad, smem_desc(db_base, b_hi + (kc * 4 + k) * (kUmmaK * 128)), idesc,
                    (part != 0 || (kc | k) != 0) ? 1u : 0u);
          if (X3 && part == 0) umma_bf16(tmem_base + mh * 64, ad, smem_desc(db_base, b_lo + (kc * 4 + k) * (kUmmaK * 128)), idesc, 1u);
        }
      }
  };
  auto issue_mmas = [&](int i) {  // by ONE thread; the part-0 loads of tile i have been issued
    edge_wait(bar_full, X3 ? 0u : ((uint32_t)i & 1u), p.dbg, 7, (uint32_t)i);
    tc_fence_after();
    mma_block(0);
    if (X3) {
      umma_commit(bar_mid);
      edge_wait(bar_mid, (uint32_t)i & 1u, p.dbg, 9, (uint32_t)i);  // the hi halves have been consumed
      issue_loads(i, 1);
      edge_wait(bar_full, 1u, p.dbg, 10, (uint32_t)i);
      tc_fence_after();
      mma_block(1);
    }
    umma_commit(bar_acc);
  };

  if (tid == 0 && n_local > 0) issue_loads(0, 0);
  if (tid == 32 && n_local > 0) issue_mmas(0);

  const int lr = tid / Ws, iw = tid - lr * Ws;
  const bool left_edge = iw == 0, right_edge = iw == Ws - 1;
  const bool halo_warp = warp * 32 < 2 * Ws;  // its threads also own one pixel of the rows above / below the tile
  const bool halo_bottom = tid >= Ws;
  const int Ho = 2 * Hs, Wo = 2 * Ws;
  float bias3[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) bias3[c] = p.bias != nullptr ? __ldg(p.bias + c) : 0.f;
  // exchange arrays: [kind (0: kh = 3, for the row below; 1: kh = 0, for the row above)][c][b][192]; slot rows: owned pixel t at
  // Ws + t, the row above the tile at 0 .. Ws - 1, the row below at Ws + 128 ..
  const uint32_t ex = smem_u32(s_ex);
  auto ex_at = [&](int kind, int c, int b, int pos) -> uint32_t { return ex + (uint32_t)((((kind * 3 + c) * 2 + b) * kExRows + pos) * 4); };
  const uint32_t acc = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

  for (int i = 0; i < n_local; ++i) {
    edge_wait(bar_acc, (uint32_t)i & 1u, p.dbg, 8, (uint32_t)i);
    tc_fence_after();
    if (tid == 0 && i + 1 < n_local) issue_loads(i + 1, 0);  // the MMAs of tile i are done: the operand tile is free

    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int nimg = tile / p.tiles_per_img, ih0 = (tile - nimg * p.tiles_per_img) * R;
    float H[3][2][2];  // [c][a: output row 2ih + a][b: output column 2iw + b], own-row terms (kh = 1 / kh = 2)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint32_t v[16];
      tmem_ld_32x16(acc + c * 16, v);
      // horizontal pair sums for every kh: ow = 2iw <- kw = 1 (own) + kw = 3 (pixel iw - 1); ow = 2iw + 1 <- kw = 2 (own) + kw = 0 (pixel iw + 1)
      float hs[4][2];
#pragma unroll
      for (int kh = 0; kh < 4; ++kh) {
        const float from_left = __shfl_up_sync(0xffffffffu, __uint_as_float(v[kh * 4 + 3]), 1);
        const float from_right = __shfl_down_sync(0xffffffffu, __uint_as_float(v[kh * 4 + 0]), 1);
        hs[kh][0] = __uint_as_float(v[kh * 4 + 1]) + (left_edge ? 0.f : from_left);
        hs[kh][1] = __uint_as_float(v[kh * 4 + 2]) + (right_edge ? 0.f : from_right);
      }
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        H[c][0][b] = hs[1][b];
        H[c][1][b] = hs[2][b];
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(ex_at(0, c, b, Ws + tid)), "f"(hs[3][b]) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(ex_at(1, c, b, Ws + tid)), "f"(hs[0][b]) : "memory");
      }
    }
    if (halo_warp) {  // warp-uniform: pixel tid of the halo half (row above: tid < Ws, row below: Ws <= tid < 2 Ws)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint32_t v[16];
        tmem_ld_32x16(acc + 64 + c * 16, v);
        // the row above contributes its kh = 3 sums, the row below its kh = 0 sums
        const float k1 = halo_bottom ? __uint_as_float(v[1]) : __uint_as_float(v[13]);
        const float k2 = halo_bottom ? __uint_as_float(v[2]) : __uint_as_float(v[14]);
        const float k3 = halo_bottom ? __uint_as_float(v[3]) : __uint_as_float(v[15]);
        const float k0 = halo_bottom ? __uint_as_float(v[0]) : __uint_as_float(v[12]);
        const float from_left = __shfl_up_sync(0xffffffffu, k3, 1);
        const float from_right = __shfl_down_sync(0xffffffffu, k0, 1);
        const float s0 = k1 + (left_edge ? 0.f : from_left), s1 = k2 + (right_edge ? 0.f : from_right);
        if (tid < 2 * Ws) {
          const int pos = halo_bottom ? 128 + tid : tid;  // (R + 1) * Ws + iw = 128 + tid for the row below
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(ex_at(halo_bottom ? 1 : 0, c, 0, pos)), "f"(s0) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(ex_at(halo_bottom ? 1 : 0, c, 1, pos)), "f"(s1) : "memory");
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // partial sums published; the accumulator is drained
    if (tid == 32 && i + 1 < n_local) issue_mmas(i + 1);  // the loads had the first half of the epilogue to land
    {
      const int oh = 2 * (ih0 + lr);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float o[2][2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          float up, dn;  // kh = 3 sums of the row above (slot row lr), kh = 0 sums of the row below (slot row lr + 2)
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(up) : "r"(ex_at(0, c, b, tid)) : "memory");
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(dn) : "r"(ex_at(1, c, b, 2 * Ws + tid)) : "memory");
          o[0][b] = act_fwd(H[c][0][b] + up + bias3[c], p.act);
          o[1][b] = act_fwd(H[c][1][b] + dn + bias3[c], p.act);
        }
        float* dst = p.img + (((long long)nimg * 3 + c) * Ho + oh) * Wo + 2 * iw;
        *reinterpret_cast<float2*>(dst) = make_float2(o[0][0], o[0][1]);
        *reinterpret_cast<float2*>(dst + Wo) = make_float2(o[1][0], o[1][1]);
      }
    }
    __syncthreads();  // the exchange arrays are rewritten by the next tile
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

static bool edge_geometry_ok(int ch, int Hi, int Wi, int C) {
  if (ch != 3 || Hi <= 0 || Wi <= 0 || (Hi & 1) || (Wi & 3)) return false;
  const int Ho = Hi / 2, Wo = Wi / 2;
  if (Wo > 64 || 128 % Wo != 0 || (Ho * Wo) % 128 != 0) return false;
  return C == 64 || C == 128;
}

// GP_EDGE_DEBUG=1: every launch is followed by a stream synchronisation; a failure prints what the timed-out wait recorded
static unsigned int* edge_debug_buffer() {
  static unsigned int* buf = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* e = getenv("GP_EDGE_DEBUG");
    if (e != nullptr && e[0] == '1' && cudaHostAlloc(&buf, 64, cudaHostAllocMapped) == cudaSuccess) memset(buf, 0, 64);
    else buf = nullptr;
  }
  return buf;
}
static int edge_debug_check(const char* what, void* stream) {
  unsigned int* buf = edge_debug_buffer();
  if (buf == nullptr) return GP_OK;
  const cudaError_t e = cudaStreamSynchronize(as_stream(stream));
  if (e != cudaSuccess || buf[0] != 0) {
    fprintf(stderr, "[gp edge debug] %s: %s; wait code %u block %u thread %u parity %u aux %u\n", what, cudaGetErrorString(e),
            buf[0], buf[1], buf[2], buf[3], buf[4]);
    return set_error(GP_ERR_CUDA, "%s failed: %s (wait code %u block %u thread %u parity %u aux %u)", what, cudaGetErrorString(e),
                     buf[0], buf[1], buf[2], buf[3], buf[4]);
  }
  return GP_OK;
}

// resident CTAs per SM by shared memory (228 KB per SM, 1 KB reserved per CTA) and registers
template <typename K>
static int edge_occupancy(K kfn, int smem) {
  cudaFuncAttributes attr;
  int by_regs = 16;
  if (cudaFuncGetAttributes(&attr, kfn) == cudaSuccess && attr.numRegs > 0) by_regs = 65536 / (attr.numRegs * kEdgeThreads);
  const int by_smem = (228 * 1024) / (smem + 1024);
  const int occ = by_regs < by_smem ? by_regs : by_smem;
  return occ < 1 ? 1 : occ;
}

}  // namespace gp

using namespace gp;

extern "C" int gp_image_conv_k4s2_fwd(const float* img, const float* mul, const float* w, const float* bias, void* out,
                                      void* out_comp, int comp_fmt, int NB, int ch, int Hi, int Wi, int Cout, int act,
                                      void* stream) {
  GP_REQUIRE(img && w && out && NB > 0, "gp_image_conv_k4s2_fwd: null pointer / empty batch");
  GP_REQUIRE(edge_geometry_ok(ch, Hi, Wi, Cout),
             "gp_image_conv_k4s2_fwd: unsupported geometry ch=%d %dx%d Cout=%d (ch == 3, Wo | 128 <= 64, Ho*Wo %% 128 == 0, Cout 64 or 128)",
             ch, Hi, Wi, Cout);
  GP_REQUIRE(comp_fmt >= GP_COMP_NONE && comp_fmt <= GP_COMP_F16, "gp_image_conv_k4s2_fwd: unknown companion format %d", comp_fmt);
  GP_REQUIRE(comp_fmt == GP_COMP_NONE || out_comp != nullptr, "gp_image_conv_k4s2_fwd: companion format %d without out_comp", comp_fmt);
  GP_REQUIRE(mul == nullptr || comp_fmt == GP_COMP_NONE, "gp_image_conv_k4s2_fwd: the tanh' factor is a backward-side (bf16) feature");
  GP_REQUIRE(act == GP_ACT_NONE || act == GP_ACT_RELU || act == GP_ACT_LRELU, "gp_image_conv_k4s2_fwd: activation %d not supported", act);
  GP_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, "gp_image_conv_k4s2_fwd: weights not 16-byte aligned");
  EdgeFwdParams p;
  memset(&p, 0, sizeof(p));
  p.g = make_geom(NB, Hi, Wi);
  p.w = w, p.bias = bias, p.N = Cout;
  p.dbg = edge_debug_buffer();
  p.slope = act == GP_ACT_NONE ? 1.f : (act == GP_ACT_RELU ? 0.f : 0.2f);
  const long long P = (long long)p.g.tiles * 128;
  int rc = make_map_image(&p.map_img, img, NB, ch, Hi, Wi, 2 * p.g.R + 2);
  if (rc == GP_OK && mul != nullptr) rc = make_map_image(&p.map_mul, mul, NB, ch, Hi, Wi, 2 * p.g.R + 2);
  if (rc == GP_OK) rc = make_map_rows(&p.map_out, out, P, Cout, 32);
  if (rc == GP_OK && comp_fmt != GP_COMP_NONE) rc = make_map_rows(&p.map_comp, out_comp, P, Cout, 32);
  if (rc != GP_OK) return rc;
  const int halves = comp_fmt == GP_COMP_LO ? 2 : 1;
  const int smem = 1024 + halves * (16384 + Cout * 128) + 32768 + (mul != nullptr ? 2 : 1) * p.g.patch_stride + 64;
  auto launch = [&](auto kfn) -> int {
    GP_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = edge_occupancy(kfn, smem);
    if (per_sm > 512 / Cout) per_sm = 512 / Cout;  // TMEM columns
    const long long cap = (long long)num_sms() * per_sm;
    const int grid = (int)(p.g.tiles < cap ? p.g.tiles : cap);
    gp::launch_pdl(kfn, grid, kEdgeThreads, smem, as_stream(stream), p);
    GP_CHECK_LAUNCH();
    return edge_debug_check(__func__, stream);
  };
  if (mul != nullptr) return launch(image_conv_fwd_kernel<GP_COMP_NONE, true>);
  switch (comp_fmt) {
    case GP_COMP_LO: return launch(image_conv_fwd_kernel<GP_COMP_LO, false>);
    case GP_COMP_F16: return launch(image_conv_fwd_kernel<GP_COMP_F16, false>);
    default: return launch(image_conv_fwd_kernel<GP_COMP_NONE, false>);
  }
}

extern "C" int gp_image_conv_k4s2_wgrad(const void* dense, const float* img, const float* mul, float* dw, float* dbias,
                                        int NB, int ch, int Hi, int Wi, int M, void* stream) {
  GP_REQUIRE(dense && img && dw && NB > 0, "gp_image_conv_k4s2_wgrad: null pointer / empty batch");
  GP_REQUIRE(edge_geometry_ok(ch, Hi, Wi, M),
             "gp_image_conv_k4s2_wgrad: unsupported geometry ch=%d %dx%d M=%d (ch == 3, Wo | 128 <= 64, Ho*Wo %% 128 == 0, M 64 or 128)",
             ch, Hi, Wi, M);
  GP_REQUIRE((reinterpret_cast<uintptr_t>(dw) & 15) == 0, "gp_image_conv_k4s2_wgrad: dw not 16-byte aligned");
  EdgeWgradParams p;
  memset(&p, 0, sizeof(p));
  p.g = make_geom(NB, Hi, Wi);
  p.dw = dw, p.dbias = dbias, p.M = M;
  p.dbg = edge_debug_buffer();
  const long long P = (long long)p.g.tiles * 128;
  int rc = make_map_image(&p.map_img, img, NB, ch, Hi, Wi, 2 * p.g.R + 2);
  if (rc == GP_OK && mul != nullptr) rc = make_map_image(&p.map_mul, mul, NB, ch, Hi, Wi, 2 * p.g.R + 2);
  if (rc == GP_OK) rc = make_map_rows(&p.map_dense, dense, P, M, 128);
  if (rc != GP_OK) return rc;
  const int grid = p.g.tiles < num_sms() ? p.g.tiles : num_sms();
  auto launch = [&](auto kfn, int stages) -> int {
    const int smem = 1024 + stages * 32768 + 2 * 16384 + stages * (mul != nullptr ? 2 : 1) * p.g.patch_stride + 128;
    GP_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    gp::launch_pdl(kfn, grid, kEdgeThreads, smem, as_stream(stream), p);
    GP_CHECK_LAUNCH();
    return edge_debug_check(__func__, stream);
  };
  if (mul != nullptr) return launch(image_conv_wgrad_kernel<true, 3>, 3);
  return launch(image_conv_wgrad_kernel<false, 4>, 4);
}

extern "C" int gp_image_convt_k4s2_fwd(const void* x, const void* x_lo, int fmt, const float* w, const float* bias, float* img,
                                       int NB, int Hs, int Ws, int C, int ch, int act, void* stream) {
  GP_REQUIRE(x && w && img && NB > 0, "gp_image_convt_k4s2_fwd: null pointer / empty batch");
  GP_REQUIRE(fmt >= GP_COMP_NONE && fmt <= GP_COMP_F16, "gp_image_convt_k4s2_fwd: unknown operand format %d", fmt);
  GP_REQUIRE(fmt != GP_COMP_LO || x_lo != nullptr, "gp_image_convt_k4s2_fwd: bf16x3 operands need x_lo");
  GP_REQUIRE(ch == 3 && Hs > 0 && (Ws == 16 || Ws == 32) && (Hs * Ws) % 128 == 0 && (C == 64 || C == 128),
             "gp_image_convt_k4s2_fwd: unsupported geometry ch=%d %dx%d C=%d (ch == 3, Ws 16 or 32, Hs*Ws %% 128 == 0, C 64 or 128)",
             ch, Hs, Ws, C);
  GP_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(img) & 7) == 0,
             "gp_image_convt_k4s2_fwd: weights / image not aligned");
  EdgeConvTParams p;
  memset(&p, 0, sizeof(p));
  p.w = w, p.bias = bias, p.img = img, p.act = act, p.NB = NB, p.Hs = Hs, p.Ws = Ws, p.C = C;
  p.R = 128 / Ws;
  p.dbg = edge_debug_buffer();
  p.tiles_per_img = Hs * Ws / 128;
  p.tiles = NB * p.tiles_per_img;
  int rc = make_map_nhwc_rows(&p.map_x, x, C, Ws, Hs, NB, p.R);
  if (rc == GP_OK) rc = make_map_nhwc_rows(&p.map_x_halo, x, C, Ws, Hs, NB, 1);
  if (rc == GP_OK && fmt == GP_COMP_LO) rc = make_map_nhwc_rows(&p.map_lo, x_lo, C, Ws, Hs, NB, p.R);
  if (rc == GP_OK && fmt == GP_COMP_LO) rc = make_map_nhwc_rows(&p.map_lo_halo, x_lo, C, Ws, Hs, NB, 1);
  if (rc != GP_OK) return rc;
  const bool x3 = fmt == GP_COMP_LO;
  const int half = (C / 64) * (16384 + 2 * Ws * 128);
  const int smem = 1024 + half + (x3 ? 2 : 1) * C * 128 + 12 * kExRows * 4 + 128;
  auto launch = [&](auto kfn) -> int {
    GP_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = edge_occupancy(kfn, smem);
    if (per_sm > 4) per_sm = 4;  // TMEM: 128 columns per CTA
    const long long cap = (long long)num_sms() * per_sm;
    const int grid = (int)(p.tiles < cap ? p.tiles : cap);
    gp::launch_pdl(kfn, grid, kEdgeThreads, smem, as_stream(stream), p);
    GP_CHECK_LAUNCH();
    return edge_debug_check(__func__, stream);
  };
  switch (fmt) {
    case GP_COMP_LO: return launch(image_convt_fwd_kernel<GP_COMP_LO>);
    case GP_COMP_F16: return launch(image_convt_fwd_kernel<GP_COMP_F16>);
    default: return launch(image_convt_fwd_kernel<GP_COMP_NONE>);
  }
}
