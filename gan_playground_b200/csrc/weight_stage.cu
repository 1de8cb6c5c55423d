// Batched weight staging: every 4x4 conv / conv-transpose weight of one network, in every GEMM-operand format and
// orientation the step needs, from ONE launch that reads each fp32 parameter once.
//
// The per-tensor entry points (gp_pack_conv_weight / gp_split_conv_weight + gp_pair_to_f16) cost one launch per layer,
// orientation and format — ~30 launches per DCGAN step whose bytes do not shrink with the batch, i.e. 15-20 % of the step at
// 128 images per GPU. Here a descriptor table lists the layers (src (D0, D1, 16) fp32) and, per layer, up to four
// destinations {pointer, orientation n_dim, format}; a block takes a 32 x 32 x 16 tile of a source through shared memory
// (contiguous 2 KB reads per d0) and writes it to every destination of its layer in 16-byte vectors:
//   n_dim 0: dst[d0][t*D1 + d1]  (N = D0, C = D1: Conv2d forward, ConvTranspose2d data gradient)
//   n_dim 1: dst[d1][t*D0 + d0]  (N = D1, C = D0: ConvTranspose2d forward, Conv2d data gradient)
//   format: bf16 | fp16 | bf16 hi and lo blocks per row (hi | lo, the bf16x3 operand).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstring>

#include "common.h"

namespace gp {

constexpr int kStageP1 = 17;             // floats per (d0, d1): 16 taps + 1
constexpr int kStageP0 = 32 * 17 + 1;    // floats per d0 (32 d1 + 1): both transposed read patterns are conflict-free
constexpr int kStageSmem = 32 * kStageP0 * 4;

__device__ __forceinline__ void stage_store8(const gp_stage_dst_t& d, long long off, long long lo_off, const float (&f)[8]) {
  if (d.fmt == GP_STAGE_F16) {
    uint4 o;
    __half2* ph = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) ph[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(static_cast<__half*>(d.ptr) + off) = o;
    return;
  }
  uint4 o;
  __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) ph[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  __nv_bfloat16* base = static_cast<__nv_bfloat16*>(d.ptr);
  *reinterpret_cast<uint4*>(base + off) = o;
  if (d.fmt == GP_STAGE_SPLIT) {
    uint4 l;
    __nv_bfloat162* pl = reinterpret_cast<__nv_bfloat162*>(&l);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 hf = __bfloat1622float2(ph[i]);
      pl[i] = __floats2bfloat162_rn(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
    }
    *reinterpret_cast<uint4*>(base + off + lo_off) = l;
  }
}

__global__ void __launch_bounds__(256) stage_conv16_kernel(const __grid_constant__ gp_stage_table_t tb) {
  gp::pdl_sync();
  extern __shared__ float s_tile[];  // [32 d0][32 d1][16 t], pitches kStageP0 / kStageP1
  const int tid = threadIdx.x;
  for (int tile = blockIdx.x; tile < tb.total_tiles; tile += gridDim.x) {
    int li = 0;
    while (li + 1 < tb.count && tile >= tb.layer[li + 1].tile0) ++li;
    const gp_stage_layer_t& L = tb.layer[li];
    const int lt = tile - L.tile0;
    const int nb1 = L.D1 / 32;
    const int b0 = lt / nb1, b1 = lt - b0 * nb1;
    // ---- load: for each d0 of the tile 32 d1 x 16 taps = 512 contiguous floats
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
      const int i = tid + k * 256;        // float4 index over [32 d0][128 float4]
      const int d0 = i >> 7, q = i & 127;  // q: d1 = q >> 2, taps 4 * (q & 3) ..
      const float4 v = __ldg(reinterpret_cast<const float4*>(L.src + ((long long)(b0 * 32 + d0) * L.D1 + b1 * 32) * 16) + q);
      float* d = s_tile + d0 * kStageP0 + (q >> 2) * kStageP1 + (q & 3) * 4;
      d[0] = v.x, d[1] = v.y, d[2] = v.z, d[3] = v.w;
    }
    __syncthreads();
    // ---- write every destination of the layer: 2048 items of 8 consecutive C elements each
    for (int di = 0; di < L.ndst; ++di) {
      const gp_stage_dst_t& d = L.dst[di];
#pragma unroll 2
      for (int k = 0; k < 8; ++k) {
        const int i = tid + k * 256;
        const int g = i & 3, t = (i >> 2) & 15, r = i >> 6;  // r: the N index inside the tile
        float f[8];
        long long off, lo_off;
        if (d.n_dim == 0) {  // N = d0 = r, C = d1 = g*8 + j
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = s_tile[r * kStageP0 + (g * 8 + j) * kStageP1 + t];
          off = (long long)(b0 * 32 + r) * d.ld + (long long)t * L.D1 + b1 * 32 + g * 8;
          lo_off = 16LL * L.D1;
        } else {             // N = d1 = r, C = d0 = g*8 + j
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = s_tile[(g * 8 + j) * kStageP0 + r * kStageP1 + t];
          off = (long long)(b1 * 32 + r) * d.ld + (long long)t * L.D0 + b0 * 32 + g * 8;
          lo_off = 16LL * L.D0;
        }
        stage_store8(d, off, lo_off, f);
      }
    }
    __syncthreads();  // the tile is overwritten by the next one
  }
}

}  // namespace gp

using namespace gp;

extern "C" int gp_stage_conv_weights(const gp_stage_table_t* table, void* stream) {
  GP_REQUIRE(table != nullptr && table->count > 0 && table->count <= GP_STAGE_MAX_LAYERS, "gp_stage_conv_weights: 1..%d layers",
             GP_STAGE_MAX_LAYERS);
  gp_stage_table_t tb;
  memcpy(&tb, table, sizeof(tb));
  int total = 0;
  for (int i = 0; i < tb.count; ++i) {
    gp_stage_layer_t& L = tb.layer[i];
    GP_REQUIRE(L.src != nullptr && L.D0 > 0 && L.D1 > 0 && L.D0 % 32 == 0 && L.D1 % 32 == 0,
               "gp_stage_conv_weights: layer %d: (D0, D1) = (%d, %d) must be multiples of 32 (16-tap weights only)", i, L.D0, L.D1);
    GP_REQUIRE((reinterpret_cast<uintptr_t>(L.src) & 15) == 0, "gp_stage_conv_weights: layer %d: source not 16-byte aligned", i);
    GP_REQUIRE(L.ndst > 0 && L.ndst <= GP_STAGE_MAX_DST, "gp_stage_conv_weights: layer %d: 1..%d destinations", i, GP_STAGE_MAX_DST);
    for (int j = 0; j < L.ndst; ++j) {
      const gp_stage_dst_t& d = L.dst[j];
      const int C = d.n_dim == 0 ? L.D1 : L.D0;
      GP_REQUIRE(d.ptr != nullptr && (d.n_dim == 0 || d.n_dim == 1) && d.fmt >= GP_STAGE_BF16 && d.fmt <= GP_STAGE_SPLIT,
                 "gp_stage_conv_weights: layer %d destination %d: bad descriptor", i, j);
      GP_REQUIRE(d.ld >= (d.fmt == GP_STAGE_SPLIT ? 32LL : 16LL) * C && d.ld % 8 == 0 && (reinterpret_cast<uintptr_t>(d.ptr) & 15) == 0,
                 "gp_stage_conv_weights: layer %d destination %d: row pitch / alignment", i, j);
    }
    L.tile0 = total;
    total += (L.D0 / 32) * (L.D1 / 32);
  }
  tb.total_tiles = total;
  static bool attr_set = false;
  if (!attr_set) {
    GP_CHECK_CUDA(cudaFuncSetAttribute(stage_conv16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageSmem));
    attr_set = true;
  }
  const int cap = num_sms() * 3;
  gp::launch_pdl(stage_conv16_kernel, total < cap ? total : cap, 256, kStageSmem, as_stream(stream), tb);
  GP_CHECK_LAUNCH();
  return 0;
}
