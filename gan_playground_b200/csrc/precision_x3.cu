// Kernels of the "bf16x3" forward precision mode: forward GEMM operands are hi/lo bf16 pairs
// (x = hi + lo to ~16 mantissa bits) and pre-BatchNorm conv outputs are kept in fp32, so that the forward pass
// carries fp32-class activations while every tensor-core instruction stays bf16 (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo).
// These are the HBM-bound producers/consumers of those formats; the GEMM side lives in conv_gemm.cuh (extra K taps).
#include <cuda_bf16.h>

#include "common.h"

namespace gp {
namespace x3 {

__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == GP_ACT_RELU) return fmaxf(v, 0.f);
  if (act == GP_ACT_LRELU) return v > 0.f ? v : 0.2f * v;
  if (act == GP_ACT_TANH) return tanhf(v);
  return v;
}
__device__ __forceinline__ float act_grad(float z, int act) {
  if (act == GP_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == GP_ACT_LRELU) return z > 0.f ? 1.f : 0.2f;
  return 1.f;
}
__device__ __forceinline__ void split(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16(v);
  lo = __float2bfloat16(v - __bfloat162float(hi));
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x, f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void store8_split(__nv_bfloat16* hi, __nv_bfloat16* lo, const float (&f)[8]) {
  uint4 rh, rl;
  __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&rh);
  __nv_bfloat162* pl = reinterpret_cast<__nv_bfloat162*>(&rl);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    const float2 hf = __bfloat1622float2(h2);
    ph[i] = h2;
    pl[i] = __floats2bfloat162_rn(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
  }
  *reinterpret_cast<uint4*>(hi) = rh;
  if (lo != nullptr) *reinterpret_cast<uint4*>(lo) = rl;
}
__device__ __forceinline__ void store8(__nv_bfloat16* dst, const float (&f)[8]) { store8_split(dst, nullptr, f); }

// thread = one 8-channel group x row lane (same layout as elementwise.cu)
struct ColLayout {
  int g, rl, lanes;
  bool active;
};
__device__ __forceinline__ ColLayout col_layout(int C) {
  ColLayout L;
  const int cgs = C / 8;
  if (cgs <= (int)blockDim.x && gridDim.y == 1) {
    L.g = threadIdx.x % cgs;
    L.lanes = blockDim.x / cgs;
    L.rl = threadIdx.x / cgs;
    L.active = L.rl < L.lanes;
  } else {
    L.g = blockIdx.y * blockDim.x + threadIdx.x;
    L.lanes = 1;
    L.rl = 0;
    L.active = L.g < cgs;
  }
  return L;
}
template <int NQ>
__device__ __forceinline__ void col_flush(const ColLayout& L, int C, float (&acc)[NQ][8], float* const (&out)[NQ]) {
  extern __shared__ float s_red[];
  const int cgs = C / 8;
  const bool tiled = !(cgs <= (int)blockDim.x && gridDim.y == 1);
  const int cols = tiled ? blockDim.x * 8 : C;
  const int base = tiled ? blockIdx.y * blockDim.x * 8 : 0;
  for (int i = threadIdx.x; i < NQ * cols; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  if (L.active) {
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&s_red[q * cols + L.g * 8 + i - base], acc[q][i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NQ * cols; i += blockDim.x) {
    const int q = i / cols, c = base + i % cols;
    if (c < C) atomicAdd(out[q] + c, s_red[i]);
  }
}

// ------------------------------------------------------------------------------------------ operand staging
// dst[r*ld + k] = hi(src[map(r)*s_r + k*s_k]), dst[lo_off + r*ld + k] = lo(...); zero padding to [Rpad][ld].
__global__ void split_matrix_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int R, int K,
                                    int Rpad, int ld, int width, long long s_r, long long s_k, int perm,
                                    long long lo_off) {
  const long long total = (long long)Rpad * width;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / width), k = (int)(i % width);
    float v = 0.f;
    if (r < R && k < K) {
      int rs = r;
      if (perm > 1) {
        const int inner = R / perm;
        rs = (r % inner) * perm + r / inner;
      }
      v = __ldg(src + rs * s_r + k * s_k);
    }
    __nv_bfloat16 hi, lo;
    split(v, hi, lo);
    dst[(long long)r * ld + k] = hi;
    dst[lo_off + (long long)r * ld + k] = lo;
  }
}

// conv weight (D0, D1, taps) fp32 -> bf16 [N][2][taps][C]  (hi block, then lo block, per output row)
__global__ void split_conv_weight_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int D0, int D1,
                                         int taps, int n_dim) {
  const int N = n_dim == 0 ? D0 : D1, C = n_dim == 0 ? D1 : D0;
  const long long total = (long long)N * taps * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int t = (int)((i / C) % taps);
    const int n = (int)(i / ((long long)C * taps));
    const int d0 = n_dim == 0 ? n : c, d1 = n_dim == 0 ? c : n;
    __nv_bfloat16 hi, lo;
    split(__ldg(src + ((long long)d0 * D1 + d1) * taps + t), hi, lo);
    const long long row = (long long)n * 2 * taps * C;
    dst[row + (long long)t * C + c] = hi;
    dst[row + (long long)taps * C + (long long)t * C + c] = lo;
  }
}

// ------------------------------------------------------------------------------------------ BatchNorm on fp32 y
__global__ void bn_stats_f32_kernel(const float* __restrict__ y, long long P, int C, float* __restrict__ sum,
                                    float* __restrict__ sumsq, int rows_per_block) {
  const ColLayout L = col_layout(C);
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (L.active) {
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > P) r1 = P;
    for (long long r = r0 + L.rl; r < r1; r += L.lanes) {
      float f[8];
      load8(y + r * C + L.g * 8, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[0][i] += f[i];
        acc[1][i] += f[i] * f[i];
      }
    }
  }
  float* const outs[2] = {sum, sumsq};
  col_flush<2>(L, C, acc, outs);
}

// (hi, lo) = split(act(y * scale + shift))
__global__ void bn_apply_split_kernel(const float* __restrict__ y, __nv_bfloat16* __restrict__ out_hi,
                                      __nv_bfloat16* __restrict__ out_lo, long long P, int C,
                                      const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                      int rows_per_block) {
  const ColLayout L = col_layout(C);
  if (!L.active) return;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[L.g * 8 + j];
    sh[j] = shift[L.g * 8 + j];
  }
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > P) r1 = P;
  for (long long r = r0 + L.rl; r < r1; r += L.lanes) {
    float f[8];
    load8(y + r * C + L.g * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = act_fwd(f[j] * sc[j] + sh[j], act);
    store8_split(out_hi + r * C + L.g * 8, out_lo ? out_lo + r * C + L.g * 8 : nullptr, f);
  }
}

__global__ void bn_bwd_reduce_f32_kernel(const __nv_bfloat16* __restrict__ da, const float* __restrict__ y, long long P,
                                         int C, const float* __restrict__ scale, const float* __restrict__ shift,
                                         const float* __restrict__ mean, const float* __restrict__ rstd, int act,
                                         float* __restrict__ sum_dz, float* __restrict__ sum_dzx, int rows_per_block) {
  const ColLayout L = col_layout(C);
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (L.active) {
    float sc[8], sh[8], mu[8], rs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sc[i] = scale[L.g * 8 + i], sh[i] = shift[L.g * 8 + i], mu[i] = mean[L.g * 8 + i], rs[i] = rstd[L.g * 8 + i];
    }
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > P) r1 = P;
    for (long long r = r0 + L.rl; r < r1; r += L.lanes) {
      float fy[8], fd[8];
      load8(y + r * C + L.g * 8, fy);
      load8(da + r * C + L.g * 8, fd);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dz = fd[i] * act_grad(fy[i] * sc[i] + sh[i], act);
        acc[0][i] += dz;
        acc[1][i] += dz * (fy[i] - mu[i]) * rs[i];
      }
    }
  }
  float* const outs[2] = {sum_dz, sum_dzx};
  col_flush<2>(L, C, acc, outs);
}

__global__ void bn_bwd_apply_f32_kernel(const __nv_bfloat16* __restrict__ da, const float* __restrict__ y,
                                        __nv_bfloat16* __restrict__ dy, long long P, int C,
                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                        const float* __restrict__ sum_dz, const float* __restrict__ sum_dzx,
                                        float inv_count, int act, int rows_per_block) {
  const ColLayout L = col_layout(C);
  if (!L.active) return;
  float sc[8], sh[8], mu[8], k0[8], k1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = L.g * 8 + j;
    sc[j] = scale[c], sh[j] = shift[c], mu[j] = mean[c];
    k0[j] = sc[j] * sum_dz[c] * inv_count;
    k1[j] = sc[j] * rstd[c] * sum_dzx[c] * inv_count;
  }
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > P) r1 = P;
  for (long long r = r0 + L.rl; r < r1; r += L.lanes) {
    float fy[8], fd[8], o[8];
    load8(y + r * C + L.g * 8, fy);
    load8(da + r * C + L.g * 8, fd);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dz = fd[j] * act_grad(fy[j] * sc[j] + sh[j], act);
      o[j] = sc[j] * dz - k0[j] - (fy[j] - mu[j]) * k1[j];
    }
    store8(dy + r * C + L.g * 8, o);
  }
}

// ------------------------------------------------------------------------------------------ image-side layers
__global__ void im2col_k4s2_split_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ col_hi,
                                         __nv_bfloat16* __restrict__ col_lo, int NB, int ch, int Hi, int Wi) {
  const int Ho = Hi / 2, Wo = Wi / 2;
  const long long total = (long long)NB * Ho * Wo * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % 8);
    const long long p = i / 8;
    const int ow = (int)(p % Wo), oh = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col_idx = g * 8 + j;
      const int c = col_idx / 16, kh = (col_idx / 4) % 4, kw = col_idx % 4;
      const int ih = 2 * oh - 1 + kh, iw = 2 * ow - 1 + kw;
      f[j] = (c < ch && ih >= 0 && ih < Hi && iw >= 0 && iw < Wi) ? __ldg(img + (((long long)n * ch + c) * Hi + ih) * Wi + iw) : 0.f;
    }
    store8_split(col_hi + i * 8, col_lo + i * 8, f);
  }
}

__global__ void col2im_k4s2_f32_kernel(const float* __restrict__ col, const float* __restrict__ bias,
                                       float* __restrict__ img, int NB, int ch, int Hi, int Wi, int act) {
  const int Ho = Hi / 2, Wo = Wi / 2;
  const long long total = (long long)NB * ch * Hi * Wi;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int iw = (int)(i % Wi), ih = (int)((i / Wi) % Hi);
    const int c = (int)((i / ((long long)Wi * Hi)) % ch), n = (int)(i / ((long long)Wi * Hi * ch));
    float acc = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int kh = ((ih + 1) & 1) + 2 * a;
      const int oh2 = ih + 1 - kh;
      if (oh2 < 0 || oh2 >= 2 * Ho) continue;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int kw = ((iw + 1) & 1) + 2 * b;
        const int ow2 = iw + 1 - kw;
        if (ow2 < 0 || ow2 >= 2 * Wo) continue;
        acc += col[(((long long)n * Ho + oh2 / 2) * Wo + ow2 / 2) * 64 + (c * 4 + kh) * 4 + kw];
      }
    }
    img[i] = act_fwd(acc, act);
  }
}

// out[b][o] = bias[o] + sum_{hw,c} (a_hi + a_lo)[b,hw,c] * w[o, c, hw]
__global__ void head_fwd_split_kernel(const __nv_bfloat16* __restrict__ a_hi, const __nv_bfloat16* __restrict__ a_lo,
                                      const float* __restrict__ w, const float* __restrict__ bias,
                                      float* __restrict__ out, int HW, int C, int O, long long s_o, long long s_c,
                                      long long s_hw) {
  const int b = blockIdx.x, o = blockIdx.y;
  float acc = 0.f;
  const long long base = (long long)b * HW * C;
  for (int i = threadIdx.x; i < HW * C; i += blockDim.x) {
    const int c = i % C, hw = i / C;
    const float v = __bfloat162float(a_hi[base + i]) + __bfloat162float(a_lo[base + i]);
    acc += v * __ldg(w + o * s_o + c * s_c + hw * s_hw);
  }
  __shared__ float red[32];
  for (int k = 16; k > 0; k >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, k);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int k = 16; k > 0; k >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, k);
    if (threadIdx.x == 0) out[(long long)b * O + o] = acc + (bias ? bias[o] : 0.f);
  }
}

static inline int grid1(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}
struct ColLaunch {
  dim3 grid;
  int block;
  size_t smem;
  int rpb;
};
static ColLaunch col_launch(long long P, int C, int nq) {
  ColLaunch L;
  long long rpb = (P + (long long)num_sms() * 8 - 1) / ((long long)num_sms() * 8);
  if (rpb < 32) rpb = 32;
  L.rpb = (int)rpb;
  const int gx = (int)((P + L.rpb - 1) / L.rpb);
  const int cgs = C / 8;
  if (cgs <= 256 && 256 % cgs == 0) {
    L.grid = dim3(gx, 1);
    L.block = 256;
    L.smem = (size_t)nq * C * sizeof(float);
  } else {
    L.block = 128;
    int gy = (cgs + 127) / 128;
    if (gy < 2) gy = 2;
    L.grid = dim3(gx, gy);
    L.smem = (size_t)nq * 128 * 8 * sizeof(float);
  }
  return L;
}

}  // namespace x3
}  // namespace gp

using namespace gp;
using namespace gp::x3;

extern "C" {

int gp_split_matrix(const float* src, void* dst, int R, int K, int Rpad, int ld, int width, long long s_r, long long s_k,
                    int perm, long long lo_off, void* stream) {
  GP_REQUIRE(src && dst && R > 0 && K > 0 && Rpad >= R && width >= K && ld >= width && lo_off > 0, "gp_split_matrix: bad arguments");
  GP_REQUIRE(perm <= 1 || R % perm == 0, "gp_split_matrix: perm must divide R");
  split_matrix_kernel<<<grid1((long long)Rpad * width), 256, 0, as_stream(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), R, K, Rpad, ld, width, s_r, s_k, perm, lo_off);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_split_conv_weight(const float* src, void* dst, int D0, int D1, int taps, int n_dim, void* stream) {
  GP_REQUIRE(src && dst && D0 > 0 && D1 > 0 && taps > 0 && (n_dim == 0 || n_dim == 1), "gp_split_conv_weight: bad arguments");
  split_conv_weight_kernel<<<grid1((long long)D0 * D1 * taps), 256, 0, as_stream(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), D0, D1, taps, n_dim);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_stats_f32(const float* y, long long P, int C, float* sum, float* sumsq, void* stream) {
  GP_REQUIRE(y && sum && sumsq && P > 0 && C > 0 && C % 8 == 0, "gp_bn_stats_f32: bad arguments");
  const ColLaunch L = col_launch(P, C, 2);
  bn_stats_f32_kernel<<<L.grid, L.block, L.smem, as_stream(stream)>>>(y, P, C, sum, sumsq, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_apply_act_split(const float* y, void* out_hi, void* out_lo, long long P, int C, const float* scale,
                          const float* shift, int act, void* stream) {
  GP_REQUIRE(y && out_hi && scale && shift && P > 0 && C % 8 == 0, "gp_bn_apply_act_split: bad arguments");
  const ColLaunch L = col_launch(P, C, 0);
  bn_apply_split_kernel<<<L.grid, L.block, 0, as_stream(stream)>>>(y, static_cast<__nv_bfloat16*>(out_hi),
                                                                   static_cast<__nv_bfloat16*>(out_lo), P, C, scale,
                                                                   shift, act, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_bwd_reduce_f32(const void* da, const float* y, long long P, int C, const float* scale, const float* shift,
                         const float* mean, const float* rstd, int act, float* sum_dz, float* sum_dzx, void* stream) {
  GP_REQUIRE(da && y && sum_dz && sum_dzx && P > 0 && C % 8 == 0, "gp_bn_bwd_reduce_f32: bad arguments");
  const ColLaunch L = col_launch(P, C, 2);
  bn_bwd_reduce_f32_kernel<<<L.grid, L.block, L.smem, as_stream(stream)>>>(
      static_cast<const __nv_bfloat16*>(da), y, P, C, scale, shift, mean, rstd, act, sum_dz, sum_dzx, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_bwd_apply_f32(const void* da, const float* y, void* dy, long long P, int C, const float* scale,
                        const float* shift, const float* mean, const float* rstd, const float* sum_dz,
                        const float* sum_dzx, double count, int act, void* stream) {
  GP_REQUIRE(da && y && dy && P > 0 && C % 8 == 0 && count > 0, "gp_bn_bwd_apply_f32: bad arguments");
  const ColLaunch L = col_launch(P, C, 0);
  bn_bwd_apply_f32_kernel<<<L.grid, L.block, 0, as_stream(stream)>>>(
      static_cast<const __nv_bfloat16*>(da), y, static_cast<__nv_bfloat16*>(dy), P, C, scale, shift, mean, rstd, sum_dz,
      sum_dzx, (float)(1.0 / count), act, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_im2col_k4s2_split(const float* img, void* col_hi, void* col_lo, int NB, int ch, int Hi, int Wi, void* stream) {
  GP_REQUIRE(img && col_hi && col_lo && NB > 0 && ch > 0 && ch <= 4 && Hi % 2 == 0 && Wi % 2 == 0, "gp_im2col_k4s2_split: bad arguments");
  im2col_k4s2_split_kernel<<<grid1((long long)NB * (Hi / 2) * (Wi / 2) * 8), 256, 0, as_stream(stream)>>>(
      img, static_cast<__nv_bfloat16*>(col_hi), static_cast<__nv_bfloat16*>(col_lo), NB, ch, Hi, Wi);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_col2im_k4s2_f32(const float* col, const float* bias, float* img, int NB, int ch, int Hi, int Wi, int act,
                       void* stream) {
  GP_REQUIRE(img && col && NB > 0 && ch > 0 && ch <= 4 && Hi % 2 == 0 && Wi % 2 == 0, "gp_col2im_k4s2_f32: bad arguments");
  col2im_k4s2_f32_kernel<<<grid1((long long)NB * ch * Hi * Wi), 256, 0, as_stream(stream)>>>(col, bias, img, NB, ch, Hi, Wi, act);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_head_fwd_split(const void* a_hi, const void* a_lo, const float* w, const float* bias, float* out, int NB, int HW,
                      int C, int O, long long s_o, long long s_c, long long s_hw, void* stream) {
  GP_REQUIRE(a_hi && a_lo && w && out && NB > 0 && HW > 0 && C > 0 && O > 0, "gp_head_fwd_split: bad arguments");
  dim3 grid(NB, O);
  head_fwd_split_kernel<<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(a_hi),
                                                             static_cast<const __nv_bfloat16*>(a_lo), w, bias, out, HW,
                                                             C, O, s_o, s_c, s_hw);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

}  // extern "C"
