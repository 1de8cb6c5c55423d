// HBM-bound kernels of the hot path: weight packing, BatchNorm (statistics / apply / backward) with fused
// ReLU / LeakyReLU, im2col / col2im for the 3-channel image layers, heads, losses.
// All activations are NHWC bf16 viewed as [P pixels][C channels]; reductions accumulate in fp32.
#include <cuda_bf16.h>

#include "act_io.cuh"
#include "bn_stream.cuh"
#include "common.h"

namespace gp {

// ------------------------------------------------------------------------------------------------ packing
// dst[r][k] (bf16, row length ld_dst) = scale * src[map(r) * s_r + k * s_k] for r < R, k < K; zero elsewhere.
// perm > 1 permutes rows: dst row r reads source row (r % (R/perm)) * perm + r / (R/perm)
// (NCHW-flatten -> NHWC-flatten of the generator's first Linear, models/dcgan.py:50-51).
__global__ void pack_matrix_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int R, int K,
                                   int Rpad, int ld_dst, long long s_r, long long s_k, int perm,
                                   const float* __restrict__ inv_scale) {
  gp::pdl_sync();
  const long long total = (long long)Rpad * ld_dst;
  const float sc = inv_scale ? 1.f / __ldg(inv_scale) : 1.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld_dst), k = (int)(i % ld_dst);
    float v = 0.f;
    if (r < R && k < K) {
      int rs = r;
      if (perm > 1) {
        const int inner = R / perm;  // channels
        rs = (r % inner) * perm + r / inner;
      }
      v = __ldg(src + rs * s_r + k * s_k) * sc;
    }
    dst[i] = __float2bfloat16(v);
  }
}

// conv weight (D0, D1, taps) fp32 -> packed bf16 rows [N][tap][C], (N, C) = (D0, D1) if n_dim == 0 else (D1, D0).
// Row n starts at dst + n*ld; with lo != nullptr the rounding residuals go to the same position of `lo` (bf16x3
// weight operand, hi block | lo block per row). Both kernels transpose through shared memory so that global reads
// are contiguous runs of the source and global writes are contiguous runs of C.
constexpr int kPackCT = 64;  // channels per tile (n_dim == 0)
__global__ void pack_conv_weight_n0_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                           __nv_bfloat16* __restrict__ lo, int N, int C, int taps, long long ld,
                                           int flip, const float* __restrict__ inv_scale) {
  gp::pdl_sync();
  extern __shared__ float s_tile[];  // [kPackCT][taps + 1]
  const int n = blockIdx.y, c0 = blockIdx.x * kPackCT;
  const int ct = min(kPackCT, C - c0);
  const float sc = inv_scale ? 1.f / __ldg(inv_scale) : 1.f;
  const float* sp = src + ((long long)n * C + c0) * taps;  // contiguous ct*taps floats
  for (int i = threadIdx.x; i < ct * taps; i += blockDim.x) s_tile[(i / taps) * (taps + 1) + i % taps] = __ldg(sp + i) * sc;
  __syncthreads();
  for (int j = threadIdx.x; j < taps * ct; j += blockDim.x) {
    const int t = j / ct, c = j % ct;
    const float v = s_tile[c * (taps + 1) + (flip ? taps - 1 - t : t)];
    const __nv_bfloat16 h = __float2bfloat16(v);
    const long long o = (long long)n * ld + (long long)t * C + c0 + c;
    dst[o] = h;
    if (lo != nullptr) lo[o] = __float2bfloat16(v - __bfloat162float(h));
  }
}
constexpr int kPackC1 = 32, kPackN1 = 8;  // tile of the n_dim == 1 variant: 32 source rows (C) x 8 output rows (N)
__global__ void pack_conv_weight_n1_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                           __nv_bfloat16* __restrict__ lo, int N, int C, int taps, long long ld,
                                           int flip, const float* __restrict__ inv_scale) {
  gp::pdl_sync();
  extern __shared__ float s_tile[];  // [kPackC1][kPackN1 * taps + 1]
  const int c0 = blockIdx.x * kPackC1, n0 = blockIdx.y * kPackN1;
  const int ct = min(kPackC1, C - c0), nt = min(kPackN1, N - n0);
  const int run = nt * taps, pitch = kPackN1 * taps + 1;
  const float sc = inv_scale ? 1.f / __ldg(inv_scale) : 1.f;
  for (int i = threadIdx.x; i < ct * run; i += blockDim.x) {
    const int c = i / run, rem = i % run;
    s_tile[c * pitch + rem] = __ldg(src + ((long long)(c0 + c) * N + n0) * taps + rem) * sc;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < run * ct; j += blockDim.x) {
    const int c = j % ct, q = j / ct;  // q = n_local * taps + t
    const int nl = q / taps, t = q % taps;
    const float v = s_tile[c * pitch + nl * taps + (flip ? taps - 1 - t : t)];
    const __nv_bfloat16 h = __float2bfloat16(v);
    const long long o = (long long)(n0 + nl) * ld + (long long)t * C + c0 + c;
    dst[o] = h;
    if (lo != nullptr) lo[o] = __float2bfloat16(v - __bfloat162float(h));
  }
}

// ---- 16-tap (4x4 kernel) fast paths: 16-byte global accesses on both sides of the shared-memory transpose.
// n_dim == 0: one block = one output row n x 128 channels: reads 128*16 contiguous floats, writes 16 runs of 128 bf16.
constexpr int kPack16C = 128;
__global__ void __launch_bounds__(256) pack_conv_weight16_n0_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                                    __nv_bfloat16* __restrict__ lo, int N, int C, long long ld,
                                                                    const float* __restrict__ inv_scale) {
  gp::pdl_sync();
  __shared__ float s_tile[kPack16C * 17];  // [c][16 taps + 1 pad]
  const int n = blockIdx.y, c0 = blockIdx.x * kPack16C;  // C % 128 == 0 on this path
  const float sc = inv_scale ? 1.f / __ldg(inv_scale) : 1.f;
  const float4* sp = reinterpret_cast<const float4*>(src + ((long long)n * C + c0) * 16);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int i = threadIdx.x + k * 256;  // float4 index: c = i / 4, taps 4*(i%4)..+3
    const float4 v = __ldg(sp + i);
    float* d = s_tile + (i >> 2) * 17 + (i & 3) * 4;
    d[0] = v.x * sc, d[1] = v.y * sc, d[2] = v.z * sc, d[3] = v.w * sc;
  }
  __syncthreads();
  // thread -> (tap t, group of 8 channels)
  const int t = threadIdx.x >> 4, g = threadIdx.x & 15;
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = s_tile[(g * 8 + j) * 17 + t];
  const long long o = (long long)n * ld + (long long)t * C + c0 + g * 8;
  store8_bf16(dst + o, lo ? lo + o : nullptr, f);
}
// n_dim == 1: one block = 32 source rows (channels c) x 8 output rows n: reads 32 runs of 128 floats, writes
// 8*16 runs of 32 bf16.
__global__ void __launch_bounds__(256) pack_conv_weight16_n1_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                                    __nv_bfloat16* __restrict__ lo, int N, int C, long long ld,
                                                                    const float* __restrict__ inv_scale) {
  gp::pdl_sync();
  __shared__ float s_tile[32 * 129];  // [c][8 n x 16 taps + 1 pad]
  const int c0 = blockIdx.x * 32, n0 = blockIdx.y * 8;  // C % 32 == 0, N % 8 == 0 on this path
  const float sc = inv_scale ? 1.f / __ldg(inv_scale) : 1.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x + k * 256;  // float4 index over [32 c][32 float4]
    const int c = i >> 5, q = i & 31;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + ((long long)(c0 + c) * N + n0) * 16) + q);
    float* d = s_tile + c * 129 + q * 4;
    d[0] = v.x * sc, d[1] = v.y * sc, d[2] = v.z * sc, d[3] = v.w * sc;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int i = threadIdx.x + k * 256;  // (n_local*16 + t) * 4 + channel group of 8
    const int g = i & 3, q = i >> 2;      // q = n_local * 16 + t
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = s_tile[(g * 8 + j) * 129 + q];
    const long long o = (long long)(n0 + (q >> 4)) * ld + (long long)(q & 15) * C + c0 + g * 8;
    store8_bf16(dst + o, lo ? lo + o : nullptr, f);
  }
}
// packed gradient [M][16][N] -> (M, N, 16): one block = one m x 64 n
// acc != 0: dst += value (the parameter's .grad buffer is the destination: no AccumulateGrad pass afterwards)
__global__ void __launch_bounds__(256) unpack_conv_wgrad16_kernel(const float* __restrict__ src, float* __restrict__ dst, int M,
                                                                  int N, int acc) {
  gp::pdl_sync();
  __shared__ float s_tile[16 * 65];  // [t][64 n + 1 pad]
  const int m = blockIdx.y, n0 = blockIdx.x * 64;  // N % 64 == 0 on this path
  {
    const int t = threadIdx.x >> 4, q = threadIdx.x & 15;  // 16 taps x 16 float4
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + ((long long)m * 16 + t) * N + n0) + q);
    float* d = s_tile + t * 65 + q * 4;
    d[0] = v.x, d[1] = v.y, d[2] = v.z, d[3] = v.w;
  }
  __syncthreads();
  {
    const int n = threadIdx.x >> 2, tq = threadIdx.x & 3;  // 64 n x 4 float4 of taps
    float4 v;
    v.x = s_tile[(tq * 4 + 0) * 65 + n];
    v.y = s_tile[(tq * 4 + 1) * 65 + n];
    v.z = s_tile[(tq * 4 + 2) * 65 + n];
    v.w = s_tile[(tq * 4 + 3) * 65 + n];
    float4* d4 = reinterpret_cast<float4*>(dst + ((long long)m * N + n0 + n) * 16) + tq;
    if (acc) {
      const float4 o = *d4;
      v.x += o.x, v.y += o.y, v.z += o.z, v.w += o.w;
    }
    *d4 = v;
  }
}

// packed fp32 gradient [M][tap][N] -> torch layout (M, N, tap) fp32, tile = one m x 64 n, transposed through smem
constexpr int kUnpackNT = 64;
__global__ void unpack_conv_wgrad_kernel(const float* __restrict__ src, float* __restrict__ dst, int M, int N, int taps,
                                         int acc) {
  gp::pdl_sync();
  extern __shared__ float s_tile[];  // [taps][kUnpackNT + 1]
  const int m = blockIdx.y, n0 = blockIdx.x * kUnpackNT;
  const int nt = min(kUnpackNT, N - n0);
  for (int i = threadIdx.x; i < taps * nt; i += blockDim.x) {
    const int t = i / nt, n = i % nt;
    s_tile[t * (kUnpackNT + 1) + n] = src[((long long)m * taps + t) * N + n0 + n];
  }
  __syncthreads();
  float* dp = dst + ((long long)m * N + n0) * taps;  // contiguous nt*taps floats
  for (int j = threadIdx.x; j < nt * taps; j += blockDim.x) {
    const float v = s_tile[(j % taps) * (kUnpackNT + 1) + j / taps];
    dp[j] = acc ? dp[j] + v : v;
  }
}

// fp32 [Rpad][ld_src] (gradient of a packed matrix) -> dst[map(r) * s_r + k * s_k] for r < R, k < K (inverse of pack_matrix)
__global__ void unpack_matrix_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int K, int ld_src,
                                     long long s_r, long long s_k, int perm, int acc) {
  gp::pdl_sync();
  const long long total = (long long)R * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / K), k = (int)(i % K);
    int rs = r;
    if (perm > 1) {
      const int inner = R / perm;
      rs = (r % inner) * perm + r / inner;
    }
    const float v = src[(long long)r * ld_src + k];
    float* d = dst + rs * s_r + k * s_k;
    *d = acc ? *d + v : v;
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm
// (statistics / apply / backward kernels: bn_stream.cuh)

// mean / rstd / fused scale & shift, running statistics (torch: aten::native_batch_norm semantics:
// biased variance to normalise, unbiased into running_var, momentum 0.1, num_batches_tracked += 1).
__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq, double count, int C,
                                   float eps, float momentum, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ scale, float* __restrict__ shift, float* running_mean,
                                   float* running_var, long long* num_batches_tracked) {
  gp::pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
  if (c >= C) return;
  const double m = (double)sum[c] / count;
  double var = (double)sumsq[c] / count - m * m;
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

// eval-mode parameters from the running statistics
__global__ void bn_eval_kernel(const float* __restrict__ rm, const float* __restrict__ rv,
                               const float* __restrict__ gamma, const float* __restrict__ beta, int C, float eps,
                               float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ scale,
                               float* __restrict__ shift) {
  gp::pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float r = rsqrtf(rv[c] + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  mean[c] = rm[c];
  rstd[c] = r;
  scale[c] = g * r;
  shift[c] = b - rm[c] * g * r;
}

// dy = da * act'(a)   (activation applied directly on the conv output: a has the sign of the pre-activation)
__global__ void act_bwd_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ a,
                               __nv_bfloat16* __restrict__ dy, long long n8, int act) {
  gp::pdl_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    Vec8<__nv_bfloat16> va, vd;
    va.load(a + i * 8);
    vd.load(da + i * 8);
    float fa[8], fd[8];
    va.unpack(fa);
    vd.unpack(fd);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float gr;
      if (act == GP_ACT_TANH) gr = 1.f - fa[j] * fa[j];
      else gr = act_grad(fa[j], act);
      fd[j] *= gr;
    }
    store8_bf16(dy + i * 8, nullptr, fd);
  }
}

}  // namespace gp

namespace gp {

// ------------------------------------------------------------------------------------------------ image layers
// im2col of a k4 s2 p1 window over an NCHW fp32 image with `ch` (<= 4) channels:
//   col[(n, oh, ow)][(c*4 + kh)*4 + kw] = img[n, c, 2oh-1+kh, 2ow-1+kw] * (mul ? 1 - mul[same]^2 : 1); columns >= ch*16 are 0.
// Used for D's first conv (models/dcgan.py:106, Cin = img_dim) and for the gradient of G's last ConvT + Tanh
// (models/dcgan.py:41-44): `mul` is then the tanh output, fusing tanh'. col_lo (optional) receives the bf16 rounding
// residuals (bf16x3 operand pair).
// One block = one image x kIm2colRows output rows: the 2*rows+2 input rows of every channel are staged in shared memory
// with float4 loads, then every thread emits 16-byte column groups (8 groups = one 128-byte col row per pixel).
constexpr int kIm2colRows = 4;
__global__ void im2col_k4s2_kernel(const float* __restrict__ img, const float* __restrict__ mul,
                                   __nv_bfloat16* __restrict__ col, __nv_bfloat16* __restrict__ col_lo, int NB, int ch,
                                   int Hi, int Wi) {
  gp::pdl_sync();
  extern __shared__ float s_img[];  // [ch][2*kIm2colRows + 2][Wi]
  constexpr int kInRows = 2 * kIm2colRows + 2;
  const int Ho = Hi / 2, Wo = Wi / 2;
  const int n = blockIdx.y, oh0 = blockIdx.x * kIm2colRows;
  const int w4 = Wi / 4;
  for (int i = threadIdx.x; i < ch * kInRows * w4; i += blockDim.x) {
    const int q = i % w4, rr = (i / w4) % kInRows, c = i / (w4 * kInRows);
    const int ih = 2 * oh0 - 1 + rr;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ih >= 0 && ih < Hi) {
      const long long off = (((long long)n * ch + c) * Hi + ih) * Wi + 4 * q;
      v = __ldg(reinterpret_cast<const float4*>(img + off));
      if (mul != nullptr) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(mul + off));
        v.x *= 1.f - t.x * t.x, v.y *= 1.f - t.y * t.y, v.z *= 1.f - t.z * t.z, v.w *= 1.f - t.w * t.w;
      }
    }
    *reinterpret_cast<float4*>(s_img + (c * kInRows + rr) * Wi + 4 * q) = v;
  }
  __syncthreads();
  const int rows = min(kIm2colRows, Ho - oh0);
  for (int i = threadIdx.x; i < rows * Wo * 8; i += blockDim.x) {
    const int g = i & 7, ow = (i >> 3) % Wo, ol = (i >> 3) / Wo;
    const int c = g >> 1;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kh = (g & 1) * 2 + (j >> 2), kw = j & 3;
      const int iw = 2 * ow - 1 + kw;
      f[j] = (c < ch && iw >= 0 && iw < Wi) ? s_img[(c * kInRows + 2 * ol + kh) * Wi + iw] : 0.f;
    }
    const long long o = ((((long long)n * Ho + oh0 + ol) * Wo + ow) * 8 + g) * 8;
    store8_bf16(col + o, col_lo ? col_lo + o : nullptr, f);
  }
}

// col2im (transpose of the above): img[n, c, ih, iw] = act( bias[c] + sum_{(oh,kh): 2oh-1+kh = ih} sum_{(ow,kw)} col[(n,oh,ow)][(c*4+kh)*4+kw] )
// Used for G's last ConvT (+Tanh) forward and for the image gradient of D's first conv. TC = bf16 or fp32 col.
// One block = one image x kCol2imRows image rows: the rows/2 + 2 col rows they gather from are staged in shared
// memory with 16-byte loads (only the ch*16 live columns), then one thread per output pixel.
constexpr int kCol2imRows = 8;
template <typename TC, int CH>
__global__ void __launch_bounds__(256) col2im_k4s2_kernel(const TC* __restrict__ col, const float* __restrict__ bias,
                                                          float* __restrict__ img, int NB, int Hi, int Wi, int act) {
  gp::pdl_sync();
  extern __shared__ float s_col[];  // [kCol2imRows/2 + 2][Wo][CH*16 + 1]
  constexpr int kColRows = kCol2imRows / 2 + 2;
  constexpr int live = CH * 16, pitch = live + 1, groups = live / 8;
  const int Ho = Hi / 2, Wo = Wi / 2;
  const int n = blockIdx.y, ih0 = blockIdx.x * kCol2imRows;
  const int oh_base = ih0 / 2 - 1;
  // staging: (ow, 8-column group) pairs of one col row per iteration; divisions are by compile-time constants
  for (int r = 0; r < kColRows; ++r) {
    const int oh = oh_base + r;
    const bool row_ok = oh >= 0 && oh < Ho;
    const TC* crow = col + ((long long)n * Ho + (row_ok ? oh : 0)) * Wo * 64;
    for (int i = threadIdx.x; i < Wo * groups; i += 256) {
      const int ow = i / groups, g = i % groups;
      float f[8];
      if (row_ok) {
        Vec8<TC> v;
        v.load(crow + ow * 64 + g * 8);
        v.unpack(f);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
      float* d = s_col + (r * Wo + ow) * pitch + g * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = f[j];
    }
  }
  __syncthreads();
  // one thread per output pixel: x = threadIdx.x % 64 (+64k), (channel, row) pairs strided by threadIdx.x / 64
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  for (int cr = ty; cr < CH * kCol2imRows; cr += 4) {
    const int c = cr / kCol2imRows, il = cr % kCol2imRows;
    const int ih = ih0 + il;
    if (ih >= Hi) continue;
    const float b0 = bias ? __ldg(bias + c) : 0.f;
    // ih = 2*oh - 1 + kh  ->  kh in {(ih+1)&1, (ih+1)&1 + 2}
    const int kh0 = (ih + 1) & 1;
    for (int iw = tx; iw < Wi; iw += 64) {
      const int kw0 = (iw + 1) & 1;
      float acc = b0;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int kh = kh0 + 2 * a;
        const int oh2 = ih + 1 - kh;
        if (oh2 < 0 || oh2 >= 2 * Ho) continue;
        const int r = oh2 / 2 - oh_base;
#pragma unroll
        for (int bb = 0; bb < 2; ++bb) {
          const int kw = kw0 + 2 * bb;
          const int ow2 = iw + 1 - kw;
          if (ow2 < 0 || ow2 >= 2 * Wo) continue;
          acc += s_col[(r * Wo + ow2 / 2) * pitch + (c * 4 + kh) * 4 + kw];
        }
      }
      img[(((long long)n * CH + c) * Hi + ih) * Wi + iw] = act_fwd(acc, act);
    }
  }
}

// dbias[c] += sum_{n,h,w} dout[n,c,h,w] * (mul ? 1 - mul^2 : 1)   (bias gradient of the last ConvT under Tanh)
__global__ void image_bias_grad_kernel(const float* __restrict__ dout, const float* __restrict__ mul,
                                       float* __restrict__ dbias, int NB, int ch, int HW) {
  gp::pdl_sync();
  const int c = blockIdx.y;
  float acc = 0.f;
  const long long total = (long long)NB * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long off = ((i / HW) * ch + c) * HW + i % HW;
    float v = __ldg(dout + off);
    if (mul != nullptr) {
      const float t = __ldg(mul + off);
      v *= 1.f - t * t;
    }
    acc += v;
  }
  __shared__ float red[32];
  for (int k = 16; k > 0; k >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, k);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int k = 16; k > 0; k >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, k);
    if (threadIdx.x == 0) atomicAdd(dbias + c, acc);
  }
}

// ------------------------------------------------------------------------------------------------ heads
// out[b][o] = bias[o] + sum_{hw, c} (a + a_lo)[b, hw, c] * w[o*s_o + c*s_c + hw*s_hw]
// (sum-pool + Linear: s_hw = 0, models/dcgan.py:121-122; flatten + Linear: s_c = HW, s_hw = 1, dcgan_specnorm.py:125-126)
// One block per (b, o); threads read 8 channels (16 bytes) at a time. a_lo: optional low halves (bf16x3 mode).
__global__ void head_fwd_kernel(const __nv_bfloat16* __restrict__ a, const void* __restrict__ a_comp, int fmt,
                                const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
                                int HW, int C, int O, long long s_o, long long s_c, long long s_hw) {
  gp::pdl_sync();
  const int b = blockIdx.x, o = blockIdx.y;
  float acc = 0.f;
  const long long base = (long long)b * HW * C;
  const float* wo = w + o * s_o;
  const int c8 = C / 8;
  const int total = HW * c8;
  if (s_c == 1 && ((s_o | s_hw) & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0) {
    // channels contiguous in the weight (pooled Linear, or the flatten head with its weight staged as [O][HW][C]):
    // 16-byte weight loads, four feature groups in flight per thread
    for (int i0 = threadIdx.x; i0 < total; i0 += 4 * blockDim.x) {
      float f[4][8];
      float4 w0[4], w1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < total) {
          load8c(a, a_comp, fmt, base + (long long)i * 8, f[u]);   // most precise view of the features (act_io.cuh)
          const float4* wp = reinterpret_cast<const float4*>(wo + (i % c8) * 8 + (long long)(i / c8) * s_hw);
          w0[u] = __ldg(wp);
          w1[u] = __ldg(wp + 1);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[u][j] = 0.f;
          w0[u] = w1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        acc += f[u][0] * w0[u].x + f[u][1] * w0[u].y + f[u][2] * w0[u].z + f[u][3] * w0[u].w + f[u][4] * w1[u].x +
               f[u][5] * w1[u].y + f[u][6] * w1[u].z + f[u][7] * w1[u].w;
    }
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int c = (i % c8) * 8, hw = i / c8;
      float f[8];
      load8c(a, a_comp, fmt, base + (long long)i * 8, f);
      const float* wp = wo + c * s_c + hw * s_hw;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += f[j] * __ldg(wp + j * s_c);
    }
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

// Several outputs per sample (ACGAN's packed adversarial + auxiliary heads, models/acgan.py:122-126): ONE pass over the
// sample's features accumulates all O <= kHeadMaxO outputs (the per-(b, o) kernel above would re-read them O times).
constexpr int kHeadMaxO = 16;
__global__ void __launch_bounds__(256) head_fwd_multi_kernel(const __nv_bfloat16* __restrict__ a, const void* __restrict__ a_comp,
                                                             int fmt, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ out, int HW,
                                                             int C, int O, long long s_o, long long s_c, long long s_hw) {
  gp::pdl_sync();
  const int b = blockIdx.x;
  float acc[kHeadMaxO];
#pragma unroll
  for (int o = 0; o < kHeadMaxO; ++o) acc[o] = 0.f;
  const long long base = (long long)b * HW * C;
  const int c8 = C / 8;
  for (int i = threadIdx.x; i < HW * c8; i += blockDim.x) {
    const int c = (i % c8) * 8, hw = i / c8;
    float f[8];
    load8c(a, a_comp, fmt, base + (long long)i * 8, f);
    const float* wp = w + c * s_c + hw * s_hw;
#pragma unroll
    for (int o = 0; o < kHeadMaxO; ++o) {
      if (o < O) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += f[j] * __ldg(wp + o * s_o + j * s_c);
        acc[o] += t;
      }
    }
  }
  __shared__ float red[kHeadMaxO][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 0; o < kHeadMaxO; ++o) {
    float v = acc[o];
    for (int k = 16; k > 0; k >>= 1) v += __shfl_xor_sync(0xffffffffu, v, k);
    if (lane == 0) red[o][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < O) {
    float v = 0.f;
    for (int k = 0; k < 8; ++k) v += red[threadIdx.x][k];
    out[(long long)b * O + threadIdx.x] = v + (bias ? bias[threadIdx.x] : 0.f);
  }
}

// da[b, hw, c] = sum_o dout[b][o] * w[o, c, hw]      (one thread per 8 channels)
__global__ void head_bwd_data_kernel(const float* __restrict__ dout, const float* __restrict__ w,
                                     __nv_bfloat16* __restrict__ da, int NB, int HW, int C, int O, long long s_o,
                                     long long s_c, long long s_hw) {
  gp::pdl_sync();
  const int c8 = C / 8;
  const long long total = (long long)NB * HW * c8;
  const bool vec = s_c == 1 && ((s_o | s_hw) & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8) * 8, hw = (int)((i / c8) % HW);
    const int b = (int)(i / ((long long)c8 * HW));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int o = 0; o < O; ++o) {
      const float d = __ldg(dout + (long long)b * O + o);
      const float* wp = w + o * s_o + c * s_c + hw * s_hw;
      if (vec) {   // channels contiguous in the weight: two 16-byte loads
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp)), w1 = __ldg(reinterpret_cast<const float4*>(wp) + 1);
        acc[0] += d * w0.x, acc[1] += d * w0.y, acc[2] += d * w0.z, acc[3] += d * w0.w;
        acc[4] += d * w1.x, acc[5] += d * w1.y, acc[6] += d * w1.z, acc[7] += d * w1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += d * __ldg(wp + j * s_c);
      }
    }
    store8_bf16(da + i * 8, nullptr, acc);
  }
}

// dw[o, c, hw] (+)= sum_b dout[b][o] * a[b, hw, c]  (sum-pool: s_hw = 0 so the hw terms accumulate); dbias[o] = sum_b dout[b][o]
// One thread per (8-channel group, hw) x batch chunk (grid.z); dw must be zeroed by the caller.
__global__ void head_bwd_weight_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ a,
                                       float* __restrict__ dw, float* __restrict__ dbias, int NB, int HW, int C, int O,
                                       long long s_o, long long s_c, long long s_hw) {
  gp::pdl_sync();
  const int o = blockIdx.y;
  const int c8 = C / 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // index over HW * C/8
  const int bchunk = (NB + gridDim.z - 1) / gridDim.z;
  const int b0 = blockIdx.z * bchunk, b1 = min(b0 + bchunk, NB);
  if (i < HW * c8) {
    const int c = (i % c8) * 8, hw = i / c8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const __nv_bfloat16* ap = a + (long long)i * 8;
    const long long sb = (long long)HW * C;
    int b = b0;
    for (; b + 4 <= b1; b += 4) {
      Vec8<__nv_bfloat16> v[4];
      float d[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u].load(ap + (b + u) * sb);
        d[u] = __ldg(dout + (long long)(b + u) * O + o);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        v[u].unpack(f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += d[u] * f[j];
      }
    }
    for (; b < b1; ++b) {
      Vec8<__nv_bfloat16> v;
      v.load(ap + b * sb);
      float f[8];
      v.unpack(f);
      const float d = __ldg(dout + (long long)b * O + o);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += d * f[j];
    }
    float* wp = dw + o * s_o + c * s_c + hw * s_hw;
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(wp + j * s_c, acc[j]);
  }
  if (dbias != nullptr && blockIdx.x == 0 && blockIdx.z == 0) {  // dbias[o] = sum_b dout[b][o]: block-wide reduction
    float s = 0.f;
    for (int b = threadIdx.x; b < NB; b += blockDim.x) s += __ldg(dout + (long long)b * O + o);
    __shared__ float red[32];
    for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
      for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
      if (threadIdx.x == 0) dbias[o] = s;
    }
  }
}

// Sum-pool head (s_hw == 0, models/dcgan.py:121-122): every spatial position adds into the same dw[o][c], so the
// reduction over (batch chunk, hw) is done in registers and across the 8 warps of the block in shared memory before one
// atomic per channel per block. Block = 8 warps x (32 lanes x 8 channels); warp w takes samples b0 + w, b0 + w + 8, ...
__global__ void __launch_bounds__(256) head_bwd_weight_pool_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ a,
                                                                   float* __restrict__ dw, int NB, int HW, int C, int O,
                                                                   long long s_o, long long s_c) {
  gp::pdl_sync();
  __shared__ float s_acc[8][256];
  const int o = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = (blockIdx.x * 32 + lane) * 8;
  const int bchunk = (NB + gridDim.z - 1) / gridDim.z;
  const int b0 = blockIdx.z * bchunk, b1 = min(b0 + bchunk, NB);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c < C) {
    for (int b = b0 + warp; b < b1; b += 8) {
      const float d = __ldg(dout + (long long)b * O + o);
      const __nv_bfloat16* ap = a + (long long)b * HW * C + c;
      float t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = 0.f;
      int hw = 0;
      for (; hw + 4 <= HW; hw += 4) {
        Vec8<__nv_bfloat16> v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u].load(ap + (long long)(hw + u) * C);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float f[8];
          v[u].unpack(f);
#pragma unroll
          for (int j = 0; j < 8; ++j) t[j] += f[j];
        }
      }
      for (; hw < HW; ++hw) {
        Vec8<__nv_bfloat16> v;
        v.load(ap + (long long)hw * C);
        float f[8];
        v.unpack(f);
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] += f[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += d * t[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s_acc[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  {
    const int i = threadIdx.x;  // channel slot within the block
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += s_acc[w][i];
    const int cc = blockIdx.x * 256 + i;
    if (cc < C) atomicAdd(dw + o * s_o + cc * s_c, v);
  }
}

// The same for O <= kHeadMaxO outputs at once: the pooled features sum_hw a[b, hw, c] do not depend on o, so one pass
// over `a` serves every output (grid: channel blocks x batch chunks).
__global__ void __launch_bounds__(256) head_bwd_weight_pool_multi_kernel(const float* __restrict__ dout,
                                                                         const __nv_bfloat16* __restrict__ a,
                                                                         float* __restrict__ dw, int NB, int HW, int C, int O,
                                                                         long long s_o, long long s_c) {
  gp::pdl_sync();
  __shared__ float s_acc[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = (blockIdx.x * 32 + lane) * 8;
  const int bchunk = (NB + gridDim.y - 1) / gridDim.y;
  const int b0 = blockIdx.y * bchunk, b1 = min(b0 + bchunk, NB);
  float acc[kHeadMaxO][8];
#pragma unroll
  for (int o = 0; o < kHeadMaxO; ++o)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[o][j] = 0.f;
  if (c < C) {
    for (int b = b0 + warp; b < b1; b += 8) {
      const __nv_bfloat16* ap = a + (long long)b * HW * C + c;
      float t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = 0.f;
      for (int hw = 0; hw < HW; ++hw) {
        Vec8<__nv_bfloat16> v;
        v.load(ap + (long long)hw * C);
        float f[8];
        v.unpack(f);
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] += f[j];
      }
#pragma unroll
      for (int o = 0; o < kHeadMaxO; ++o) {
        if (o < O) {
          const float d = __ldg(dout + (long long)b * O + o);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[o][j] += d * t[j];
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < kHeadMaxO; ++o) {
    if (o < O) {   // uniform across the block
#pragma unroll
      for (int j = 0; j < 8; ++j) s_acc[warp][lane * 8 + j] = acc[o][j];
      __syncthreads();
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += s_acc[w][threadIdx.x];
      const int cc = blockIdx.x * 256 + threadIdx.x;
      if (cc < C) atomicAdd(dw + o * s_o + cc * s_c, v);
      __syncthreads();
    }
  }
}

// dbias[o] = sum_b dout[b][o]   (one block per output)
__global__ void head_bias_grad_kernel(const float* __restrict__ dout, float* __restrict__ dbias, int NB, int O) {
  gp::pdl_sync();
  const int o = blockIdx.x;
  float s = 0.f;
  for (int b = threadIdx.x; b < NB; b += blockDim.x) s += __ldg(dout + (long long)b * O + o);
  __shared__ float red[32];
  for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    if (threadIdx.x == 0) dbias[o] = s;
  }
}

// ------------------------------------------------------------------------------------------------ losses
// GANLoss (utils/criterion.py:30-41), value and d(loss)/d(pred) in one pass; single block.
// mode 0: BCE-with-logits vs constant target; 1: MSE vs constant target; 2: hinge real relu(1-p); 3: hinge fake relu(1+p); 4: -p
__device__ __forceinline__ void gan_loss_term(int mode, float x, float target, float& l, float& d) {
  if (mode == 0) {
    l = fmaxf(x, 0.f) - x * target + log1pf(expf(-fabsf(x)));
    d = 1.f / (1.f + expf(-x)) - target;
  } else if (mode == 1) {
    l = (x - target) * (x - target);
    d = 2.f * (x - target);
  } else if (mode == 2) {
    l = fmaxf(1.f - x, 0.f);
    d = (1.f - x) > 0.f ? -1.f : 0.f;
  } else if (mode == 3) {
    l = fmaxf(1.f + x, 0.f);
    d = (1.f + x) > 0.f ? 1.f : 0.f;
  } else {
    l = -x;
    d = -1.f;
  }
}

// sum over the (single) block; the result is valid in thread 0. `red` is 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float acc, float* red) {
  for (int k = 16; k > 0; k >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, k);
  __syncthreads();                       // `red` may still be read by a previous call
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32)
    for (int k = 16; k > 0; k >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, k);
  return acc;
}

__global__ void gan_loss_kernel(const float* __restrict__ pred, int n, int mode, float target, float inv_n,
                                float* __restrict__ loss, float* __restrict__ dpred) {
  gp::pdl_sync();
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float l, d;
    gan_loss_term(mode, pred[i], target, l, d);
    acc += l;
    dpred[i] = d * inv_n;
  }
  __shared__ float red[32];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) *loss = acc * inv_n;
}

// ACGAN objective (main_acgan.py:95-97,114-116,129-131) on the packed two-head logits [NB][1 + K] (column 0 = the
// adversarial logit, columns 1..K = the auxiliary head): GANLoss term on column 0 + aux_weight * MSELoss(mean over NB*K)
// against the float label vectors, plus mean(sigmoid(adv)) — the D(x) / D(G(z)) number the script prints (:94,112,127).
// out[0] = adversarial term, out[1] = auxiliary term (unweighted), out[2] = adv + aux_weight * aux, out[3] = sigmoid mean;
// dlogits = d out[2] / d logits in the packed layout. Single block.
__global__ void acgan_loss_kernel(const float* __restrict__ logits, const float* __restrict__ labels, int NB, int K,
                                  int mode, float target, float aux_weight, float* __restrict__ out,
                                  float* __restrict__ dlogits) {
  gp::pdl_sync();
  const int ld = K + 1;
  const float inv_nb = 1.f / (float)NB, inv_aux = 1.f / ((float)NB * (float)K);
  float s_adv = 0.f, s_aux = 0.f, s_sig = 0.f;
  for (int i = threadIdx.x; i < NB * ld; i += blockDim.x) {
    const int b = i / ld, j = i - b * ld;
    const float x = logits[i];
    if (j == 0) {
      float l, d;
      gan_loss_term(mode, x, target, l, d);
      s_adv += l;
      s_sig += 1.f / (1.f + expf(-x));
      dlogits[i] = d * inv_nb;
    } else {
      const float e = x - labels[b * K + (j - 1)];
      s_aux += e * e;
      dlogits[i] = 2.f * e * inv_aux * aux_weight;
    }
  }
  __shared__ float red[32];
  s_adv = block_sum(s_adv, red);
  s_aux = block_sum(s_aux, red);
  s_sig = block_sum(s_sig, red);
  if (threadIdx.x == 0) {
    out[0] = s_adv * inv_nb;
    out[1] = s_aux * inv_aux;
    out[2] = s_adv * inv_nb + aux_weight * (s_aux * inv_aux);
    out[3] = s_sig * inv_nb;
  }
}

static inline int grid_for(long long n, int block = 256, int max_blocks = 148 * 16) {
  long long g = (n + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return (int)g;
}

// fp32 (D0, D1, taps) -> bf16 rows [N][tap][C] at pitch ld, optional residual block `lo` (see the kernels above)
static int launch_pack_conv_weight(const float* src, __nv_bfloat16* dst, __nv_bfloat16* lo, int D0, int D1, int taps,
                                   int n_dim_flags, long long ld, const float* inv_scale, cudaStream_t st) {
  const int n_dim = n_dim_flags & 1, flip = (n_dim_flags >> 1) & 1;
  const int N = n_dim == 0 ? D0 : D1, C = n_dim == 0 ? D1 : D0;
  if (taps == 16 && !flip && n_dim == 0 && C % kPack16C == 0) {
    gp::launch_pdl(pack_conv_weight16_n0_kernel, dim3(C / kPack16C, N), 256, 0, st, src, dst, lo, N, C, ld, inv_scale);
  } else if (taps == 16 && !flip && n_dim == 1 && C % 32 == 0 && N % 8 == 0) {
    gp::launch_pdl(pack_conv_weight16_n1_kernel, dim3(C / 32, N / 8), 256, 0, st, src, dst, lo, N, C, ld, inv_scale);
  } else if (n_dim == 0) {
    dim3 grid((D1 + kPackCT - 1) / kPackCT, D0);
    gp::launch_pdl(pack_conv_weight_n0_kernel, grid, 256, (size_t)kPackCT * (taps + 1) * sizeof(float), st, src, dst, lo, D0, D1, taps,
                                                                                             ld, flip, inv_scale);
  } else {
    dim3 grid((D0 + kPackC1 - 1) / kPackC1, (D1 + kPackN1 - 1) / kPackN1);
    gp::launch_pdl(pack_conv_weight_n1_kernel, grid, 256, (size_t)kPackC1 * (kPackN1 * taps + 1) * sizeof(float), st, src, dst, lo, D1, D0, taps, ld, flip, inv_scale);
  }
  GP_CHECK_LAUNCH();
  return GP_OK;
}

}  // namespace gp

using namespace gp;

extern "C" {

int gp_pack_matrix(const float* src, void* dst, int R, int K, int Rpad, int ld_dst, long long s_r, long long s_k,
                   int perm, const float* inv_scale, void* stream) {
  GP_REQUIRE(src && dst && R > 0 && K > 0 && Rpad >= R && ld_dst >= K, "gp_pack_matrix: bad arguments");
  GP_REQUIRE(perm <= 1 || R % perm == 0, "gp_pack_matrix: perm must divide R");
  gp::launch_pdl(pack_matrix_kernel, grid_for((long long)Rpad * ld_dst), 256, 0, as_stream(stream), src, static_cast<__nv_bfloat16*>(dst), R, K, Rpad, ld_dst, s_r, s_k, perm, inv_scale);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_unpack_matrix(const float* src, float* dst, int R, int K, int ld_src, long long s_r, long long s_k, int perm,
                     void* stream) {
  GP_REQUIRE(src && dst && R > 0 && K > 0 && ld_src >= K, "gp_unpack_matrix: bad arguments");
  gp::launch_pdl(unpack_matrix_kernel, grid_for((long long)R * K), 256, 0, as_stream(stream), src, dst, R, K, ld_src, s_r, s_k,
                                                                                  perm & 0x3fffffff, (perm >> 30) & 1);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_pack_conv_weight(const float* src, void* dst, int D0, int D1, int taps, int n_dim, const float* inv_scale,
                        void* stream) {
  GP_REQUIRE(src && dst && D0 > 0 && D1 > 0 && taps > 0 && n_dim >= 0 && n_dim <= 3, "gp_pack_conv_weight: bad arguments");
  GP_REQUIRE(taps <= 64, "gp_pack_conv_weight: at most 64 taps");
  const int C = (n_dim & 1) == 0 ? D1 : D0;
  return launch_pack_conv_weight(src, static_cast<__nv_bfloat16*>(dst), nullptr, D0, D1, taps, n_dim,
                                 (long long)taps * C, inv_scale, as_stream(stream));
}

// bf16x3 weight operand: rows [N][2][taps][C] = hi block | lo block (rounding residuals)
int gp_split_conv_weight(const float* src, void* dst, int D0, int D1, int taps, int n_dim, void* stream) {
  GP_REQUIRE(src && dst && D0 > 0 && D1 > 0 && taps > 0 && (n_dim == 0 || n_dim == 1), "gp_split_conv_weight: bad arguments");
  GP_REQUIRE(taps <= 64, "gp_split_conv_weight: at most 64 taps");
  const int C = n_dim == 0 ? D1 : D0;
  __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(dst);
  return launch_pack_conv_weight(src, hi, hi + (long long)taps * C, D0, D1, taps, n_dim, 2LL * taps * C, nullptr,
                                 as_stream(stream));
}

int gp_unpack_conv_wgrad(const float* src, float* dst, int M, int N, int taps, void* stream) {
  // bit 30 of `taps` (GP_UNPACK_ACCUMULATE): dst += value instead of dst = value
  const int acc = (taps >> 30) & 1;
  taps &= 0x3fffffff;
  GP_REQUIRE(src && dst && M > 0 && N > 0 && taps > 0 && taps <= 64, "gp_unpack_conv_wgrad: bad arguments");
  if (taps == 16 && N % 64 == 0) {
    gp::launch_pdl(unpack_conv_wgrad16_kernel, dim3(N / 64, M), 256, 0, as_stream(stream), src, dst, M, N, acc);
    GP_CHECK_LAUNCH();
    return GP_OK;
  }
  dim3 grid((N + kUnpackNT - 1) / kUnpackNT, M);
  gp::launch_pdl(unpack_conv_wgrad_kernel, grid, 256, (size_t)taps * (kUnpackNT + 1) * sizeof(float), as_stream(stream), src, dst, M,
                                                                                                          N, taps, acc);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_stats(const void* x, long long P, int C, float* sum, float* sumsq, void* stream) {
  GP_REQUIRE(x && sum && sumsq && P > 0 && C > 0 && C % 8 == 0, "gp_bn_stats: bad arguments (C %% 8 == 0 required)");
  const ColLaunch L = col_launch(P, C, 2);
  gp::launch_pdl(col_stats_kernel<__nv_bfloat16, true>, L.grid, L.block, L.smem, as_stream(stream), static_cast<const __nv_bfloat16*>(x), P, C, sum, sumsq, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_finalize(const float* sum, const float* sumsq, double count, int C, float eps, float momentum,
                   const float* gamma, const float* beta, float* mean, float* rstd, float* scale, float* shift,
                   float* running_mean, float* running_var, long long* num_batches_tracked, void* stream) {
  GP_REQUIRE(sum && sumsq && mean && rstd && scale && shift && C > 0 && count > 0, "gp_bn_finalize: bad arguments");
  gp::launch_pdl(bn_finalize_kernel, (C + 127) / 128, 128, 0, as_stream(stream), sum, sumsq, count, C, eps, momentum, gamma, beta,
                                                                    mean, rstd, scale, shift, running_mean,
                                                                    running_var, num_batches_tracked);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_eval_params(const float* running_mean, const float* running_var, const float* gamma, const float* beta, int C,
                      float eps, float* mean, float* rstd, float* scale, float* shift, void* stream) {
  GP_REQUIRE(running_mean && running_var && mean && rstd && scale && shift && C > 0, "gp_bn_eval_params: bad arguments");
  gp::launch_pdl(bn_eval_kernel, (C + 127) / 128, 128, 0, as_stream(stream), running_mean, running_var, gamma, beta, C, eps, mean,
                                                                rstd, scale, shift);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_apply_act(const void* y, void* out, long long P, int C, const float* scale, const float* shift, int act,
                    void* stream) {
  GP_REQUIRE(y && out && scale && shift && P > 0 && C % 8 == 0, "gp_bn_apply_act: bad arguments");
  const ColLaunch L = col_launch(P, C, 0);
  gp::launch_pdl(bn_apply_kernel<__nv_bfloat16>, L.grid, L.block, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(out), nullptr, P, C, scale, shift, act, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_bwd_reduce(const void* da, const void* y, long long P, int C, const float* scale, const float* shift,
                     const float* mean, const float* rstd, int act, float* sum_dz, float* sum_dzx, void* stream) {
  GP_REQUIRE(da && y && sum_dz && sum_dzx && P > 0 && C % 8 == 0, "gp_bn_bwd_reduce: bad arguments");
  const ColLaunch L = col_launch(P, C, 2, 2);
  gp::launch_pdl(bn_bwd_reduce_kernel<__nv_bfloat16>, L.grid, L.block, L.smem, as_stream(stream), static_cast<const __nv_bfloat16*>(da), static_cast<const __nv_bfloat16*>(y), P, C, scale, shift, mean, rstd, act,
      sum_dz, sum_dzx, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_bwd_apply(const void* da, const void* y, void* dy, long long P, int C, const float* scale,
                    const float* shift, const float* mean, const float* rstd, const float* sum_dz,
                    const float* sum_dzx, double count, int act, float* acc_dbeta, float* acc_dgamma, float acc_scale,
                    void* stream) {
  GP_REQUIRE(da && y && dy && P > 0 && C % 8 == 0 && count > 0, "gp_bn_bwd_apply: bad arguments");
  GP_REQUIRE((acc_dbeta == nullptr) == (acc_dgamma == nullptr), "gp_bn_bwd_apply: acc_dbeta and acc_dgamma go together");
  const ColLaunch L = col_launch(P, C, 0);
  gp::launch_pdl(bn_bwd_apply_kernel<__nv_bfloat16>, L.grid, L.block, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(da), static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(dy), P, C,
      scale, shift, mean, rstd, sum_dz, sum_dzx, (float)(1.0 / count), act, L.rpb, acc_dbeta, acc_dgamma, acc_scale);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_act_bwd(const void* da, const void* a, void* dy, long long n, int act, void* stream) {
  GP_REQUIRE(da && a && dy && n > 0 && n % 8 == 0, "gp_act_bwd: bad arguments");
  gp::launch_pdl(act_bwd_kernel, grid_for(n / 8), 256, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(da),
                                                                 static_cast<const __nv_bfloat16*>(a),
                                                                 static_cast<__nv_bfloat16*>(dy), n / 8, act);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_colsum(const void* x, long long P, int C, float* out, void* stream) {
  GP_REQUIRE(x && out && P > 0 && C % 8 == 0, "gp_colsum: bad arguments");
  const ColLaunch L = col_launch(P, C, 1);
  gp::launch_pdl(col_stats_kernel<__nv_bfloat16, false>, L.grid, L.block, L.smem, as_stream(stream), static_cast<const __nv_bfloat16*>(x), P, C, out, nullptr, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

}  // extern "C"

static int launch_im2col(const float* img, const float* mul, void* col, void* col_lo, int NB, int ch, int Hi, int Wi,
                         void* stream) {
  GP_REQUIRE(img && col && NB > 0 && ch > 0 && ch <= 4 && Hi % 2 == 0 && Wi % 8 == 0, "gp_im2col_k4s2: bad arguments (Wi %% 8 == 0)");
  const size_t smem = (size_t)ch * (2 * kIm2colRows + 2) * Wi * sizeof(float);
  GP_REQUIRE(smem <= 48 * 1024, "gp_im2col_k4s2: image rows too wide (Wi=%d)", Wi);
  dim3 grid((Hi / 2 + kIm2colRows - 1) / kIm2colRows, NB);
  gp::launch_pdl(im2col_k4s2_kernel, grid, 256, smem, as_stream(stream), img, mul, static_cast<__nv_bfloat16*>(col),
                                                             static_cast<__nv_bfloat16*>(col_lo), NB, ch, Hi, Wi);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

template <typename TC, int CH>
static int launch_col2im_ch(const TC* col, const float* bias, float* img, int NB, int Hi, int Wi, int act, cudaStream_t st) {
  const size_t smem = (size_t)(kCol2imRows / 2 + 2) * (Wi / 2) * (CH * 16 + 1) * sizeof(float);
  auto kfn = col2im_k4s2_kernel<TC, CH>;
  if (smem > 48 * 1024) {
    GP_REQUIRE(smem <= 200 * 1024, "gp_col2im_k4s2: image rows too wide (Wi=%d)", Wi);
    GP_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid((Hi + kCol2imRows - 1) / kCol2imRows, NB);
  gp::launch_pdl(kfn, grid, 256, smem, st, col, bias, img, NB, Hi, Wi, act);
  GP_CHECK_LAUNCH();
  return GP_OK;
}
template <typename TC>
static int launch_col2im(const TC* col, const float* bias, float* img, int NB, int ch, int Hi, int Wi, int act,
                         void* stream) {
  GP_REQUIRE(img && col && NB > 0 && ch > 0 && ch <= 4 && Hi % 2 == 0 && Wi % 2 == 0, "gp_col2im_k4s2: bad arguments");
  cudaStream_t st = as_stream(stream);
  switch (ch) {
    case 1: return launch_col2im_ch<TC, 1>(col, bias, img, NB, Hi, Wi, act, st);
    case 2: return launch_col2im_ch<TC, 2>(col, bias, img, NB, Hi, Wi, act, st);
    case 3: return launch_col2im_ch<TC, 3>(col, bias, img, NB, Hi, Wi, act, st);
    default: return launch_col2im_ch<TC, 4>(col, bias, img, NB, Hi, Wi, act, st);
  }
}

extern "C" {

int gp_im2col_k4s2(const float* img, const float* mul, void* col, int NB, int ch, int Hi, int Wi, void* stream) {
  return launch_im2col(img, mul, col, nullptr, NB, ch, Hi, Wi, stream);
}

int gp_im2col_k4s2_split(const float* img, void* col_hi, void* col_lo, int NB, int ch, int Hi, int Wi, void* stream) {
  GP_REQUIRE(col_lo != nullptr, "gp_im2col_k4s2_split: null pointer");
  return launch_im2col(img, nullptr, col_hi, col_lo, NB, ch, Hi, Wi, stream);
}

int gp_col2im_k4s2(const void* col, const float* bias, float* img, int NB, int ch, int Hi, int Wi, int act,
                   void* stream) {
  return launch_col2im(static_cast<const __nv_bfloat16*>(col), bias, img, NB, ch, Hi, Wi, act, stream);
}

int gp_col2im_k4s2_f32(const float* col, const float* bias, float* img, int NB, int ch, int Hi, int Wi, int act,
                       void* stream) {
  return launch_col2im(col, bias, img, NB, ch, Hi, Wi, act, stream);
}

int gp_image_bias_grad(const float* dout, const float* mul, float* dbias, int NB, int ch, int HW, void* stream) {
  GP_REQUIRE(dout && dbias && NB > 0 && ch > 0 && HW > 0, "gp_image_bias_grad: bad arguments");
  dim3 grid(grid_for((long long)NB * HW, 256, 148 * 2), ch);
  gp::launch_pdl(image_bias_grad_kernel, grid, 256, 0, as_stream(stream), dout, mul, dbias, NB, ch, HW);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_head_fwd(const void* a, const float* w, const float* bias, float* out, int NB, int HW, int C, int O,
                long long s_o, long long s_c, long long s_hw, void* stream) {
  GP_REQUIRE(a && w && out && NB > 0 && HW > 0 && C > 0 && C % 8 == 0 && O > 0, "gp_head_fwd: bad arguments");
  if (O > 1 && O <= kHeadMaxO) {
    gp::launch_pdl(head_fwd_multi_kernel, NB, 256, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(a), nullptr, GP_COMP_NONE, w, bias,
                                                             out, HW, C, O, s_o, s_c, s_hw);
    GP_CHECK_LAUNCH();
    return GP_OK;
  }
  dim3 grid(NB, O);
  gp::launch_pdl(head_fwd_kernel, grid, 256, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(a), nullptr, GP_COMP_NONE, w, bias,
                                                       out, HW, C, O, s_o, s_c, s_hw);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_head_fwd_split(const void* a_hi, const void* a_lo, const float* w, const float* bias, float* out, int NB, int HW,
                      int C, int O, long long s_o, long long s_c, long long s_hw, void* stream) {
  GP_REQUIRE(a_hi && a_lo && w && out && NB > 0 && HW > 0 && C > 0 && C % 8 == 0 && O > 0, "gp_head_fwd_split: bad arguments");
  dim3 grid(NB, O);
  gp::launch_pdl(head_fwd_kernel, grid, 256, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(a_hi), a_lo, GP_COMP_LO, w, bias,
                                                       out, HW, C, O, s_o, s_c, s_hw);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_head_fwd_comp(const void* a, const void* a_comp, int comp_fmt, const float* w, const float* bias, float* out, int NB,
                     int HW, int C, int O, long long s_o, long long s_c, long long s_hw, void* stream) {
  GP_REQUIRE(a && w && out && NB > 0 && HW > 0 && C > 0 && C % 8 == 0 && O > 0, "gp_head_fwd_comp: bad arguments");
  GP_REQUIRE(comp_fmt >= GP_COMP_NONE && comp_fmt <= GP_COMP_F16, "gp_head_fwd_comp: unknown companion format %d", comp_fmt);
  if (O > 1 && O <= kHeadMaxO) {   // several heads packed: read the features once for all of them
    gp::launch_pdl(head_fwd_multi_kernel, NB, 256, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(a), a_comp, comp_fmt, w, bias, out,
                                                             HW, C, O, s_o, s_c, s_hw);
    GP_CHECK_LAUNCH();
    return GP_OK;
  }
  dim3 grid(NB, O);
  gp::launch_pdl(head_fwd_kernel, grid, 256, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(a), a_comp, comp_fmt, w, bias, out,
                                                       HW, C, O, s_o, s_c, s_hw);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_head_bwd(const float* dout, const void* a, const float* w, void* da, float* dw, float* dbias, int NB, int HW,
                int C, int O, long long s_o, long long s_c, long long s_hw, void* stream) {
  GP_REQUIRE(dout && a && w && NB > 0 && HW > 0 && C > 0 && C % 8 == 0 && O > 0, "gp_head_bwd: bad arguments");
  if (da != nullptr) {
    gp::launch_pdl(head_bwd_data_kernel, grid_for((long long)NB * HW * (C / 8)), 256, 0, as_stream(stream), dout, w, static_cast<__nv_bfloat16*>(da), NB, HW, C, O, s_o, s_c, s_hw);
    GP_CHECK_LAUNCH();
  }
  if (dw != nullptr && s_hw == 0 && O > 1 && O <= kHeadMaxO) {
    const int gx = (C + 255) / 256;
    int zsplit = (2 * num_sms()) / gx;
    if (zsplit < 1) zsplit = 1;
    if (zsplit > (NB + 7) / 8) zsplit = (NB + 7) / 8;
    gp::launch_pdl(head_bwd_weight_pool_multi_kernel, dim3(gx, zsplit), 256, 0, as_stream(stream), dout, static_cast<const __nv_bfloat16*>(a), dw, NB, HW, C, O, s_o, s_c);
    GP_CHECK_LAUNCH();
    if (dbias != nullptr) {
      gp::launch_pdl(head_bias_grad_kernel, O, 256, 0, as_stream(stream), dout, dbias, NB, O);
      GP_CHECK_LAUNCH();
    }
  } else if (dw != nullptr && s_hw == 0) {
    const int gx = (C + 255) / 256;
    int zsplit = (4 * num_sms()) / (gx * O);
    if (zsplit < 1) zsplit = 1;
    if (zsplit > NB) zsplit = NB;
    gp::launch_pdl(head_bwd_weight_pool_kernel, dim3(gx, O, zsplit), 256, 0, as_stream(stream), dout, static_cast<const __nv_bfloat16*>(a),
                                                                                 dw, NB, HW, C, O, s_o, s_c);
    GP_CHECK_LAUNCH();
    if (dbias != nullptr) {
      gp::launch_pdl(head_bias_grad_kernel, O, 256, 0, as_stream(stream), dout, dbias, NB, O);
      GP_CHECK_LAUNCH();
    }
  } else if (dw != nullptr) {
    const int gx = (HW * (C / 8) + 127) / 128;
    int zsplit = (4 * num_sms()) / (gx * O);
    if (zsplit < 1) zsplit = 1;
    if (zsplit > 64) zsplit = 64;
    if (zsplit > NB) zsplit = NB;
    dim3 grid(gx, O, zsplit);
    gp::launch_pdl(head_bwd_weight_kernel, grid, 128, 0, as_stream(stream), dout, static_cast<const __nv_bfloat16*>(a), dw, dbias, NB,
                                                                HW, C, O, s_o, s_c, s_hw);
    GP_CHECK_LAUNCH();
  }
  return GP_OK;
}

int gp_gan_loss(const float* pred, int n, int mode, float target, float* loss, float* dpred, void* stream) {
  GP_REQUIRE(pred && loss && dpred && n > 0 && mode >= 0 && mode <= 4, "gp_gan_loss: bad arguments");
  gp::launch_pdl(gan_loss_kernel, 1, 256, 0, as_stream(stream), pred, n, mode, target, 1.f / (float)n, loss, dpred);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_acgan_loss(const float* logits, const float* labels, int NB, int K, int mode, float target, float aux_weight,
                  float* out4, float* dlogits, void* stream) {
  GP_REQUIRE(logits && labels && out4 && dlogits && NB > 0 && K > 0 && mode >= 0 && mode <= 4 &&
                 (long long)NB * (K + 1) < (1ll << 30),
             "gp_acgan_loss: bad arguments");
  gp::launch_pdl(acgan_loss_kernel, 1, 256, 0, as_stream(stream), logits, labels, NB, K, mode, target, aux_weight, out4, dlogits);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

}  // extern "C"
