// HBM-bound kernels of the hot path: weight packing, BatchNorm (statistics / apply / backward) with fused
// ReLU / LeakyReLU, im2col / col2im for the 3-channel image layers, heads, losses.
// All activations are NHWC bf16 viewed as [P pixels][C channels]; reductions accumulate in fp32.
#include <cuda_bf16.h>

#include "common.h"

namespace gp {

__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == GP_ACT_RELU) return fmaxf(v, 0.f);
  if (act == GP_ACT_LRELU) return v > 0.f ? v : 0.2f * v;
  if (act == GP_ACT_TANH) return tanhf(v);
  return v;
}
// derivative w.r.t. the pre-activation z, given z (ReLU / LeakyReLU only need its sign)
__device__ __forceinline__ float act_grad(float z, int act) {
  if (act == GP_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == GP_ACT_LRELU) return z > 0.f ? 1.f : 0.2f;
  if (act == GP_ACT_TANH) {
    float t = tanhf(z);
    return 1.f - t * t;
  }
  return 1.f;
}

struct bf16x8 {
  uint4 raw;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ __forceinline__ void pack(const float (&f)[8]) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  }
};

// ------------------------------------------------------------------------------------------------ packing
// dst[r][k] (bf16, row length ld_dst) = scale * src[map(r) * s_r + k * s_k] for r < R, k < K; zero elsewhere.
// perm > 1 permutes rows: dst row r reads source row (r % (R/perm)) * perm + r / (R/perm)
// (NCHW-flatten -> NHWC-flatten of the generator's first Linear, models/dcgan.py:50-51).
__global__ void pack_matrix_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int R, int K,
                                   int Rpad, int ld_dst, long long s_r, long long s_k, int perm,
                                   const float* __restrict__ inv_scale) {
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

// conv weight (D0, D1, KH*KW) fp32 -> packed bf16 [N][tap][C] with (N, C) = (D0, D1) if n_dim == 0 else (D1, D0)
__global__ void pack_conv_weight_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int D0, int D1,
                                        int taps, int n_dim_flags, const float* __restrict__ inv_scale) {
  const int n_dim = n_dim_flags & 1;
  const bool flip = (n_dim_flags & 2) != 0;  // reversed tap order: dgrad of a stride-1 conv is a conv with the flipped kernel
  const int N = n_dim == 0 ? D0 : D1, C = n_dim == 0 ? D1 : D0;
  const long long total = (long long)N * taps * C;
  const float sc = inv_scale ? 1.f / __ldg(inv_scale) : 1.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int t = (int)((i / C) % taps);
    const int n = (int)(i / ((long long)C * taps));
    const int d0 = n_dim == 0 ? n : c, d1 = n_dim == 0 ? c : n;
    dst[i] = __float2bfloat16(__ldg(src + ((long long)d0 * D1 + d1) * taps + (flip ? taps - 1 - t : t)) * sc);
  }
}

// packed fp32 gradient [M][tap][N] -> torch layout (M, N, tap) fp32
__global__ void unpack_conv_wgrad_kernel(const float* __restrict__ src, float* __restrict__ dst, int M, int N, int taps) {
  const long long total = (long long)M * N * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const int n = (int)((i / taps) % N);
    const int m = (int)(i / ((long long)taps * N));
    dst[i] = src[((long long)m * taps + t) * N + n];
  }
}

// fp32 [Rpad][ld_src] (gradient of a packed matrix) -> dst[map(r) * s_r + k * s_k] for r < R, k < K (inverse of pack_matrix)
__global__ void unpack_matrix_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int K, int ld_src,
                                     long long s_r, long long s_k, int perm) {
  const long long total = (long long)R * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / K), k = (int)(i % K);
    int rs = r;
    if (perm > 1) {
      const int inner = R / perm;
      rs = (r % inner) * perm + r / inner;
    }
    dst[rs * s_r + k * s_k] = src[(long long)r * ld_src + k];
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm
// Column-reduction thread layout shared by bn_stats / bn_bwd_reduce / colsum: a block of 256 threads covers
// (C/8 column groups) x (256 / (C/8) row lanes) when C/8 divides 256, else grid.y tiles the column groups.
// Partial sums are combined in shared memory first (one global atomic per channel per block).
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
// smem: NQ * cols floats, cols = number of channels covered by this block.
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

// Per-channel sum and sum of squares over P rows.
__global__ void bn_stats_kernel(const __nv_bfloat16* __restrict__ x, long long P, int C, float* __restrict__ sum,
                                float* __restrict__ sumsq, int rows_per_block) {
  const ColLayout L = col_layout(C);
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.f;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > P) r1 = P;
  if (L.active) {
    for (long long r = r0 + L.rl; r < r1; r += L.lanes) {
      bf16x8 v;
      v.load(x + r * C + L.g * 8);
      float f[8];
      v.unpack(f);
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

// mean / rstd / fused scale & shift, running statistics (torch: aten::native_batch_norm semantics:
// biased variance to normalise, unbiased into running_var, momentum 0.1, num_batches_tracked += 1).
__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq, double count, int C,
                                   float eps, float momentum, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ scale, float* __restrict__ shift, float* running_mean,
                                   float* running_var, long long* num_batches_tracked) {
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
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float r = rsqrtf(rv[c] + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  mean[c] = rm[c];
  rstd[c] = r;
  scale[c] = g * r;
  shift[c] = b - rm[c] * g * r;
}

// out = act(y * scale[c] + shift[c])
// Each thread owns one group of 8 channels (per-channel parameters live in registers) and strides over rows.
__global__ void bn_apply_act_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ out, long long P,
                                    int C, const float* __restrict__ scale, const float* __restrict__ shift, int act,
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
    bf16x8 v;
    v.load(y + r * C + L.g * 8);
    float f[8];
    v.unpack(f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = act_fwd(f[j] * sc[j] + sh[j], act);
    v.pack(f);
    v.store(out + r * C + L.g * 8);
  }
}

// Backward reduction: sum_dz[c] = sum dz, sum_dzx[c] = sum dz * xhat, with z = y*scale+shift, dz = da*act'(z),
// xhat = (y - mean) * rstd.   Same thread layout as bn_stats_kernel.
__global__ void bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ y,
                                     long long P, int C, const float* __restrict__ scale,
                                     const float* __restrict__ shift, const float* __restrict__ mean,
                                     const float* __restrict__ rstd, int act, float* __restrict__ sum_dz,
                                     float* __restrict__ sum_dzx, int rows_per_block) {
  const ColLayout L = col_layout(C);
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (L.active) {
    float sc[8], sh[8], mu[8], rs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sc[i] = scale[L.g * 8 + i];
      sh[i] = shift[L.g * 8 + i];
      mu[i] = mean[L.g * 8 + i];
      rs[i] = rstd[L.g * 8 + i];
    }
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > P) r1 = P;
    for (long long r = r0 + L.rl; r < r1; r += L.lanes) {
      bf16x8 vy, vd;
      vy.load(y + r * C + L.g * 8);
      vd.load(da + r * C + L.g * 8);
      float fy[8], fd[8];
      vy.unpack(fy);
      vd.unpack(fd);
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

// dy = gamma*rstd * (dz - sum_dz/M - xhat * sum_dzx/M)
__global__ void bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ y,
                                    __nv_bfloat16* __restrict__ dy, long long n8, int C,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ sum_dz, const float* __restrict__ sum_dzx,
                                    float inv_count, int act, int rows_per_block) {
  const long long P = n8;  // rows
  const ColLayout L = col_layout(C);
  if (!L.active) return;
  // dy = sc*dz - k0 - (y - mu) * k1   with k0 = sc*sum_dz/M, k1 = sc*rstd*sum_dzx/M
  float sc[8], sh[8], mu[8], k0[8], k1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = L.g * 8 + j;
    sc[j] = scale[c];
    sh[j] = shift[c];
    mu[j] = mean[c];
    k0[j] = sc[j] * sum_dz[c] * inv_count;
    k1[j] = sc[j] * rstd[c] * sum_dzx[c] * inv_count;
  }
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > P) r1 = P;
  for (long long r = r0 + L.rl; r < r1; r += L.lanes) {
    bf16x8 vy, vd;
    vy.load(y + r * C + L.g * 8);
    vd.load(da + r * C + L.g * 8);
    float fy[8], fd[8], o[8];
    vy.unpack(fy);
    vd.unpack(fd);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dz = fd[j] * act_grad(fy[j] * sc[j] + sh[j], act);
      o[j] = sc[j] * dz - k0[j] - (fy[j] - mu[j]) * k1[j];
    }
    vd.pack(o);
    vd.store(dy + r * C + L.g * 8);
  }
}

// dy = da * act'(a)   (activation applied directly on the conv output: a has the sign of the pre-activation)
__global__ void act_bwd_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ a,
                               __nv_bfloat16* __restrict__ dy, long long n8, int act) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    bf16x8 va, vd;
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
    vd.pack(fd);
    vd.store(dy + i * 8);
  }
}

// column sums of a bf16 [P][C] matrix into fp32 (bias gradients)
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ x, long long P, int C, float* __restrict__ out,
                              int rows_per_block) {
  const ColLayout L = col_layout(C);
  float acc[1][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = 0.f;
  if (L.active) {
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > P) r1 = P;
    for (long long r = r0 + L.rl; r < r1; r += L.lanes) {
      bf16x8 v;
      v.load(x + r * C + L.g * 8);
      float f[8];
      v.unpack(f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[0][i] += f[i];
    }
  }
  float* const outs[1] = {out};
  col_flush<1>(L, C, acc, outs);
}

// ------------------------------------------------------------------------------------------------ image layers
// im2col of a k4 s2 p1 window over an NCHW fp32 image with `ch` (<= 4) channels:
//   col[(n, oh, ow)][(c*4 + kh)*4 + kw] = img[n, c, 2oh-1+kh, 2ow-1+kw] * (mul ? 1 - mul[same]^2 : 1); columns >= ch*16 are 0.
// Used for D's first conv (models/dcgan.py:106, Cin = img_dim) and for the gradient of G's last ConvT + Tanh
// (models/dcgan.py:41-44): `mul` is then the tanh output, fusing tanh'.
__global__ void im2col_k4s2_kernel(const float* __restrict__ img, const float* __restrict__ mul,
                                   __nv_bfloat16* __restrict__ col, int NB, int ch, int Hi, int Wi) {
  const int Ho = Hi / 2, Wo = Wi / 2;
  const long long total = (long long)NB * Ho * Wo * 8;  // 8 groups of 8 columns per pixel
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
      float v = 0.f;
      if (c < ch && ih >= 0 && ih < Hi && iw >= 0 && iw < Wi) {
        const long long off = (((long long)n * ch + c) * Hi + ih) * Wi + iw;
        v = __ldg(img + off);
        if (mul != nullptr) {
          const float t = __ldg(mul + off);
          v *= 1.f - t * t;
        }
      }
      f[j] = v;
    }
    bf16x8 o;
    o.pack(f);
    o.store(col + i * 8);
  }
}

// col2im (transpose of the above): img[n, c, ih, iw] = act( bias[c] + sum_{(oh,kh): 2oh-1+kh = ih} sum_{(ow,kw)} col[(n,oh,ow)][(c*4+kh)*4+kw] )
// Used for G's last ConvT (+Tanh) forward and for the image gradient of D's first conv.
__global__ void col2im_k4s2_kernel(const __nv_bfloat16* __restrict__ col, const float* __restrict__ bias,
                                   float* __restrict__ img, int NB, int ch, int Hi, int Wi, int act) {
  const int Ho = Hi / 2, Wo = Wi / 2;
  const long long total = (long long)NB * ch * Hi * Wi;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int iw = (int)(i % Wi), ih = (int)((i / Wi) % Hi);
    const int c = (int)((i / ((long long)Wi * Hi)) % ch), n = (int)(i / ((long long)Wi * Hi * ch));
    float acc = bias ? __ldg(bias + c) : 0.f;
    // ih = 2*oh - 1 + kh  ->  kh in {(ih+1)&1, (ih+1)&1 + 2}
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int kh = ((ih + 1) & 1) + 2 * a;
      const int oh2 = ih + 1 - kh;
      if (oh2 < 0 || oh2 >= 2 * Ho) continue;
      const int oh = oh2 / 2;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int kw = ((iw + 1) & 1) + 2 * b;
        const int ow2 = iw + 1 - kw;
        if (ow2 < 0 || ow2 >= 2 * Wo) continue;
        const int ow = ow2 / 2;
        acc += __bfloat162float(col[(((long long)n * Ho + oh) * Wo + ow) * 64 + (c * 4 + kh) * 4 + kw]);
      }
    }
    img[i] = act_fwd(acc, act);
  }
}

// dbias[c] += sum_{n,h,w} dout[n,c,h,w] * (mul ? 1 - mul^2 : 1)   (bias gradient of the last ConvT under Tanh)
__global__ void image_bias_grad_kernel(const float* __restrict__ dout, const float* __restrict__ mul,
                                       float* __restrict__ dbias, int NB, int ch, int HW) {
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
// out[b][o] = bias[o] + sum_{hw, c} a[b, hw, c] * w[o*s_o + c*s_c + hw*s_hw]
// (sum-pool + Linear: s_hw = 0, models/dcgan.py:121-122; flatten + Linear: s_c = HW, s_hw = 1, dcgan_specnorm.py:125-126)
__global__ void head_fwd_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ w,
                                const float* __restrict__ bias, float* __restrict__ out, int HW, int C, int O,
                                long long s_o, long long s_c, long long s_hw) {
  const int b = blockIdx.x, o = blockIdx.y;
  float acc = 0.f;
  const __nv_bfloat16* ab = a + (long long)b * HW * C;
  for (int i = threadIdx.x; i < HW * C; i += blockDim.x) {
    const int c = i % C, hw = i / C;
    acc += __bfloat162float(ab[i]) * __ldg(w + o * s_o + c * s_c + hw * s_hw);
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

// da[b, hw, c] = sum_o dout[b][o] * w[o, c, hw]
__global__ void head_bwd_data_kernel(const float* __restrict__ dout, const float* __restrict__ w,
                                     __nv_bfloat16* __restrict__ da, int NB, int HW, int C, int O, long long s_o,
                                     long long s_c, long long s_hw) {
  const long long total = (long long)NB * HW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), hw = (int)((i / C) % HW);
    const long long b = i / ((long long)C * HW);
    float acc = 0.f;
    for (int o = 0; o < O; ++o) acc += __ldg(dout + b * O + o) * __ldg(w + o * s_o + c * s_c + hw * s_hw);
    da[i] = __float2bfloat16(acc);
  }
}

// dw[o, c, hw] (+)= sum_b dout[b][o] * a[b, hw, c]  (sum-pool: s_hw = 0 so the hw terms accumulate); dbias[o] = sum_b dout[b][o]
// One block per (c-chunk, o); dw must be zeroed by the caller.
__global__ void head_bwd_weight_kernel(const float* __restrict__ dout, const __nv_bfloat16* __restrict__ a,
                                       float* __restrict__ dw, float* __restrict__ dbias, int NB, int HW, int C, int O,
                                       long long s_o, long long s_c, long long s_hw) {
  const int o = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // index over HW*C
  const int bchunk = (NB + gridDim.z - 1) / gridDim.z;
  const int b0 = blockIdx.z * bchunk, b1 = min(b0 + bchunk, NB);
  if (i < HW * C) {
    const int c = i % C, hw = i / C;
    float acc = 0.f;
    for (int b = b0; b < b1; ++b) acc += __ldg(dout + (long long)b * O + o) * __bfloat162float(a[((long long)b * HW) * C + i]);
    atomicAdd(dw + o * s_o + c * s_c + hw * s_hw, acc);
  }
  if (dbias != nullptr && blockIdx.x == 0 && blockIdx.z == 0 && threadIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < NB; ++b) s += dout[(long long)b * O + o];
    dbias[o] = s;
  }
}

// ------------------------------------------------------------------------------------------------ losses
// GANLoss (utils/criterion.py:30-41), value and d(loss)/d(pred) in one pass; single block.
// mode 0: BCE-with-logits vs constant target; 1: MSE vs constant target; 2: hinge real relu(1-p); 3: hinge fake relu(1+p); 4: -p
__global__ void gan_loss_kernel(const float* __restrict__ pred, int n, int mode, float target, float inv_n,
                                float* __restrict__ loss, float* __restrict__ dpred) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = pred[i];
    float l, d;
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
    acc += l;
    dpred[i] = d * inv_n;
  }
  __shared__ float red[32];
  for (int k = 16; k > 0; k >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, k);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int k = 16; k > 0; k >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, k);
    if (threadIdx.x == 0) *loss = acc * inv_n;
  }
}

static inline int grid_for(long long n, int block = 256, int max_blocks = 148 * 16) {
  long long g = (n + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return (int)g;
}

// choose rows-per-block for the column-reduction kernels so the grid is ~8 blocks per SM
static inline int rows_per_block_for(long long P) {
  long long target_blocks = (long long)num_sms() * 8;
  long long rpb = (P + target_blocks - 1) / target_blocks;
  if (rpb < 32) rpb = 32;
  return (int)rpb;
}

}  // namespace gp

using namespace gp;

extern "C" {

int gp_pack_matrix(const float* src, void* dst, int R, int K, int Rpad, int ld_dst, long long s_r, long long s_k,
                   int perm, const float* inv_scale, void* stream) {
  GP_REQUIRE(src && dst && R > 0 && K > 0 && Rpad >= R && ld_dst >= K, "gp_pack_matrix: bad arguments");
  GP_REQUIRE(perm <= 1 || R % perm == 0, "gp_pack_matrix: perm must divide R");
  pack_matrix_kernel<<<grid_for((long long)Rpad * ld_dst), 256, 0, as_stream(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), R, K, Rpad, ld_dst, s_r, s_k, perm, inv_scale);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_unpack_matrix(const float* src, float* dst, int R, int K, int ld_src, long long s_r, long long s_k, int perm,
                     void* stream) {
  GP_REQUIRE(src && dst && R > 0 && K > 0 && ld_src >= K, "gp_unpack_matrix: bad arguments");
  unpack_matrix_kernel<<<grid_for((long long)R * K), 256, 0, as_stream(stream)>>>(src, dst, R, K, ld_src, s_r, s_k, perm);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_pack_conv_weight(const float* src, void* dst, int D0, int D1, int taps, int n_dim, const float* inv_scale,
                        void* stream) {
  GP_REQUIRE(src && dst && D0 > 0 && D1 > 0 && taps > 0 && n_dim >= 0 && n_dim <= 3, "gp_pack_conv_weight: bad arguments");
  pack_conv_weight_kernel<<<grid_for((long long)D0 * D1 * taps), 256, 0, as_stream(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), D0, D1, taps, n_dim, inv_scale);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_unpack_conv_wgrad(const float* src, float* dst, int M, int N, int taps, void* stream) {
  GP_REQUIRE(src && dst && M > 0 && N > 0 && taps > 0, "gp_unpack_conv_wgrad: bad arguments");
  unpack_conv_wgrad_kernel<<<grid_for((long long)M * N * taps), 256, 0, as_stream(stream)>>>(src, dst, M, N, taps);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

// launch geometry of the column-reduction kernels
struct ColLaunch {
  dim3 grid;
  int block;
  size_t smem;
  int rpb;
};
static ColLaunch col_launch(long long P, int C, int nq) {
  ColLaunch L;
  L.rpb = rows_per_block_for(P);
  const int gx = (int)((P + L.rpb - 1) / L.rpb);
  const int cgs = C / 8;
  if (cgs <= 256 && 256 % cgs == 0) {
    L.grid = dim3(gx, 1);
    L.block = 256;
    L.smem = (size_t)nq * C * sizeof(float);
  } else {
    L.block = 128;
    int gy = (cgs + 127) / 128;
    if (gy < 2) gy = 2;  // gridDim.y > 1 selects the tiled layout inside the kernels
    L.grid = dim3(gx, gy);
    L.smem = (size_t)nq * 128 * 8 * sizeof(float);
  }
  return L;
}

int gp_bn_stats(const void* x, long long P, int C, float* sum, float* sumsq, void* stream) {
  GP_REQUIRE(x && sum && sumsq && P > 0 && C > 0 && C % 8 == 0, "gp_bn_stats: bad arguments (C %% 8 == 0 required)");
  const ColLaunch L = col_launch(P, C, 2);
  bn_stats_kernel<<<L.grid, L.block, L.smem, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x), P, C, sum, sumsq,
                                                                  L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_finalize(const float* sum, const float* sumsq, double count, int C, float eps, float momentum,
                   const float* gamma, const float* beta, float* mean, float* rstd, float* scale, float* shift,
                   float* running_mean, float* running_var, long long* num_batches_tracked, void* stream) {
  GP_REQUIRE(sum && sumsq && mean && rstd && scale && shift && C > 0 && count > 0, "gp_bn_finalize: bad arguments");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(sum, sumsq, count, C, eps, momentum, gamma, beta,
                                                                    mean, rstd, scale, shift, running_mean,
                                                                    running_var, num_batches_tracked);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_eval_params(const float* running_mean, const float* running_var, const float* gamma, const float* beta, int C,
                      float eps, float* mean, float* rstd, float* scale, float* shift, void* stream) {
  GP_REQUIRE(running_mean && running_var && mean && rstd && scale && shift && C > 0, "gp_bn_eval_params: bad arguments");
  bn_eval_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(running_mean, running_var, gamma, beta, C, eps, mean,
                                                                rstd, scale, shift);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_apply_act(const void* y, void* out, long long P, int C, const float* scale, const float* shift, int act,
                    void* stream) {
  GP_REQUIRE(y && out && scale && shift && P > 0 && C % 8 == 0, "gp_bn_apply_act: bad arguments");
  const ColLaunch L = col_launch(P, C, 0);
  bn_apply_act_kernel<<<L.grid, L.block, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(y),
                                                                 static_cast<__nv_bfloat16*>(out), P, C, scale, shift,
                                                                 act, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_bwd_reduce(const void* da, const void* y, long long P, int C, const float* scale, const float* shift,
                     const float* mean, const float* rstd, int act, float* sum_dz, float* sum_dzx, void* stream) {
  GP_REQUIRE(da && y && sum_dz && sum_dzx && P > 0 && C % 8 == 0, "gp_bn_bwd_reduce: bad arguments");
  const ColLaunch L = col_launch(P, C, 2);
  bn_bwd_reduce_kernel<<<L.grid, L.block, L.smem, as_stream(stream)>>>(
      static_cast<const __nv_bfloat16*>(da), static_cast<const __nv_bfloat16*>(y), P, C, scale, shift, mean, rstd, act,
      sum_dz, sum_dzx, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_bwd_apply(const void* da, const void* y, void* dy, long long P, int C, const float* scale,
                    const float* shift, const float* mean, const float* rstd, const float* sum_dz,
                    const float* sum_dzx, double count, int act, void* stream) {
  GP_REQUIRE(da && y && dy && P > 0 && C % 8 == 0 && count > 0, "gp_bn_bwd_apply: bad arguments");
  const ColLaunch L = col_launch(P, C, 0);
  bn_bwd_apply_kernel<<<L.grid, L.block, 0, as_stream(stream)>>>(
      static_cast<const __nv_bfloat16*>(da), static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(dy), P, C,
      scale, shift, mean, rstd, sum_dz, sum_dzx, (float)(1.0 / count), act, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_act_bwd(const void* da, const void* a, void* dy, long long n, int act, void* stream) {
  GP_REQUIRE(da && a && dy && n > 0 && n % 8 == 0, "gp_act_bwd: bad arguments");
  act_bwd_kernel<<<grid_for(n / 8), 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(da),
                                                                 static_cast<const __nv_bfloat16*>(a),
                                                                 static_cast<__nv_bfloat16*>(dy), n / 8, act);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_colsum(const void* x, long long P, int C, float* out, void* stream) {
  GP_REQUIRE(x && out && P > 0 && C % 8 == 0, "gp_colsum: bad arguments");
  const ColLaunch L = col_launch(P, C, 1);
  colsum_kernel<<<L.grid, L.block, L.smem, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x), P, C, out, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_im2col_k4s2(const float* img, const float* mul, void* col, int NB, int ch, int Hi, int Wi, void* stream) {
  GP_REQUIRE(img && col && NB > 0 && ch > 0 && ch <= 4 && Hi % 2 == 0 && Wi % 2 == 0, "gp_im2col_k4s2: bad arguments");
  const long long total = (long long)NB * (Hi / 2) * (Wi / 2) * 8;
  im2col_k4s2_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(img, mul, static_cast<__nv_bfloat16*>(col), NB, ch,
                                                                     Hi, Wi);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_col2im_k4s2(const void* col, const float* bias, float* img, int NB, int ch, int Hi, int Wi, int act,
                   void* stream) {
  GP_REQUIRE(img && col && NB > 0 && ch > 0 && ch <= 4 && Hi % 2 == 0 && Wi % 2 == 0, "gp_col2im_k4s2: bad arguments");
  const long long total = (long long)NB * ch * Hi * Wi;
  col2im_k4s2_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(col), bias, img,
                                                                     NB, ch, Hi, Wi, act);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_image_bias_grad(const float* dout, const float* mul, float* dbias, int NB, int ch, int HW, void* stream) {
  GP_REQUIRE(dout && dbias && NB > 0 && ch > 0 && HW > 0, "gp_image_bias_grad: bad arguments");
  dim3 grid(grid_for((long long)NB * HW, 256, 148 * 2), ch);
  image_bias_grad_kernel<<<grid, 256, 0, as_stream(stream)>>>(dout, mul, dbias, NB, ch, HW);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_head_fwd(const void* a, const float* w, const float* bias, float* out, int NB, int HW, int C, int O,
                long long s_o, long long s_c, long long s_hw, void* stream) {
  GP_REQUIRE(a && w && out && NB > 0 && HW > 0 && C > 0 && O > 0, "gp_head_fwd: bad arguments");
  dim3 grid(NB, O);
  head_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(a), w, bias, out, HW, C, O,
                                                       s_o, s_c, s_hw);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_head_bwd(const float* dout, const void* a, const float* w, void* da, float* dw, float* dbias, int NB, int HW,
                int C, int O, long long s_o, long long s_c, long long s_hw, void* stream) {
  GP_REQUIRE(dout && a && w && NB > 0 && HW > 0 && C > 0 && O > 0, "gp_head_bwd: bad arguments");
  if (da != nullptr) {
    head_bwd_data_kernel<<<grid_for((long long)NB * HW * C), 256, 0, as_stream(stream)>>>(
        dout, w, static_cast<__nv_bfloat16*>(da), NB, HW, C, O, s_o, s_c, s_hw);
    GP_CHECK_LAUNCH();
  }
  if (dw != nullptr) {
    int zsplit = (4 * num_sms()) / (((HW * C + 255) / 256) * O);
    if (zsplit < 1) zsplit = 1;
    if (zsplit > 32) zsplit = 32;
    if (zsplit > NB) zsplit = NB;
    dim3 grid((HW * C + 255) / 256, O, zsplit);
    head_bwd_weight_kernel<<<grid, 256, 0, as_stream(stream)>>>(dout, static_cast<const __nv_bfloat16*>(a), dw, dbias, NB,
                                                                HW, C, O, s_o, s_c, s_hw);
    GP_CHECK_LAUNCH();
  }
  return GP_OK;
}

int gp_gan_loss(const float* pred, int n, int mode, float target, float* loss, float* dpred, void* stream) {
  GP_REQUIRE(pred && loss && dpred && n > 0 && mode >= 0 && mode <= 4, "gp_gan_loss: bad arguments");
  gan_loss_kernel<<<1, 256, 0, as_stream(stream)>>>(pred, n, mode, target, 1.f / (float)n, loss, dpred);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

}  // extern "C"
