// HBM-bound kernels specific to the SNGAN projection networks (reference: models/sngan_projection.py):
// conditional BatchNorm (embedding gather fused into the scale/shift pass, optional fused nearest x2 upsample),
// 2x2 pooling / upsampling, 3x3 image-side im2col / col2im, ReLU + global sum pooling, the projection head.
#include <cuda_bf16.h>

#include "act_io.cuh"
#include "common.h"

namespace gp {

__device__ __forceinline__ float r_act_fwd(float v, int act) {
  if (act == GP_ACT_RELU) return fmaxf(v, 0.f);
  if (act == GP_ACT_LRELU) return v > 0.f ? v : 0.2f * v;
  return v;
}
__device__ __forceinline__ float r_act_grad(float z, int act) {
  if (act == GP_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == GP_ACT_LRELU) return z > 0.f ? 1.f : 0.2f;
  return 1.f;
}

__device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return raw;
}

// ---------------------------------------------------------------------------------------- conditional BatchNorm
// out[n, (up)h, (up)w, c] = act( (y - mean[c]) * rstd[c] * gamma[n][c] + beta[n][c] ),
// gamma[n] = emb[label[n]][0:C], beta[n] = emb[label[n]][C:2C]   (models/sngan_projection.py:15-19).
// emb == NULL: plain non-affine normalisation. grid = (pixel chunks, NB); thread = one 8-channel group x pixel lane.
// y_comp / out_comp / fmt: companion tensors of the forward precision mode (act_io.cuh).
__global__ void cbn_apply_kernel(const __nv_bfloat16* __restrict__ y, const void* __restrict__ y_comp,
                                 __nv_bfloat16* __restrict__ out, void* __restrict__ out_comp, int fmt, int H, int W,
                                 int C, const float* __restrict__ mean, const float* __restrict__ rstd,
                                 const float* __restrict__ emb, const long long* __restrict__ labels, int act, int up,
                                 int px_per_block) {
  gp::pdl_sync();
  const int n = blockIdx.y;
  const int cgs = C / 8;
  const int g = threadIdx.x % cgs, lane = threadIdx.x / cgs, lanes = blockDim.x / cgs;
  if (lane >= lanes) return;
  float sc[8], sh[8];
  const float* e = emb ? emb + (long long)labels[n] * 2 * C : nullptr;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    const float gm = e ? e[c] : 1.f, bt = e ? e[C + c] : 0.f;
    sc[j] = rstd[c] * gm;
    sh[j] = bt - mean[c] * rstd[c] * gm;
  }
  const int HW = H * W;
  const int p0 = blockIdx.x * px_per_block, p1 = min(p0 + px_per_block, HW);
  const int oW = up ? 2 * W : W;
  for (int p = p0 + lane; p < p1; p += lanes) {
    float f[8];
    load8c(y, y_comp, fmt, ((long long)n * HW + p) * C + g * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = r_act_fwd(f[j] * sc[j] + sh[j], act);
    const Packed8c o = pack8c(out_comp != nullptr ? fmt : GP_COMP_NONE, f);
    if (up) {
      const int h = p / W, w = p % W;
      const long long base = (((long long)n * 2 * H + 2 * h) * oW + 2 * w) * C + g * 8;
      put8c(out, out_comp, base, o);
      put8c(out, out_comp, base + C, o);
      put8c(out, out_comp, base + (long long)oW * C, o);
      put8c(out, out_comp, base + (long long)oW * C + C, o);
    } else {
      put8c(out, out_comp, ((long long)n * HW + p) * C + g * 8, o);
    }
  }
}

// da of the (possibly upsampled) output gathered back onto the small grid (sum of the 2x2 copies)
__device__ __forceinline__ void load_da(const __nv_bfloat16* __restrict__ da, int n, int p, int H, int W, int C, int g,
                                        int up, float (&f)[8]) {
  if (!up) {
    unpack8(*reinterpret_cast<const uint4*>(da + ((long long)n * H * W + p) * C + g * 8), f);
    return;
  }
  const int h = p / W, w = p % W, oW = 2 * W;
  const __nv_bfloat16* base = da + (((long long)n * 2 * H + 2 * h) * oW + 2 * w) * C + g * 8;
  float t[8];
  unpack8(*reinterpret_cast<const uint4*>(base), f);
  unpack8(*reinterpret_cast<const uint4*>(base + C), t);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] += t[j];
  unpack8(*reinterpret_cast<const uint4*>(base + (long long)oW * C), t);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] += t[j];
  unpack8(*reinterpret_cast<const uint4*>(base + (long long)oW * C + C), t);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] += t[j];
}

// per-sample sums: part[n][0][c] += sum_hw dz, part[n][1][c] += sum_hw dz * xhat   (dz = da * act'(z))
__global__ void cbn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ y,
                                      const void* __restrict__ y_comp, int fmt, int H, int W, int C, const float* __restrict__ mean, const float* __restrict__ rstd,
                                      const float* __restrict__ emb, const long long* __restrict__ labels, int act,
                                      int up, float* __restrict__ part, int px_per_block) {
  gp::pdl_sync();
  const int n = blockIdx.y;
  const int cgs = C / 8;
  const int g = threadIdx.x % cgs, lane = threadIdx.x / cgs, lanes = blockDim.x / cgs;
  float a0[8], a1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a0[j] = a1[j] = 0.f;
  if (lane < lanes) {
    float gm[8], bt[8], mu[8], rs[8];
    const float* e = emb ? emb + (long long)labels[n] * 2 * C : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      gm[j] = e ? e[c] : 1.f;
      bt[j] = e ? e[C + c] : 0.f;
      mu[j] = mean[c];
      rs[j] = rstd[c];
    }
    const int HW = H * W;
    const int p0 = blockIdx.x * px_per_block, p1 = min(p0 + px_per_block, HW);
    for (int p = p0 + lane; p < p1; p += lanes) {
      float fy[8], fd[8];
      load8c(y, y_comp, fmt, ((long long)n * HW + p) * C + g * 8, fy);   // the value the forward normalised
      load_da(da, n, p, H, W, C, g, up, fd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (fy[j] - mu[j]) * rs[j];
        const float dz = fd[j] * r_act_grad(xh * gm[j] + bt[j], act);
        a0[j] += dz;
        a1[j] += dz * xh;
      }
    }
  }
  extern __shared__ float s_red[];  // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  if (lane < lanes) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&s_red[g * 8 + j], a0[j]);
      atomicAdd(&s_red[C + g * 8 + j], a1[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(part + (long long)n * 2 * C + i, s_red[i]);
}

// S[0][c] = sum_n gamma[n][c] * part[n][0][c];  S[1][c] = sum_n gamma[n][c] * part[n][1][c];
// demb[label[n]][c] += part[n][1][c] (d gamma), demb[label[n]][C + c] += part[n][0][c] (d beta)
__global__ void cbn_bwd_finalize_kernel(const float* __restrict__ part, int NB, int C, const float* __restrict__ emb,
                                        const long long* __restrict__ labels, float* __restrict__ S,
                                        float* __restrict__ demb) {
  gp::pdl_sync();
  // grid (channel blocks, sample slices): S is zeroed by the caller and receives one atomic per slice
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int per = (NB + gridDim.y - 1) / gridDim.y;
  const int n0 = blockIdx.y * per, n1 = min(n0 + per, NB);
  float s0 = 0.f, s1 = 0.f;
  for (int n = n0; n < n1; ++n) {
    const float a = part[(long long)n * 2 * C + c], b = part[(long long)n * 2 * C + C + c];
    const long long lb = emb ? labels[n] : 0;
    const float gm = emb ? emb[lb * 2 * C + c] : 1.f;
    s0 += gm * a;
    s1 += gm * b;
    if (demb != nullptr) {
      atomicAdd(demb + lb * 2 * C + c, b);
      atomicAdd(demb + lb * 2 * C + C + c, a);
    }
  }
  if (n1 > n0) {
    atomicAdd(S + c, s0);
    atomicAdd(S + C + c, s1);
  }
}

// dy[n,hw,c] = rstd[c] * (gamma[n][c] * dz - S0[c]/M - xhat * S1[c]/M)
__global__ void cbn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ y,
                                     const void* __restrict__ y_comp, int fmt, __nv_bfloat16* __restrict__ dy, int H,
                                     int W, int C,
                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                     const float* __restrict__ emb, const long long* __restrict__ labels,
                                     const float* __restrict__ S, float inv_count, int act, int up, int px_per_block) {
  gp::pdl_sync();
  const int n = blockIdx.y;
  const int cgs = C / 8;
  const int g = threadIdx.x % cgs, lane = threadIdx.x / cgs, lanes = blockDim.x / cgs;
  if (lane >= lanes) return;
  float gm[8], bt[8], mu[8], rs[8], k0[8], k1[8];
  const float* e = emb ? emb + (long long)labels[n] * 2 * C : nullptr;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    gm[j] = e ? e[c] : 1.f;
    bt[j] = e ? e[C + c] : 0.f;
    mu[j] = mean[c];
    rs[j] = rstd[c];
    k0[j] = S[c] * inv_count;
    k1[j] = S[C + c] * inv_count;
  }
  const int HW = H * W;
  const int p0 = blockIdx.x * px_per_block, p1 = min(p0 + px_per_block, HW);
  for (int p = p0 + lane; p < p1; p += lanes) {
    float fy[8], fd[8], o[8];
    load8c(y, y_comp, fmt, ((long long)n * HW + p) * C + g * 8, fy);
    load_da(da, n, p, H, W, C, g, up, fd);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (fy[j] - mu[j]) * rs[j];
      const float dz = fd[j] * r_act_grad(xh * gm[j] + bt[j], act);
      o[j] = rs[j] * (gm[j] * dz - k0[j] - xh * k1[j]);
    }
    *reinterpret_cast<uint4*>(dy + ((long long)n * HW + p) * C + g * 8) = pack8(o);
  }
}

// ---------------------------------------------------------------------------------------- resampling / activation
// out[n, 2h+a, 2w+b, c] = scale * in[n, h, w, c]
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long long NB,
                                  int H, int W, int C, float scale) {
  gp::pdl_sync();
  const int cgs = C / 8;
  const long long total = NB * H * W * cgs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cgs);
    const long long p = i / cgs;
    const int w = (int)(p % W), h = (int)((p / W) % H);
    const long long n = p / ((long long)W * H);
    uint4 o = *reinterpret_cast<const uint4*>(in + p * C + g * 8);
    if (scale != 1.f) {  // scale == 1: a pure 16-bit copy (also used for fp16 companion tensors)
      float f[8];
      unpack8(o, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= scale;
      o = pack8(f);
    }
    __nv_bfloat16* base = out + ((n * 2 * H + 2 * h) * (2LL * W) + 2 * w) * C + g * 8;
    *reinterpret_cast<uint4*>(base) = o;
    *reinterpret_cast<uint4*>(base + C) = o;
    *reinterpret_cast<uint4*>(base + 2LL * W * C) = o;
    *reinterpret_cast<uint4*>(base + 2LL * W * C + C) = o;
  }
}

// out[n, h, w, c] = scale * sum_{a,b} in[n, 2h+a, 2w+b, c]      (H, W: output size)
__global__ void pool2x_kernel(const __nv_bfloat16* __restrict__ in, const void* __restrict__ in_comp,
                              __nv_bfloat16* __restrict__ out, void* __restrict__ out_comp, int fmt, long long NB, int H,
                              int W, int C, float scale) {
  gp::pdl_sync();
  const int cgs = C / 8;
  const long long total = NB * H * W * cgs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cgs);
    const long long p = i / cgs;
    const int w = (int)(p % W), h = (int)((p / W) % H);
    const long long n = p / ((long long)W * H);
    const long long base = ((n * 2 * H + 2 * h) * (2LL * W) + 2 * w) * C + g * 8;
    float f[8], t[8];
    load8c(in, in_comp, fmt, base, f);
    load8c(in, in_comp, fmt, base + C, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += t[j];
    load8c(in, in_comp, fmt, base + 2LL * W * C, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += t[j];
    load8c(in, in_comp, fmt, base + 2LL * W * C + C, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (f[j] + t[j]) * scale;
    store8c(out, out_comp, fmt, p * C + g * 8, f);
  }
}

__global__ void act_fwd_kernel(const __nv_bfloat16* __restrict__ in, const void* __restrict__ in_comp,
                               __nv_bfloat16* __restrict__ out, void* __restrict__ out_comp, int fmt, long long n8,
                               int act) {
  gp::pdl_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    load8c(in, in_comp, fmt, i * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = r_act_fwd(f[j], act);
    store8c(out, out_comp, fmt, i * 8, f);
  }
}

// ---------------------------------------------------------------------------------------- 3x3 image-side layers
// col[(n,h,w)][(c*3+kh)*3+kw] = img[n, c, h+kh-1, w+kw-1] (zero padded), columns >= ch*9 are zero; col row = 32 bf16.
__global__ void im2col_k3s1_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ col,
                                   void* __restrict__ col_comp, int fmt, int NB, int ch, int H, int W) {
  gp::pdl_sync();
  const long long total = (long long)NB * H * W * 4;  // 4 groups of 8 columns
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % 4);
    const long long p = i / 4;
    const int w = (int)(p % W), h = (int)((p / W) % H), n = (int)(p / ((long long)W * H));
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = g * 8 + j;
      const int c = idx / 9, kh = (idx / 3) % 3, kw = idx % 3;
      const int ih = h + kh - 1, iw = w + kw - 1;
      f[j] = (c < ch && ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(img + (((long long)n * ch + c) * H + ih) * W + iw) : 0.f;
    }
    store8c(col, col_comp, fmt, i * 8, f);
  }
}

// img[n,c,ih,iw] = sum_{kh,kw} col[(n, ih-kh+1, iw-kw+1)][(c*3+kh)*3+kw]
__global__ void col2im_k3s1_kernel(const __nv_bfloat16* __restrict__ col, float* __restrict__ img, int NB, int ch, int H,
                                   int W) {
  gp::pdl_sync();
  const long long total = (long long)NB * ch * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int iw = (int)(i % W), ih = (int)((i / W) % H);
    const int c = (int)((i / ((long long)W * H)) % ch), n = (int)(i / ((long long)W * H * ch));
    float acc = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int h = ih - kh + 1;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int w = iw - kw + 1;
        if (w < 0 || w >= W) continue;
        acc += __bfloat162float(col[(((long long)n * H + h) * W + w) * 32 + (c * 3 + kh) * 3 + kw]);
      }
    }
    img[i] = acc;
  }
}

// NHWC bf16 with 8 channels (first `ch` valid) -> NCHW fp32 with optional tanh; and its gradient back.
// in_f32 != 0: `in` is the fp32 output of the GEMM (the precise modes keep the pre-tanh image out of bf16).
__global__ void nhwc8_to_image_kernel(const void* __restrict__ in, int in_f32, float* __restrict__ img, long long NB, int ch,
                                      int HW, int tanh_act) {
  gp::pdl_sync();
  const long long total = NB * ch * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % HW;
    const int c = (int)((i / HW) % ch);
    const long long n = i / ((long long)HW * ch);
    const long long off = (n * HW + p) * 8 + c;
    const float v = in_f32 ? static_cast<const float*>(in)[off] : __bfloat162float(static_cast<const __nv_bfloat16*>(in)[off]);
    img[i] = tanh_act ? tanhf(v) : v;
  }
}
// dy[p][c] = dout[n,c,p] * (1 - out^2) for c < ch, 0 for c >= ch
__global__ void image_to_nhwc8_grad_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                           __nv_bfloat16* __restrict__ dy, long long NB, int ch, int HW, int tanh_act) {
  gp::pdl_sync();
  const long long total = NB * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % HW, n = i / HW;
    float f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float v = 0.f;
      if (c < ch) {
        const long long off = (n * ch + c) * HW + p;
        v = dout[off];
        if (tanh_act) {
          const float t = out[off];
          v *= 1.f - t * t;
        }
      }
      f[c] = v;
    }
    *reinterpret_cast<uint4*>(dy + i * 8) = pack8(f);
  }
}

// ---------------------------------------------------------------------------------------- projection head
// h[n][c] = sum_hw relu(a[n,hw,c])     (models/sngan_projection.py:190-191)
__global__ void relu_sumpool_kernel(const __nv_bfloat16* __restrict__ a, const void* __restrict__ a_comp, int fmt,
                                    float* __restrict__ h, int HW, int C) {
  gp::pdl_sync();
  const int n = blockIdx.y;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;  // 8-channel group
  if (g * 8 >= C) return;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int p = 0; p < HW; ++p) {
    float f[8];
    load8c(a, a_comp, fmt, ((long long)n * HW + p) * C + g * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += fmaxf(f[j], 0.f);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) h[(long long)n * C + g * 8 + j] = acc[j];
}
__global__ void relu_sumpool_bwd_kernel(const float* __restrict__ dh, const __nv_bfloat16* __restrict__ a,
                                        __nv_bfloat16* __restrict__ da, long long total, int HW, int C) {
  gp::pdl_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long n = i / ((long long)C * HW);
    da[i] = __float2bfloat16(__bfloat162float(a[i]) > 0.f ? dh[n * C + c] : 0.f);
  }
}

// out[n] = b + sum_c h[n][c] * (w[c] + E[label[n]][c])      (l6(h) + sum(l_y(y) * h), sngan_projection.py:192-195)
// one warp per sample (GEMV with a gathered embedding row)
__global__ void proj_head_fwd_kernel(const float* __restrict__ h, const float* __restrict__ w, const float* __restrict__ b,
                                     const float* __restrict__ E, const long long* __restrict__ labels,
                                     float* __restrict__ out, int NB, int C) {
  gp::pdl_sync();
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= NB) return;
  const float* e = E ? E + (long long)labels[n] * C : nullptr;
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) acc += h[(long long)n * C + c] * (w[c] + (e ? e[c] : 0.f));
  for (int k = 16; k > 0; k >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, k);
  if (lane == 0) out[n] = acc + (b ? b[0] : 0.f);
}
// dh[n][c] = dout[n] * (w[c] + E[label[n]][c]); dw[c] += sum_n dout[n] h[n][c]; dE[label[n]][c] += dout[n] h[n][c]; db += sum dout
__global__ void proj_head_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ h,
                                     const float* __restrict__ w, const float* __restrict__ E,
                                     const long long* __restrict__ labels, float* __restrict__ dh,
                                     float* __restrict__ dw, float* __restrict__ db, float* __restrict__ dE, int NB,
                                     int C) {
  gp::pdl_sync();
  // grid (channel blocks, sample slices): dw / db are zeroed by the caller and receive one atomic per slice
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int per = (NB + gridDim.y - 1) / gridDim.y;
  const int n0 = blockIdx.y * per, n1 = min(n0 + per, NB);
  float accw = 0.f, accb = 0.f;
  for (int n = n0; n < n1; ++n) {
    const float d = dout[n];
    const float hv = h[(long long)n * C + c];
    const long long lb = E ? labels[n] : 0;
    dh[(long long)n * C + c] = d * (w[c] + (E ? E[lb * C + c] : 0.f));
    accw += d * hv;
    if (dE != nullptr) atomicAdd(dE + lb * C + c, d * hv);
    accb += d;
  }
  if (n1 > n0) {
    atomicAdd(dw + c, accw);
    if (c == 0 && db != nullptr) atomicAdd(db, accb);
  }
}

static inline int grid1(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// slices of ~16 samples for the per-channel kernels that loop over the batch
static inline int sample_slices(int NB) { return NB >= 32 ? (NB + 15) / 16 : 1; }

struct CbnLaunch {
  dim3 grid;
  int block, ppb;
};
static CbnLaunch cbn_launch(int NB, int HW, int C) {
  CbnLaunch L;
  L.block = 256;
  const int cgs = C / 8;
  const int lanes = L.block / cgs > 0 ? L.block / cgs : 1;
  long long target = ((long long)num_sms() * 8 + NB - 1) / NB;  // pixel chunks per sample
  if (target < 1) target = 1;
  int ppb = (int)((HW + target - 1) / target);
  if (ppb < lanes) ppb = lanes;
  L.ppb = ppb;
  L.grid = dim3((HW + ppb - 1) / ppb, NB);
  return L;
}

}  // namespace gp

using namespace gp;

extern "C" {

int gp_cbn_apply_act(const void* y, const void* y_comp, void* out, void* out_comp, int comp_fmt, int NB, int H, int W,
                     int C, const float* mean, const float* rstd, const float* emb, const long long* labels, int act,
                     int upsample, void* stream) {
  GP_REQUIRE(y && out && mean && rstd && NB > 0 && C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0,
             "gp_cbn_apply_act: bad arguments (C/8 must divide 256)");
  GP_REQUIRE(comp_fmt >= GP_COMP_NONE && comp_fmt <= GP_COMP_F16, "gp_cbn_apply_act: unknown companion format %d", comp_fmt);
  GP_REQUIRE(emb == nullptr || labels != nullptr, "gp_cbn_apply_act: labels required with an embedding table");
  const CbnLaunch L = cbn_launch(NB, H * W, C);
  gp::launch_pdl(cbn_apply_kernel, L.grid, L.block, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(y), y_comp,
                                                              static_cast<__nv_bfloat16*>(out), out_comp, comp_fmt, H, W, C,
                                                              mean, rstd, emb, labels, act, upsample, L.ppb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

// part: fp32 [NB][2][C] scratch; S: fp32 [2][C] out; demb: fp32 [ncls][2C] or NULL (both zeroed by this call as needed)
int gp_cbn_bwd_reduce(const void* da, const void* y, const void* y_comp, int comp_fmt, int NB, int H, int W, int C,
                      const float* mean, const float* rstd, const float* emb, const long long* labels, int act,
                      int upsample, float* part, float* S, float* demb, int n_classes, void* stream) {
  GP_REQUIRE(comp_fmt >= GP_COMP_NONE && comp_fmt <= GP_COMP_F16, "gp_cbn_bwd_reduce: unknown companion format %d", comp_fmt);
  GP_REQUIRE(da && y && part && S && NB > 0 && C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0,
             "gp_cbn_bwd_reduce: bad arguments");
  cudaStream_t st = as_stream(stream);
  GP_CHECK_CUDA(cudaMemsetAsync(part, 0, sizeof(float) * NB * 2 * C, st));
  if (demb != nullptr) GP_CHECK_CUDA(cudaMemsetAsync(demb, 0, sizeof(float) * n_classes * 2 * C, st));
  const CbnLaunch L = cbn_launch(NB, H * W, C);
  gp::launch_pdl(cbn_bwd_reduce_kernel, L.grid, L.block, 2 * C * sizeof(float), st, static_cast<const __nv_bfloat16*>(da), static_cast<const __nv_bfloat16*>(y), y_comp, comp_fmt, H, W, C, mean, rstd,
      emb, labels, act, upsample, part, L.ppb);
  GP_CHECK_LAUNCH();
  GP_CHECK_CUDA(cudaMemsetAsync(S, 0, sizeof(float) * 2 * C, st));
  gp::launch_pdl(cbn_bwd_finalize_kernel, dim3((C + 127) / 128, sample_slices(NB)), 128, 0, st, part, NB, C, emb, labels, S, demb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_cbn_bwd_apply(const void* da, const void* y, const void* y_comp, int comp_fmt, void* dy, int NB, int H, int W,
                     int C, const float* mean, const float* rstd, const float* emb, const long long* labels,
                     const float* S, double count, int act, int upsample, void* stream) {
  GP_REQUIRE(comp_fmt >= GP_COMP_NONE && comp_fmt <= GP_COMP_F16, "gp_cbn_bwd_apply: unknown companion format %d", comp_fmt);
  GP_REQUIRE(da && y && dy && S && NB > 0 && C % 8 == 0 && count > 0, "gp_cbn_bwd_apply: bad arguments");
  const CbnLaunch L = cbn_launch(NB, H * W, C);
  gp::launch_pdl(cbn_bwd_apply_kernel, L.grid, L.block, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(da), static_cast<const __nv_bfloat16*>(y), y_comp, comp_fmt,
      static_cast<__nv_bfloat16*>(dy), H, W, C, mean, rstd, emb, labels, S, (float)(1.0 / count), act, upsample, L.ppb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_upsample2x(const void* in, void* out, int NB, int H, int W, int C, float scale, void* stream) {
  GP_REQUIRE(in && out && NB > 0 && C % 8 == 0, "gp_upsample2x: bad arguments");
  gp::launch_pdl(upsample2x_kernel, grid1((long long)NB * H * W * (C / 8)), 256, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), NB, H, W, C, scale);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_pool2x(const void* in, const void* in_comp, void* out, void* out_comp, int comp_fmt, int NB, int Hout, int Wout,
              int C, float scale, void* stream) {
  GP_REQUIRE(in && out && NB > 0 && C % 8 == 0, "gp_pool2x: bad arguments");
  GP_REQUIRE(comp_fmt >= GP_COMP_NONE && comp_fmt <= GP_COMP_F16, "gp_pool2x: unknown companion format %d", comp_fmt);
  gp::launch_pdl(pool2x_kernel, grid1((long long)NB * Hout * Wout * (C / 8)), 256, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(in), in_comp, static_cast<__nv_bfloat16*>(out), out_comp, comp_fmt, NB, Hout, Wout,
      C, scale);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_act_fwd(const void* in, const void* in_comp, void* out, void* out_comp, int comp_fmt, long long n, int act,
               void* stream) {
  GP_REQUIRE(in && out && n > 0 && n % 8 == 0, "gp_act_fwd: bad arguments");
  GP_REQUIRE(comp_fmt >= GP_COMP_NONE && comp_fmt <= GP_COMP_F16, "gp_act_fwd: unknown companion format %d", comp_fmt);
  gp::launch_pdl(act_fwd_kernel, grid1(n / 8), 256, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(in), in_comp,
                                                              static_cast<__nv_bfloat16*>(out), out_comp, comp_fmt, n / 8,
                                                              act);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_im2col_k3s1(const float* img, void* col, void* col_comp, int comp_fmt, int NB, int ch, int H, int W, void* stream) {
  GP_REQUIRE(img && col && NB > 0 && ch > 0 && ch * 9 <= 32, "gp_im2col_k3s1: bad arguments");
  GP_REQUIRE(comp_fmt >= GP_COMP_NONE && comp_fmt <= GP_COMP_F16, "gp_im2col_k3s1: unknown companion format %d", comp_fmt);
  gp::launch_pdl(im2col_k3s1_kernel, grid1((long long)NB * H * W * 4), 256, 0, as_stream(stream), img, static_cast<__nv_bfloat16*>(col), col_comp, comp_fmt, NB, ch, H, W);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_col2im_k3s1(const void* col, float* img, int NB, int ch, int H, int W, void* stream) {
  GP_REQUIRE(img && col && NB > 0 && ch > 0 && ch * 9 <= 32, "gp_col2im_k3s1: bad arguments");
  gp::launch_pdl(col2im_k3s1_kernel, grid1((long long)NB * ch * H * W), 256, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(col), img, NB, ch, H, W);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_nhwc8_to_image(const void* in, int in_f32, float* img, int NB, int ch, int HW, int tanh_act, void* stream) {
  GP_REQUIRE(in && img && NB > 0 && ch > 0 && ch <= 8, "gp_nhwc8_to_image: bad arguments");
  gp::launch_pdl(nhwc8_to_image_kernel, grid1((long long)NB * ch * HW), 256, 0, as_stream(stream), in, in_f32, img, NB, ch, HW,
                                                                                       tanh_act);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_image_to_nhwc8_grad(const float* dout, const float* out, void* dy, int NB, int ch, int HW, int tanh_act,
                           void* stream) {
  GP_REQUIRE(dout && dy && NB > 0 && ch > 0 && ch <= 8 && (!tanh_act || out), "gp_image_to_nhwc8_grad: bad arguments");
  gp::launch_pdl(image_to_nhwc8_grad_kernel, grid1((long long)NB * HW), 256, 0, as_stream(stream), dout, out, static_cast<__nv_bfloat16*>(dy), NB, ch, HW, tanh_act);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_relu_sumpool(const void* a, const void* a_comp, int comp_fmt, float* h, int NB, int HW, int C, void* stream) {
  GP_REQUIRE(a && h && NB > 0 && HW > 0 && C > 0 && C % 8 == 0, "gp_relu_sumpool: bad arguments (C %% 8 == 0)");
  GP_REQUIRE(comp_fmt >= GP_COMP_NONE && comp_fmt <= GP_COMP_F16, "gp_relu_sumpool: unknown companion format %d", comp_fmt);
  dim3 grid((C / 8 + 63) / 64, NB);
  gp::launch_pdl(relu_sumpool_kernel, grid, 64, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(a), a_comp, comp_fmt, h, HW, C);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_relu_sumpool_bwd(const float* dh, const void* a, void* da, int NB, int HW, int C, void* stream) {
  GP_REQUIRE(dh && a && da && NB > 0, "gp_relu_sumpool_bwd: bad arguments");
  const long long total = (long long)NB * HW * C;
  gp::launch_pdl(relu_sumpool_bwd_kernel, grid1(total), 256, 0, as_stream(stream), dh, static_cast<const __nv_bfloat16*>(a),
                                                                      static_cast<__nv_bfloat16*>(da), total, HW, C);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_proj_head_fwd(const float* h, const float* w, const float* b, const float* E, const long long* labels, float* out,
                     int NB, int C, void* stream) {
  GP_REQUIRE(h && w && out && NB > 0 && C > 0 && (E == nullptr || labels != nullptr), "gp_proj_head_fwd: bad arguments");
  gp::launch_pdl(proj_head_fwd_kernel, (NB + 7) / 8, 256, 0, as_stream(stream), h, w, b, E, labels, out, NB, C);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_proj_head_bwd(const float* dout, const float* h, const float* w, const float* E, const long long* labels,
                     float* dh, float* dw, float* db, float* dE, int NB, int C, int n_classes, void* stream) {
  GP_REQUIRE(dout && h && w && dh && dw && NB > 0 && C > 0, "gp_proj_head_bwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (dE != nullptr) GP_CHECK_CUDA(cudaMemsetAsync(dE, 0, sizeof(float) * n_classes * C, st));
  GP_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * C, st));
  if (db != nullptr) GP_CHECK_CUDA(cudaMemsetAsync(db, 0, sizeof(float), st));
  gp::launch_pdl(proj_head_bwd_kernel, dim3((C + 127) / 128, sample_slices(NB)), 128, 0, st, dout, h, w, E, labels, dh, dw, db, dE, NB, C);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

}  // extern "C"
