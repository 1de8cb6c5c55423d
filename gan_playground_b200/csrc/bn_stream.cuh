// Streaming (HBM-bound) BatchNorm kernels shared by the bf16 and bf16x3 precision modes.
// Activations are NHWC viewed as [P rows][C channels]; TY = storage type of the pre-BN conv output y
// (__nv_bfloat16 in "bf16" mode, float in "bf16x3" mode). Reference semantics: torch aten::native_batch_norm
// (+ backward) as called by nn.BatchNorm2d in models/dcgan.py:37,108 with the following in-place ReLU / LeakyReLU.
//
// Thread layout: one thread = one group of 8 channels (per-channel parameters live in registers) x one row lane;
// a block covers (C/8 groups) x (256 / (C/8) row lanes) and strides over its row range keeping kRowsInFlight rows
// (16- or 32-byte loads each) in flight per thread. Column reductions go through shared-memory accumulators:
// one global atomic per channel per block.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

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

// ---- 8-element vectors in registers: raw (as loaded) and unpacked fp32
template <typename T>
struct Vec8;
template <>
struct Vec8<__nv_bfloat16> {
  uint4 raw;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void zero() { raw = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
};
template <>
struct Vec8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void zero() { a = b = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
  }
};

// store 8 fp32 values as bf16 (hi) and optionally the rounding residual as a second bf16 vector (lo)
__device__ __forceinline__ void store8_bf16(__nv_bfloat16* hi, __nv_bfloat16* lo, const float (&f)[8]) {
  uint4 rh, rl;
  __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&rh);
  __nv_bfloat162* pl = reinterpret_cast<__nv_bfloat162*>(&rl);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    ph[i] = h2;
    if (lo != nullptr) {
      const float2 hf = __bfloat1622float2(h2);
      pl[i] = __floats2bfloat162_rn(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
    }
  }
  *reinterpret_cast<uint4*>(hi) = rh;
  if (lo != nullptr) *reinterpret_cast<uint4*>(lo) = rl;
}

// store 8 fp32 values as bf16 (backward operand) and as fp16 (operand of the next forward GEMM in the "fp16" mode)
__device__ __forceinline__ void store8_bf16_f16(__nv_bfloat16* b, __half* h, const float (&f)[8]) {
  uint4 rb, rh;
  __nv_bfloat162* pb = reinterpret_cast<__nv_bfloat162*>(&rb);
  __half2* ph = reinterpret_cast<__half2*>(&rh);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    pb[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    ph[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  }
  *reinterpret_cast<uint4*>(b) = rb;
  *reinterpret_cast<uint4*>(h) = rh;
}

constexpr int kRowsInFlight = 4;

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

struct RowRange {
  long long r0, r1;
};
__device__ __forceinline__ RowRange row_range(long long P, int rows_per_block) {
  RowRange R;
  R.r0 = (long long)blockIdx.x * rows_per_block;
  R.r1 = R.r0 + rows_per_block;
  if (R.r1 > P) R.r1 = P;
  return R;
}

// Per-channel sum and sum of squares over P rows (SQ = false: sums only -> bias gradients).
template <typename TY, bool SQ>
__global__ void col_stats_kernel(const TY* __restrict__ x, long long P, int C, float* __restrict__ sum,
                                 float* __restrict__ sumsq, int rows_per_block) {
  gp::pdl_sync();
  const ColLayout L = col_layout(C);
  constexpr int NQ = SQ ? 2 : 1;
  float acc[NQ][8];
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[q][i] = 0.f;
  if (L.active) {
    const RowRange R = row_range(P, rows_per_block);
    const TY* xp = x + L.g * 8;
    for (long long r = R.r0 + L.rl; r < R.r1; r += (long long)kRowsInFlight * L.lanes) {
      Vec8<TY> v[kRowsInFlight];
#pragma unroll
      for (int u = 0; u < kRowsInFlight; ++u) {
        const long long ru = r + (long long)u * L.lanes;
        if (ru < R.r1) v[u].load(xp + ru * C);
        else v[u].zero();
      }
#pragma unroll
      for (int u = 0; u < kRowsInFlight; ++u) {
        float f[8];
        v[u].unpack(f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[0][i] += f[i];
          if (SQ) acc[NQ - 1][i] += f[i] * f[i];
        }
      }
    }
  }
  if constexpr (SQ) {
    float* const outs[2] = {sum, sumsq};
    col_flush<2>(L, C, acc, outs);
  } else {
    float* const outs[1] = {sum};
    col_flush<1>(L, C, acc, outs);
  }
}

// out (bf16 hi [+ lo]) = act(y * scale[c] + shift[c])
// LOH: out_lo is an fp16 copy of the value instead of the bf16 rounding residual
template <typename TY, bool LOH = false>
__global__ void bn_apply_kernel(const TY* __restrict__ y, __nv_bfloat16* __restrict__ out_hi,
                                __nv_bfloat16* __restrict__ out_lo, long long P, int C, const float* __restrict__ scale,
                                const float* __restrict__ shift, int act, int rows_per_block) {
  gp::pdl_sync();
  const ColLayout L = col_layout(C);
  if (!L.active) return;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[L.g * 8 + j];
    sh[j] = shift[L.g * 8 + j];
  }
  const RowRange R = row_range(P, rows_per_block);
  const TY* yp = y + L.g * 8;
  for (long long r = R.r0 + L.rl; r < R.r1; r += (long long)kRowsInFlight * L.lanes) {
    Vec8<TY> v[kRowsInFlight];
#pragma unroll
    for (int u = 0; u < kRowsInFlight; ++u) {
      const long long ru = r + (long long)u * L.lanes;
      if (ru < R.r1) v[u].load(yp + ru * C);
    }
#pragma unroll
    for (int u = 0; u < kRowsInFlight; ++u) {
      const long long ru = r + (long long)u * L.lanes;
      if (ru < R.r1) {
        float f[8];
        v[u].unpack(f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = act_fwd(f[j] * sc[j] + sh[j], act);
        const long long off = ru * C + L.g * 8;
        if constexpr (LOH)
          store8_bf16_f16(out_hi + off, reinterpret_cast<__half*>(out_lo) + off, f);
        else
          store8_bf16(out_hi + off, out_lo ? out_lo + off : nullptr, f);
      }
    }
  }
}

// Backward reduction: sum_dz[c] = sum dz, sum_dzx[c] = sum dz * xhat, with z = y*scale+shift, dz = da*act'(z),
// xhat = (y - mean) * rstd.
template <typename TY>
__global__ void bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ da, const TY* __restrict__ y, long long P, int C,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     const float* __restrict__ mean, const float* __restrict__ rstd, int act,
                                     float* __restrict__ sum_dz, float* __restrict__ sum_dzx, int rows_per_block) {
  gp::pdl_sync();
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
    const RowRange R = row_range(P, rows_per_block);
    const TY* yp = y + L.g * 8;
    const __nv_bfloat16* dp = da + L.g * 8;
    for (long long r = R.r0 + L.rl; r < R.r1; r += (long long)kRowsInFlight * L.lanes) {
      Vec8<TY> vy[kRowsInFlight];
      Vec8<__nv_bfloat16> vd[kRowsInFlight];
#pragma unroll
      for (int u = 0; u < kRowsInFlight; ++u) {
        const long long ru = r + (long long)u * L.lanes;
        if (ru < R.r1) {
          vy[u].load(yp + ru * C);
          vd[u].load(dp + ru * C);
        } else {
          vy[u].zero();
          vd[u].zero();  // da = 0 contributes nothing
        }
      }
#pragma unroll
      for (int u = 0; u < kRowsInFlight; ++u) {
        float fy[8], fd[8];
        vy[u].unpack(fy);
        vd[u].unpack(fd);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float dz = fd[i] * act_grad(fy[i] * sc[i] + sh[i], act);
          acc[0][i] += dz;
          acc[1][i] += dz * (fy[i] - mu[i]) * rs[i];
        }
      }
    }
  }
  float* const outs[2] = {sum_dz, sum_dzx};
  col_flush<2>(L, C, acc, outs);
}

// dy = gamma*rstd * (dz - sum_dz/M - xhat * sum_dzx/M)
template <typename TY>
__global__ void bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ da, const TY* __restrict__ y,
                                    __nv_bfloat16* __restrict__ dy, long long P, int C, const float* __restrict__ scale,
                                    const float* __restrict__ shift, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, const float* __restrict__ sum_dz,
                                    const float* __restrict__ sum_dzx, float inv_count, int act, int rows_per_block,
                                    float* __restrict__ acc_dbeta, float* __restrict__ acc_dgamma, float acc_scale) {
  gp::pdl_sync();
  // affine-parameter gradients delivered straight into the parameters' .grad buffers (dbeta = sum dz, dgamma =
  // sum dz * xhat, times 1 / world under data parallelism): one block does it, no separate accumulation launches
  if (acc_dbeta != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      acc_dbeta[c] += sum_dz[c] * acc_scale;
      acc_dgamma[c] += sum_dzx[c] * acc_scale;
    }
  }
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
  const RowRange R = row_range(P, rows_per_block);
  const TY* yp = y + L.g * 8;
  const __nv_bfloat16* dp = da + L.g * 8;
  for (long long r = R.r0 + L.rl; r < R.r1; r += (long long)kRowsInFlight * L.lanes) {
    Vec8<TY> vy[kRowsInFlight];
    Vec8<__nv_bfloat16> vd[kRowsInFlight];
#pragma unroll
    for (int u = 0; u < kRowsInFlight; ++u) {
      const long long ru = r + (long long)u * L.lanes;
      if (ru < R.r1) {
        vy[u].load(yp + ru * C);
        vd[u].load(dp + ru * C);
      }
    }
#pragma unroll
    for (int u = 0; u < kRowsInFlight; ++u) {
      const long long ru = r + (long long)u * L.lanes;
      if (ru < R.r1) {
        float fy[8], fd[8], o[8];
        vy[u].unpack(fy);
        vd[u].unpack(fd);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dz = fd[j] * act_grad(fy[j] * sc[j] + sh[j], act);
          o[j] = sc[j] * dz - k0[j] - (fy[j] - mu[j]) * k1[j];
        }
        store8_bf16(dy + ru * C + L.g * 8, nullptr, o);
      }
    }
  }
}

// launch geometry of the row-strided kernels: ~8 blocks per SM, 256 threads (128 x grid.y tiles when C/8 does not
// divide 256)
struct ColLaunch {
  dim3 grid;
  int block;
  size_t smem;
  int rpb;
};
static inline ColLaunch col_launch(long long P, int C, int nq, int blocks_per_sm = 8) {
  ColLaunch L;
  // blocks_per_sm: 8 for the light kernels (statistics, column sums); the BatchNorm backward reduction keeps ~128
  // registers per thread, so 2 resident blocks per SM is all it gets — one wave of them, with 4x fewer per-block
  // flushes (shared-memory + global atomics per channel) than a finer split
  const long long blocks = (long long)num_sms() * blocks_per_sm;
  long long rpb = (P + blocks - 1) / blocks;
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
    if (gy < 2) gy = 2;  // gridDim.y > 1 selects the tiled layout inside the kernels
    L.grid = dim3(gx, gy);
    L.smem = (size_t)nq * 128 * 8 * sizeof(float);
  }
  return L;
}

}  // namespace gp
