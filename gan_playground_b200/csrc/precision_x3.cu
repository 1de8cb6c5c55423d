// Kernels of the "bf16x3" forward precision mode: forward GEMM operands are hi/lo bf16 pairs
// (x = hi + lo to ~16 mantissa bits) and pre-BatchNorm conv outputs are kept in fp32, so that the forward pass
// carries fp32-class activations while every tensor-core instruction stays bf16 (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo).
// These are the HBM-bound producers/consumers of those formats: the BatchNorm kernels of bn_stream.cuh instantiated
// for fp32 y, and the hi/lo staging of plain matrices. (Conv-weight / im2col / col2im / head variants share their
// kernels with the bf16 mode and live in elementwise.cu; the GEMM side is conv_gemm.cuh with extra K taps.)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "act_io.cuh"
#include "bn_stream.cuh"
#include "common.h"

namespace gp {
namespace x3 {

// dst[r*ld + k] = hi(src[map(r)*s_r + k*s_k]), dst[lo_off + r*ld + k] = lo(...); zero padding to [Rpad][width].
__global__ void split_matrix_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int R, int K,
                                    int Rpad, int ld, int width, long long s_r, long long s_k, int perm,
                                    long long lo_off) {
  gp::pdl_sync();
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
    const __nv_bfloat16 hi = __float2bfloat16(v);
    dst[(long long)r * ld + k] = hi;
    dst[lo_off + (long long)r * ld + k] = __float2bfloat16(v - __bfloat162float(hi));
  }
}

// out[r*ld_out + k] = fp16(hi[r*ld_in + k] + lo[r*ld_in + k]): the single-MMA fp16 operand from a hi/lo bf16 pair
// (hi + lo carries ~16 significant bits, so the double rounding is below fp16's own). 8 columns per thread.
__global__ void pair_to_f16_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                   long long ld_in, __half* __restrict__ out, long long ld_out, long long rows, int cols8) {
  gp::pdl_sync();
  const long long total = rows * cols8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols8;
    const int k = (int)(i % cols8) * 8;
    const uint4 vh = *reinterpret_cast<const uint4*>(hi + r * ld_in + k);
    const uint4 vl = *reinterpret_cast<const uint4*>(lo + r * ld_in + k);
    const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&vh);
    const __nv_bfloat162* pl = reinterpret_cast<const __nv_bfloat162*>(&vl);
    uint4 o;
    __half2* po = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = __bfloat1622float2(ph[j]), b = __bfloat1622float2(pl[j]);
      po[j] = __floats2half2_rn(a.x + b.x, a.y + b.y);
    }
    *reinterpret_cast<uint4*>(out + r * ld_out + k) = o;
  }
}

// BatchNorm statistics / apply on an activation stored with a companion tensor (act_io.cuh): the nodes that normalise
// an EXISTING activation instead of a conv's fp32 output (generator's b6 of models/sngan_projection.py:92, every
// BatchNorm of models/dcgan_blur.py, which follows a BlurPool). Same thread layout as bn_stream.cuh.
__global__ void col_stats_comp_kernel(const __nv_bfloat16* __restrict__ x, const void* __restrict__ x_comp, int fmt,
                                      long long P, int C, float* __restrict__ sum, float* __restrict__ sumsq,
                                      int rows_per_block) {
  gp::pdl_sync();
  const ColLayout L = col_layout(C);
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (L.active) {
    const RowRange R = row_range(P, rows_per_block);
    for (long long r = R.r0 + L.rl; r < R.r1; r += (long long)kRowsInFlight * L.lanes) {
      float f[kRowsInFlight][8];
#pragma unroll
      for (int u = 0; u < kRowsInFlight; ++u) {
        const long long ru = r + (long long)u * L.lanes;
        if (ru < R.r1) {
          load8c(x, x_comp, fmt, ru * C + L.g * 8, f[u]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[u][i] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < kRowsInFlight; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[0][i] += f[u][i];
          acc[1][i] += f[u][i] * f[u][i];
        }
    }
  }
  float* const outs[2] = {sum, sumsq};
  col_flush<2>(L, C, acc, outs);
}

__global__ void bn_apply_comp_kernel(const __nv_bfloat16* __restrict__ y, const void* __restrict__ y_comp,
                                     __nv_bfloat16* __restrict__ out, void* __restrict__ out_comp, int fmt, long long P,
                                     int C, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                     int rows_per_block) {
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
  for (long long r = R.r0 + L.rl; r < R.r1; r += (long long)kRowsInFlight * L.lanes) {
    float f[kRowsInFlight][8];
#pragma unroll
    for (int u = 0; u < kRowsInFlight; ++u) {
      const long long ru = r + (long long)u * L.lanes;
      if (ru < R.r1) load8c(y, y_comp, fmt, ru * C + L.g * 8, f[u]);
    }
#pragma unroll
    for (int u = 0; u < kRowsInFlight; ++u) {
      const long long ru = r + (long long)u * L.lanes;
      if (ru < R.r1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[u][j] = act_fwd(f[u][j] * sc[j] + sh[j], act);
        store8c(out, out_comp, fmt, ru * C + L.g * 8, f[u]);
      }
    }
  }
}

// BatchNorm backward with y read through its companion (the SAME value the forward normalised: the activation mask
// act'(y*scale+shift) and xhat must not be recomputed from the bf16 rounding of y, or near-zero pre-activations flip
// sides — measured on dcgan_blur: G-step gradient cosine 0.9984 with the bf16 view).
__global__ void bn_bwd_reduce_comp_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ y,
                                          const void* __restrict__ y_comp, int fmt, long long P, int C,
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
    for (long long r = R.r0 + L.rl; r < R.r1; r += (long long)kRowsInFlight * L.lanes) {
      float fy[kRowsInFlight][8], fd[kRowsInFlight][8];
#pragma unroll
      for (int u = 0; u < kRowsInFlight; ++u) {
        const long long ru = r + (long long)u * L.lanes;
        if (ru < R.r1) {
          load8c(y, y_comp, fmt, ru * C + L.g * 8, fy[u]);
          load8c(da, nullptr, GP_COMP_NONE, ru * C + L.g * 8, fd[u]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) fy[u][i] = fd[u][i] = 0.f;  // da = 0 contributes nothing
        }
      }
#pragma unroll
      for (int u = 0; u < kRowsInFlight; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float dz = fd[u][i] * act_grad(fy[u][i] * sc[i] + sh[i], act);
          acc[0][i] += dz;
          acc[1][i] += dz * (fy[u][i] - mu[i]) * rs[i];
        }
    }
  }
  float* const outs[2] = {sum_dz, sum_dzx};
  col_flush<2>(L, C, acc, outs);
}

__global__ void bn_bwd_apply_comp_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ y,
                                         const void* __restrict__ y_comp, int fmt, __nv_bfloat16* __restrict__ dy,
                                         long long P, int C, const float* __restrict__ scale,
                                         const float* __restrict__ shift, const float* __restrict__ mean,
                                         const float* __restrict__ rstd, const float* __restrict__ sum_dz,
                                         const float* __restrict__ sum_dzx, float inv_count, int act, int rows_per_block,
                                         float* __restrict__ acc_dbeta, float* __restrict__ acc_dgamma, float acc_scale) {
  gp::pdl_sync();
  if (acc_dbeta != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      acc_dbeta[c] += sum_dz[c] * acc_scale;
      acc_dgamma[c] += sum_dzx[c] * acc_scale;
    }
  }
  const ColLayout L = col_layout(C);
  if (!L.active) return;
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
  for (long long r = R.r0 + L.rl; r < R.r1; r += (long long)kRowsInFlight * L.lanes) {
    float fy[kRowsInFlight][8], fd[kRowsInFlight][8];
#pragma unroll
    for (int u = 0; u < kRowsInFlight; ++u) {
      const long long ru = r + (long long)u * L.lanes;
      if (ru < R.r1) {
        load8c(y, y_comp, fmt, ru * C + L.g * 8, fy[u]);
        load8c(da, nullptr, GP_COMP_NONE, ru * C + L.g * 8, fd[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kRowsInFlight; ++u) {
      const long long ru = r + (long long)u * L.lanes;
      if (ru < R.r1) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dz = fd[u][j] * act_grad(fy[u][j] * sc[j] + sh[j], act);
          o[j] = sc[j] * dz - k0[j] - (fy[u][j] - mu[j]) * k1[j];
        }
        store8_bf16(dy + ru * C + L.g * 8, nullptr, o);
      }
    }
  }
}

}  // namespace x3
}  // namespace gp

using namespace gp;

extern "C" {

int gp_split_matrix(const float* src, void* dst, int R, int K, int Rpad, int ld, int width, long long s_r, long long s_k,
                    int perm, long long lo_off, void* stream) {
  GP_REQUIRE(src && dst && R > 0 && K > 0 && Rpad >= R && width >= K && ld >= width && lo_off > 0, "gp_split_matrix: bad arguments");
  GP_REQUIRE(perm <= 1 || R % perm == 0, "gp_split_matrix: perm must divide R");
  long long g = ((long long)Rpad * width + 255) / 256;
  if (g > (long long)num_sms() * 16) g = (long long)num_sms() * 16;
  gp::launch_pdl(x3::split_matrix_kernel, (int)g, 256, 0, as_stream(stream), src, static_cast<__nv_bfloat16*>(dst), R, K, Rpad, ld,
                                                                 width, s_r, s_k, perm, lo_off);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_stats_f32(const float* y, long long P, int C, float* sum, float* sumsq, void* stream) {
  GP_REQUIRE(y && sum && sumsq && P > 0 && C > 0 && C % 8 == 0, "gp_bn_stats_f32: bad arguments");
  const ColLaunch L = col_launch(P, C, 2);
  gp::launch_pdl(col_stats_kernel<float, true>, L.grid, L.block, L.smem, as_stream(stream), y, P, C, sum, sumsq, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_apply_act_split(const float* y, void* out_hi, void* out_lo, long long P, int C, const float* scale,
                          const float* shift, int act, void* stream) {
  GP_REQUIRE(y && out_hi && scale && shift && P > 0 && C % 8 == 0, "gp_bn_apply_act_split: bad arguments");
  const ColLaunch L = col_launch(P, C, 0);
  gp::launch_pdl(bn_apply_kernel<float>, L.grid, L.block, 0, as_stream(stream), y, static_cast<__nv_bfloat16*>(out_hi),
                                                                    static_cast<__nv_bfloat16*>(out_lo), P, C, scale,
                                                                    shift, act, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_pair_to_f16(const void* hi, const void* lo, long long ld_in, void* out, long long ld_out, long long rows, int cols,
                   void* stream) {
  GP_REQUIRE(hi && lo && out && rows > 0 && cols > 0 && cols % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 &&
                 ld_in >= cols && ld_out >= cols,
             "gp_pair_to_f16: bad arguments (cols and pitches must be multiples of 8)");
  GP_REQUIRE(((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
             "gp_pair_to_f16: pointers must be 16-byte aligned");
  long long g = (rows * (cols / 8) + 255) / 256;
  if (g > (long long)num_sms() * 16) g = (long long)num_sms() * 16;
  gp::launch_pdl(x3::pair_to_f16_kernel, (int)g, 256, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(hi),
                                                                static_cast<const __nv_bfloat16*>(lo), ld_in,
                                                                static_cast<__half*>(out), ld_out, rows, cols / 8);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_apply_act_pair(const float* y, void* out_bf16, void* out_f16, long long P, int C, const float* scale,
                         const float* shift, int act, void* stream) {
  GP_REQUIRE(y && out_bf16 && out_f16 && scale && shift && P > 0 && C % 8 == 0, "gp_bn_apply_act_pair: bad arguments");
  const ColLaunch L = col_launch(P, C, 0);
  gp::launch_pdl(bn_apply_kernel<float, true>, L.grid, L.block, 0, as_stream(stream), y, static_cast<__nv_bfloat16*>(out_bf16),
                                                                          static_cast<__nv_bfloat16*>(out_f16), P, C, scale,
                                                                          shift, act, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_stats_comp(const void* x, const void* x_comp, int comp_fmt, long long P, int C, float* sum, float* sumsq,
                     void* stream) {
  GP_REQUIRE(x && sum && sumsq && P > 0 && C > 0 && C % 8 == 0, "gp_bn_stats_comp: bad arguments (C %% 8 == 0 required)");
  GP_REQUIRE(comp_fmt >= GP_COMP_NONE && comp_fmt <= GP_COMP_F16, "gp_bn_stats_comp: unknown companion format %d", comp_fmt);
  const ColLaunch L = col_launch(P, C, 2);
  gp::launch_pdl(x3::col_stats_comp_kernel, L.grid, L.block, L.smem, as_stream(stream), static_cast<const __nv_bfloat16*>(x), x_comp,
                                                                            comp_fmt, P, C, sum, sumsq, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_apply_act_comp(const void* y, const void* y_comp, void* out, void* out_comp, int comp_fmt, long long P, int C,
                         const float* scale, const float* shift, int act, void* stream) {
  GP_REQUIRE(y && out && scale && shift && P > 0 && C % 8 == 0, "gp_bn_apply_act_comp: bad arguments");
  GP_REQUIRE(comp_fmt >= GP_COMP_NONE && comp_fmt <= GP_COMP_F16, "gp_bn_apply_act_comp: unknown companion format %d", comp_fmt);
  const ColLaunch L = col_launch(P, C, 0);
  gp::launch_pdl(x3::bn_apply_comp_kernel, L.grid, L.block, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(y), y_comp,
                                                                      static_cast<__nv_bfloat16*>(out), out_comp, comp_fmt,
                                                                      P, C, scale, shift, act, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_bwd_reduce_comp(const void* da, const void* y, const void* y_comp, int comp_fmt, long long P, int C,
                          const float* scale, const float* shift, const float* mean, const float* rstd, int act,
                          float* sum_dz, float* sum_dzx, void* stream) {
  GP_REQUIRE(da && y && sum_dz && sum_dzx && P > 0 && C % 8 == 0, "gp_bn_bwd_reduce_comp: bad arguments");
  GP_REQUIRE(comp_fmt >= GP_COMP_NONE && comp_fmt <= GP_COMP_F16, "gp_bn_bwd_reduce_comp: unknown companion format %d", comp_fmt);
  const ColLaunch L = col_launch(P, C, 2, 2);
  gp::launch_pdl(x3::bn_bwd_reduce_comp_kernel, L.grid, L.block, L.smem, as_stream(stream), static_cast<const __nv_bfloat16*>(da), static_cast<const __nv_bfloat16*>(y), y_comp, comp_fmt, P, C, scale, shift, mean,
      rstd, act, sum_dz, sum_dzx, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_bwd_apply_comp(const void* da, const void* y, const void* y_comp, int comp_fmt, void* dy, long long P, int C,
                         const float* scale, const float* shift, const float* mean, const float* rstd,
                         const float* sum_dz, const float* sum_dzx, double count, int act, float* acc_dbeta,
                         float* acc_dgamma, float acc_scale, void* stream) {
  GP_REQUIRE(da && y && dy && P > 0 && C % 8 == 0 && count > 0, "gp_bn_bwd_apply_comp: bad arguments");
  GP_REQUIRE(comp_fmt >= GP_COMP_NONE && comp_fmt <= GP_COMP_F16, "gp_bn_bwd_apply_comp: unknown companion format %d", comp_fmt);
  GP_REQUIRE((acc_dbeta == nullptr) == (acc_dgamma == nullptr), "gp_bn_bwd_apply_comp: acc_dbeta and acc_dgamma go together");
  const ColLaunch L = col_launch(P, C, 0);
  gp::launch_pdl(x3::bn_bwd_apply_comp_kernel, L.grid, L.block, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(da), static_cast<const __nv_bfloat16*>(y), y_comp, comp_fmt,
      static_cast<__nv_bfloat16*>(dy), P, C, scale, shift, mean, rstd, sum_dz, sum_dzx, (float)(1.0 / count), act, L.rpb,
      acc_dbeta, acc_dgamma, acc_scale);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_bwd_reduce_f32(const void* da, const float* y, long long P, int C, const float* scale, const float* shift,
                         const float* mean, const float* rstd, int act, float* sum_dz, float* sum_dzx, void* stream) {
  GP_REQUIRE(da && y && sum_dz && sum_dzx && P > 0 && C % 8 == 0, "gp_bn_bwd_reduce_f32: bad arguments");
  const ColLaunch L = col_launch(P, C, 2, 2);
  gp::launch_pdl(bn_bwd_reduce_kernel<float>, L.grid, L.block, L.smem, as_stream(stream), static_cast<const __nv_bfloat16*>(da), y, P, C, scale, shift, mean, rstd, act, sum_dz, sum_dzx, L.rpb);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_bn_bwd_apply_f32(const void* da, const float* y, void* dy, long long P, int C, const float* scale,
                        const float* shift, const float* mean, const float* rstd, const float* sum_dz,
                        const float* sum_dzx, double count, int act, float* acc_dbeta, float* acc_dgamma, float acc_scale,
                    void* stream) {
  GP_REQUIRE(da && y && dy && P > 0 && C % 8 == 0 && count > 0, "gp_bn_bwd_apply_f32: bad arguments");
  GP_REQUIRE((acc_dbeta == nullptr) == (acc_dgamma == nullptr), "gp_bn_bwd_apply_f32: acc_dbeta and acc_dgamma go together");
  const ColLaunch L = col_launch(P, C, 0);
  gp::launch_pdl(bn_bwd_apply_kernel<float>, L.grid, L.block, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(da), y, static_cast<__nv_bfloat16*>(dy), P, C, scale, shift, mean, rstd, sum_dz,
      sum_dzx, (float)(1.0 / count), act, L.rpb, acc_dbeta, acc_dgamma, acc_scale);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

}  // extern "C"
