// BlurPool2d(filt_size=3, pad_type='reflect') of the reference's models/ops.py:7-47, as used by models/dcgan_blur.py
// (:41 stride 1 after every generator conv, :116 stride 2 after every discriminator block but the last): reflection
// padding 1 then a depth-wise 3x3 convolution with the fixed kernel outer([1,2,1],[1,2,1])/16. NHWC bf16, HBM-bound:
// one thread per output pixel x 8-channel group (16-byte accesses), fp32 accumulation.
#include <cuda_bf16.h>

#include "act_io.cuh"
#include "common.h"

namespace gp {

__device__ __forceinline__ void blur_unpack8(const uint4& raw, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 blur_pack8(const float (&f)[8]) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return raw;
}
// index into the unpadded axis of padded position q (pad 1, reflect): -1 -> 1, n -> n-2
__device__ __forceinline__ int reflect1(int q, int n) {
  int r = q - 1;
  if (r < 0) r = -r;
  if (r >= n) r = 2 * n - 2 - r;
  return r;
}

// out[n, oh, ow, :] = sum_{kh,kw} w[kh] w[kw] / 16 * in[n, reflect(oh*s + kh), reflect(ow*s + kw), :],  w = (1, 2, 1)
// in_comp / out_comp / fmt: companion tensors of the forward precision mode (act_io.cuh)
__global__ void blur3x3_fwd_kernel(const __nv_bfloat16* __restrict__ in, const void* __restrict__ in_comp,
                                   __nv_bfloat16* __restrict__ out, void* __restrict__ out_comp, int fmt, int NB, int H,
                                   int W, int C, int s, int Ho, int Wo) {
  gp::pdl_sync();
  const int cgs = C / 8;
  const long long total = (long long)NB * Ho * Wo * cgs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cgs);
    const long long p = i / cgs;
    const int ow = (int)(p % Wo), oh = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = reflect1(oh * s + kh, H);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = reflect1(ow * s + kw, W);
        const float wgt = (kh == 1 ? 2.f : 1.f) * (kw == 1 ? 2.f : 1.f) * (1.f / 16.f);
        float f[8];
        load8c(in, in_comp, fmt, (((long long)n * H + ih) * W + iw) * C + g * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += wgt * f[j];
      }
    }
    store8c(out, out_comp, fmt, p * C + g * 8, acc);
  }
}

// Adjoint: din[n, ih, iw, :] = sum over padded positions q that reflect onto ih (q = ih+1, plus 0 when ih == 1, plus
// H+1 when ih == H-2), taps kh with (q - kh) divisible by s, and the same along the width, of w[kh] w[kw]/16 * dout.
__global__ void blur3x3_bwd_kernel(const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ din, int NB, int H,
                                   int W, int C, int s, int Ho, int Wo) {
  gp::pdl_sync();
  const int cgs = C / 8;
  const long long total = (long long)NB * H * W * cgs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cgs);
    const long long p = i / cgs;
    const int iw = (int)(p % W), ih = (int)((p / W) % H), n = (int)(p / ((long long)W * H));
    int qh[3], qw[3], nh = 0, nw = 0;
    qh[nh++] = ih + 1;
    if (ih == 1) qh[nh++] = 0;
    if (ih == H - 2) qh[nh++] = H + 1;
    qw[nw++] = iw + 1;
    if (iw == 1) qw[nw++] = 0;
    if (iw == W - 2) qw[nw++] = W + 1;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int a = 0; a < nh; ++a) {
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int th = qh[a] - kh;
        if (th < 0 || th % s != 0 || th / s >= Ho) continue;
        for (int b = 0; b < nw; ++b) {
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int tw = qw[b] - kw;
            if (tw < 0 || tw % s != 0 || tw / s >= Wo) continue;
            const float wgt = (kh == 1 ? 2.f : 1.f) * (kw == 1 ? 2.f : 1.f) * (1.f / 16.f);
            float f[8];
            blur_unpack8(*reinterpret_cast<const uint4*>(dout + (((long long)n * Ho + th / s) * Wo + tw / s) * C + g * 8), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += wgt * f[j];
          }
        }
      }
    }
    *reinterpret_cast<uint4*>(din + p * C + g * 8) = blur_pack8(acc);
  }
}

static inline int blur_grid(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace gp

using namespace gp;

extern "C" {

int gp_blur3x3_fwd(const void* in, const void* in_comp, void* out, void* out_comp, int comp_fmt, int NB, int H, int W,
                   int C, int stride, void* stream) {
  GP_REQUIRE(in && out && NB > 0 && H >= 2 && W >= 2 && C > 0 && C % 8 == 0 && (stride == 1 || stride == 2),
             "gp_blur3x3_fwd: bad arguments (C %% 8 == 0, H, W >= 2, stride 1 or 2)");
  GP_REQUIRE(comp_fmt >= GP_COMP_NONE && comp_fmt <= GP_COMP_F16, "gp_blur3x3_fwd: unknown companion format %d", comp_fmt);
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  gp::launch_pdl(blur3x3_fwd_kernel, blur_grid((long long)NB * Ho * Wo * (C / 8)), 256, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(in), in_comp, static_cast<__nv_bfloat16*>(out), out_comp, comp_fmt, NB, H, W, C,
      stride, Ho, Wo);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_blur3x3_bwd(const void* dout, void* din, int NB, int H, int W, int C, int stride, void* stream) {
  GP_REQUIRE(dout && din && NB > 0 && H >= 2 && W >= 2 && C > 0 && C % 8 == 0 && (stride == 1 || stride == 2),
             "gp_blur3x3_bwd: bad arguments (C %% 8 == 0, H, W >= 2, stride 1 or 2)");
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  gp::launch_pdl(blur3x3_bwd_kernel, blur_grid((long long)NB * H * W * (C / 8)), 256, 0, as_stream(stream), static_cast<const __nv_bfloat16*>(dout), static_cast<__nv_bfloat16*>(din), NB, H, W, C, stride, Ho, Wo);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

}  // extern "C"
