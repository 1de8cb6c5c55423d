// Activation storage conventions of the forward precision modes, as 8-channel register vectors.
//
// Every NHWC activation has a bf16 tensor `p` — the one autograd sees and every backward kernel / GEMM reads — and,
// depending on the mode of the forward pass that produced it, a COMPANION tensor of the same shape:
//   GP_COMP_NONE  "bf16" mode : value = p
//   GP_COMP_LO    "bf16x3"    : companion = bf16(value - p), value = p + comp (~16 significant bits)
//   GP_COMP_F16   "fp16"      : companion = fp16(value), the operand of the next single-MMA fp16 forward GEMM (11 bits)
// Forward element-wise kernels read the most precise view (load8c) and write both tensors (store8c); a NULL companion
// pointer means "produced in another mode": read p alone / write p alone.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../../include/gpb200.h"

namespace gp {

__device__ __forceinline__ void unpack8_bf16(const uint4& raw, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void unpack8_f16(const uint4& raw, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// 8 consecutive channels starting at element offset `off` (16-byte aligned)
__device__ __forceinline__ void load8c(const __nv_bfloat16* __restrict__ p, const void* __restrict__ comp, int fmt,
                                       long long off, float (&f)[8]) {
  if (comp != nullptr && fmt == GP_COMP_F16) {
    unpack8_f16(*reinterpret_cast<const uint4*>(static_cast<const __half*>(comp) + off), f);
    return;
  }
  unpack8_bf16(*reinterpret_cast<const uint4*>(p + off), f);
  if (comp != nullptr && fmt == GP_COMP_LO) {
    float l[8];
    unpack8_bf16(*reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(comp) + off), l);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] += l[i];
  }
}

// the pair of 16-byte vectors store8c writes (so that a kernel can write the same value to several places)
struct Packed8c {
  uint4 hi, comp;
};
__device__ __forceinline__ Packed8c pack8c(int fmt, const float (&f)[8]) {
  Packed8c o;
  __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&o.hi);
#pragma unroll
  for (int i = 0; i < 4; ++i) ph[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  if (fmt == GP_COMP_F16) {
    __half2* pc = reinterpret_cast<__half2*>(&o.comp);
#pragma unroll
    for (int i = 0; i < 4; ++i) pc[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  } else if (fmt == GP_COMP_LO) {
    __nv_bfloat162* pc = reinterpret_cast<__nv_bfloat162*>(&o.comp);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 hf = __bfloat1622float2(ph[i]);
      pc[i] = __floats2bfloat162_rn(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
    }
  } else {
    o.comp = make_uint4(0u, 0u, 0u, 0u);
  }
  return o;
}
__device__ __forceinline__ void put8c(__nv_bfloat16* __restrict__ p, void* __restrict__ comp, long long off,
                                      const Packed8c& v) {
  *reinterpret_cast<uint4*>(p + off) = v.hi;
  if (comp != nullptr) *reinterpret_cast<uint4*>(static_cast<uint16_t*>(comp) + off) = v.comp;
}
__device__ __forceinline__ void store8c(__nv_bfloat16* __restrict__ p, void* __restrict__ comp, int fmt, long long off,
                                        const float (&f)[8]) {
  put8c(p, comp, off, pack8c(comp != nullptr ? fmt : GP_COMP_NONE, f));
}

}  // namespace gp
