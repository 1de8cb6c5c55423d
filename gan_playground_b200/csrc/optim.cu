// Adam over flat fp32 buffers: one launch per network instead of one multi-tensor pass per state tensor.
// Semantics of torch.optim.Adam as the reference configures it (main_dcgan.py:55-56, main_sngan.py:55-56: no weight
// decay, no amsgrad, eps 1e-8), torch:optim/adam.py `_single_tensor_adam` capturable branch:
//   m <- m + (g - m) * (1 - beta1);  v <- beta2 * v + (1 - beta2) * g * g
//   p <- p - (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// `step` is a device scalar holding the number of steps taken BEFORE this call (so the launch can sit in a CUDA graph);
// the caller increments it afterwards. grad_scale folds the 1/world of data-parallel gradient averaging into the read.
#include "common.h"

namespace gp {

__global__ void adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n4, float lr, float beta1, float beta2, float w1,
                                 float w2, float eps, const float* __restrict__ step, float grad_scale) {
  gp::pdl_sync();
  // w1 = 1 - beta1, w2 = 1 - beta2 are rounded from the host's double arithmetic, as torch does (python floats)
  const float t = __ldg(step) + 1.f;
  const float bc1 = 1.f - powf(beta1, t);
  const float bc2 = 1.f - powf(beta2, t);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    float* pa = reinterpret_cast<float*>(&pp);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
    const float* ga = reinterpret_cast<const float*>(&gg);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = ga[k] * grad_scale;
      ma[k] = ma[k] + (gk - ma[k]) * w1;
      va[k] = beta2 * va[k] + w2 * gk * gk;
      const float denom = sqrtf(va[k]) * inv_sqrt_bc2 + eps;
      pa[k] -= step_size * (ma[k] / denom);
    }
    p4[i] = pp;
    m4[i] = mm;
    v4[i] = vv;
  }
}

}  // namespace gp

extern "C" int gp_adam_flat(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1,
                            double beta2, double eps, const float* step, double grad_scale, void* stream) {
  GP_REQUIRE(p && g && m && v && step && n > 0 && n % 4 == 0, "gp_adam_flat: bad arguments (n %% 4 == 0 required)");
  GP_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
               reinterpret_cast<uintptr_t>(v)) & 15) == 0, "gp_adam_flat: buffers must be 16-byte aligned");
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)gp::num_sms() * 16;
  if (blocks > cap) blocks = cap;
  gp::launch_pdl(gp::adam_flat_kernel, (int)blocks, 256, 0, gp::as_stream(stream), p, g, m, v, n4, (float)lr, (float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, step,
      (float)grad_scale);
  GP_CHECK_LAUNCH();
  return GP_OK;
}
