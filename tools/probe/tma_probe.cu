// Probe: which fp32 image boxes does cp.async.bulk.tensor.3d accept?  usage: tma_probe <box_w> <start_w> <rows> <start_h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int nfloats, int c0, int c1, int c2, unsigned bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar), dst = (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes));
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(&map)), "r"(bar_a), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  __syncthreads();
  uint32_t ok = 0;
  long long spins = 0;
  while (!ok && spins < 20000000) {
    asm volatile("{\n .reg .pred P1;\n mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\n selp.b32 %0, 1, 0, P1;\n}" : "=r"(ok) : "r"(bar_a));
    ++spins;
  }
  if (!ok) { if (threadIdx.x == 0) out[0] = -12345.f; return; }
  for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

int main(int argc, char** argv) {
  int box_w = atoi(argv[1]), start_w = atoi(argv[2]), rows = atoi(argv[3]), start_h = atoi(argv[4]);
  const int W = 64, H = 64, CN = 12;
  std::vector<float> h(W * H * CN);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1000) + 1.f;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int nfl = box_w * rows * 3;
  cudaMalloc(&o, nfl * 4);
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  auto fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  CUtensorMap m;
  cuuint64_t dims[3] = {W, H, CN};
  cuuint64_t strides[2] = {W * 4, (cuuint64_t)W * H * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)rows, 3};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box_w %d start_w %d rows %d start_h %d: encode %d; ", box_w, start_w, rows, start_h, (int)r);
  if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  probe<<<1, 128, 40000>>>(m, o, nfl, start_w, start_h, 3, (unsigned)(nfl * 4));
  cudaError_t e = cudaDeviceSynchronize();
  printf("run: %s; ", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<float> got(nfl);
    cudaMemcpy(got.data(), o, nfl * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < 3; ++c)
      for (int rr = 0; rr < rows; ++rr)
        for (int x = 0; x < box_w; ++x) {
          int gx = start_w + x, gy = start_h + rr;
          float want = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0.f : h[((size_t)(3 + c) * H + gy) * W + gx];
          if (got[(c * rows + rr) * box_w + x] != want) ++bad;
        }
    printf("first %.1f mismatches %d", got[0], bad);
  }
  printf("\n");
  return 0;
}
