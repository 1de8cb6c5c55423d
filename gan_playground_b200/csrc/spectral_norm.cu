// Spectral normalisation (torch.nn.utils.spectral_norm, torch:nn/utils/spectral_norm.py:62-114) as GEMV kernels.
// The weight is used in place, in torch's layout [A][B][T] (fp32):
//   dim == 0 (Conv2d / Linear / Embedding): W_mat[r = a][c = b*T + t]
//   dim == 1 (ConvTranspose2d)            : W_mat[r = b][c = a*T + t]
// One power iteration:  v <- normalize(W^T u);  u <- normalize(W v);  sigma = u . (W v)  (= ||W v|| after the update).
// HBM-bound: W is read twice per forward (once per GEMV) plus once for the W / sigma staging.
#include "common.h"

#include <cstring>

namespace gp {

struct SnLayout {
  int A, B, T, dim;
  __device__ __host__ int rows() const { return dim == 0 ? A : B; }
  __device__ __host__ int cols() const { return dim == 0 ? B * T : A * T; }
  __device__ __forceinline__ long long addr(int r, int c) const {
    if (dim == 0) return (long long)r * (B * T) + c;
    return ((long long)(c / T) * B + r) * T + (c % T);
  }
};

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[32];
  for (int k = 16; k > 0; k >>= 1) v += __shfl_xor_sync(0xffffffffu, v, k);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32)
    for (int k = 16; k > 0; k >>= 1) t += __shfl_xor_sync(0xffffffffu, t, k);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  t = red[0];
  __syncthreads();
  return t;
}

// t1[c] += sum_{r in chunk} W[r][c] * u[r]      grid: (ceil(cols/256), row chunks)
__global__ void sn_gemv_t_kernel(const float* __restrict__ w, SnLayout L, const float* __restrict__ u,
                                 float* __restrict__ t1, int rows_per_block) {
  gp::pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= L.cols()) return;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(r0 + rows_per_block, L.rows());
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) acc += __ldg(w + L.addr(r, c)) * __ldg(u + r);
  atomicAdd(t1 + c, acc);
}

// t2[r] = sum_c W[r][c] * v[c]      grid: rows
__global__ void sn_gemv_kernel(const float* __restrict__ w, SnLayout L, const float* __restrict__ v,
                               float* __restrict__ t2) {
  gp::pdl_sync();
  const int r = blockIdx.x;
  float acc = 0.f;
  for (int c = threadIdx.x; c < L.cols(); c += blockDim.x) acc += __ldg(w + L.addr(r, c)) * __ldg(v + c);
  acc = block_sum(acc);
  if (threadIdx.x == 0) t2[r] = acc;
}

// out[i] = t[i] / max(||t||, eps); sigma (optional) = sum t^2 / max(||t||, eps)      single block
__global__ void sn_normalize_kernel(const float* __restrict__ t, int n, float eps, float* __restrict__ out,
                                    float* __restrict__ sigma) {
  gp::pdl_sync();
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += t[i] * t[i];
  acc = block_sum(acc);
  const float nrm = fmaxf(sqrtf(acc), eps);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = t[i] / nrm;
  if (sigma != nullptr && threadIdx.x == 0) *sigma = acc / nrm;
}

// eval mode: sigma = u . t2      single block
__global__ void sn_dot_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, float* __restrict__ out) {
  gp::pdl_sync();
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += a[i] * b[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) *out = acc;
}

// out = w / sigma
__global__ void sn_scale_kernel(const float* __restrict__ w, const float* __restrict__ sigma, float* __restrict__ out,
                                long long n) {
  gp::pdl_sync();
  const float inv = 1.f / __ldg(sigma);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = w[i] * inv;
}

// dot += sum g * w_sn
__global__ void sn_grad_dot_kernel(const float* __restrict__ g, const float* __restrict__ wsn, long long n,
                                   float* __restrict__ dot) {
  gp::pdl_sync();
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += g[i] * wsn[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(dot, acc);
}

// dW_orig[e] = (g[e] - dot * u[r(e)] * v[c(e)]) / sigma      (gradient through W / sigma with u, v constant)
__global__ void sn_grad_kernel(const float* __restrict__ g, SnLayout L, const float* __restrict__ u,
                               const float* __restrict__ v, const float* __restrict__ sigma,
                               const float* __restrict__ dot, float* __restrict__ out) {
  gp::pdl_sync();
  const long long n = (long long)L.A * L.B * L.T;
  const float inv = 1.f / __ldg(sigma), d = __ldg(dot);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(e % L.T), b = (int)((e / L.T) % L.B), a = (int)(e / ((long long)L.T * L.B));
    const int r = L.dim == 0 ? a : b;
    const int c = L.dim == 0 ? b * L.T + t : a * L.T + t;
    out[e] = (g[e] - d * __ldg(u + r) * __ldg(v + c)) * inv;
  }
}

// ------------------------------------------------------------------------------------------------ batched over hooks
// One forward of the projection discriminator runs 17 spectral-norm hooks (models/sngan_projection.py:110-181); launched
// one by one that is ~7 launches per hook, three forwards per step. These kernels run the SAME five phases for every hook
// of a forward at once (blockIdx.y = hook; the descriptor table travels by value in the kernel parameters, so a CUDA
// graph node owns it): Wt u -> normalise v -> W v -> normalise u + sigma -> W / sigma.
constexpr int kSnRowChunk = 16;  // rows of W per block of the batched W^T u kernel

struct SnBatch {
  const float* w[GP_SN_MAX];
  float* u[GP_SN_MAX];
  float* v[GP_SN_MAX];
  float* out[GP_SN_MAX];
  float* sigma[GP_SN_MAX];
  float* u_keep[GP_SN_MAX];  // copies of the u / v this forward used (later forwards advance u, v in place)
  float* v_keep[GP_SN_MAX];
  float* t2[GP_SN_MAX];      // scratch [rows]
  float* t1[GP_SN_MAX];      // scratch [cols]
  int A[GP_SN_MAX], B[GP_SN_MAX], T[GP_SN_MAX], dim[GP_SN_MAX];
  int count;
  float eps;
};

__global__ void snb_gemv_t_kernel(const __grid_constant__ SnBatch b) {
  gp::pdl_sync();
  const int m = blockIdx.y;
  const SnLayout L{b.A[m], b.B[m], b.T[m], b.dim[m]};
  const int cb = (L.cols() + 255) / 256, rc = (L.rows() + kSnRowChunk - 1) / kSnRowChunk;
  if ((int)blockIdx.x >= cb * rc) return;
  const int c = (blockIdx.x % cb) * 256 + threadIdx.x;
  if (c >= L.cols()) return;
  const int r0 = (blockIdx.x / cb) * kSnRowChunk, r1 = min(r0 + kSnRowChunk, L.rows());
  const float* w = b.w[m];
  const float* u = b.u[m];
  float acc = 0.f;
  int r = r0;
  for (; r + 8 <= r1; r += 8) {   // eight independent loads in flight per thread
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = __ldg(w + L.addr(r + k, c));
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += x[k] * __ldg(u + r + k);
  }
  for (; r < r1; ++r) acc += __ldg(w + L.addr(r, c)) * __ldg(u + r);
  atomicAdd(b.t1[m] + c, acc);
}

// which = 0: v <- t1 / max(||t1||, eps) (+ copy);  which = 1: u <- t2 / ..., sigma = ||t2||^2 / max(||t2||, eps) (+ copy)
__global__ void snb_normalize_kernel(const __grid_constant__ SnBatch b, int which) {
  gp::pdl_sync();
  const int m = blockIdx.x;
  const SnLayout L{b.A[m], b.B[m], b.T[m], b.dim[m]};
  const int n = which == 0 ? L.cols() : L.rows();
  const float* t = which == 0 ? b.t1[m] : b.t2[m];
  float* out = which == 0 ? b.v[m] : b.u[m];
  float* keep = which == 0 ? b.v_keep[m] : b.u_keep[m];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += t[i] * t[i];
  acc = block_sum(acc);
  const float nrm = fmaxf(sqrtf(acc), b.eps);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = t[i] / nrm;
    out[i] = x;
    if (keep != nullptr) keep[i] = x;
  }
  if (which == 1 && threadIdx.x == 0) *b.sigma[m] = acc / nrm;
}

__global__ void snb_gemv_kernel(const __grid_constant__ SnBatch b) {
  gp::pdl_sync();
  const int m = blockIdx.y, r = blockIdx.x;
  const SnLayout L{b.A[m], b.B[m], b.T[m], b.dim[m]};
  if (r >= L.rows()) return;
  const float* w = b.w[m];
  const float* v = b.v[m];
  float acc = 0.f;
  const int cols = L.cols();
  if (L.dim == 0 && (cols & 3) == 0 && ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(v)) & 15) == 0) {
    // a row of the matrix is contiguous: 16-byte loads, two per thread in flight
    const float4* wr = reinterpret_cast<const float4*>(w + (long long)r * cols);
    const float4* v4 = reinterpret_cast<const float4*>(v);
    const int n4 = cols >> 2;
    int c = threadIdx.x;
    for (; c + (int)blockDim.x < n4; c += 2 * blockDim.x) {
      const float4 a0 = __ldg(wr + c), a1 = __ldg(wr + c + blockDim.x);
      const float4 b0 = __ldg(v4 + c), b1 = __ldg(v4 + c + blockDim.x);
      acc += a0.x * b0.x + a0.y * b0.y + a0.z * b0.z + a0.w * b0.w + a1.x * b1.x + a1.y * b1.y + a1.z * b1.z + a1.w * b1.w;
    }
    for (; c < n4; c += blockDim.x) {
      const float4 a0 = __ldg(wr + c), b0 = __ldg(v4 + c);
      acc += a0.x * b0.x + a0.y * b0.y + a0.z * b0.z + a0.w * b0.w;
    }
  } else {
    for (int c = threadIdx.x; c < cols; c += blockDim.x) acc += __ldg(w + L.addr(r, c)) * __ldg(v + c);
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) b.t2[m][r] = acc;
}

// eval mode: sigma = u . (W v) with the stored u, v (+ copies for the backward)
__global__ void snb_eval_sigma_kernel(const __grid_constant__ SnBatch b) {
  gp::pdl_sync();
  const int m = blockIdx.x;
  const SnLayout L{b.A[m], b.B[m], b.T[m], b.dim[m]};
  float acc = 0.f;
  for (int i = threadIdx.x; i < L.rows(); i += blockDim.x) {
    acc += b.u[m][i] * b.t2[m][i];
    if (b.u_keep[m] != nullptr) b.u_keep[m][i] = b.u[m][i];
  }
  if (b.v_keep[m] != nullptr)
    for (int i = threadIdx.x; i < L.cols(); i += blockDim.x) b.v_keep[m][i] = b.v[m][i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) *b.sigma[m] = acc;
}

__global__ void snb_scale_kernel(const __grid_constant__ SnBatch b) {
  gp::pdl_sync();
  const int m = blockIdx.y;
  const long long n = (long long)b.A[m] * b.B[m] * b.T[m];
  const float inv = 1.f / __ldg(b.sigma[m]);
  const float* w = b.w[m];
  float* out = b.out[m];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = w[i] * inv;
}

// backward, batched: dot[m] = <g, w_sn>;  out = (g - dot u v^T) / sigma. A hook whose output got no gradient has g = NULL.
struct SnGradBatch {
  const float* g[GP_SN_MAX];
  const float* w_sn[GP_SN_MAX];
  const float* u[GP_SN_MAX];
  const float* v[GP_SN_MAX];
  const float* sigma[GP_SN_MAX];
  float* out[GP_SN_MAX];
  float* dot;  // [count], zeroed by the entry point
  int A[GP_SN_MAX], B[GP_SN_MAX], T[GP_SN_MAX], dim[GP_SN_MAX];
  int count;
};

__global__ void snb_grad_dot_kernel(const __grid_constant__ SnGradBatch b) {
  gp::pdl_sync();
  const int m = blockIdx.y;
  if (b.g[m] == nullptr) return;
  const long long n = (long long)b.A[m] * b.B[m] * b.T[m];
  const float* g = b.g[m];
  const float* w = b.w_sn[m];
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += g[i] * w[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0 && acc != 0.f) atomicAdd(b.dot + m, acc);
}

__global__ void snb_grad_kernel(const __grid_constant__ SnGradBatch b) {
  gp::pdl_sync();
  const int m = blockIdx.y;
  if (b.g[m] == nullptr) return;
  const SnLayout L{b.A[m], b.B[m], b.T[m], b.dim[m]};
  const long long n = (long long)L.A * L.B * L.T;
  const float inv = 1.f / __ldg(b.sigma[m]), d = b.dot[m];
  const float* g = b.g[m];
  const float* u = b.u[m];
  const float* v = b.v[m];
  float* out = b.out[m];
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(e % L.T), bb = (int)((e / L.T) % L.B), a = (int)(e / ((long long)L.T * L.B));
    const int r = L.dim == 0 ? a : bb;
    const int c = L.dim == 0 ? bb * L.T + t : a * L.T + t;
    out[e] = (g[e] - d * __ldg(u + r) * __ldg(v + c)) * inv;
  }
}

static inline int grid1d(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace gp

using namespace gp;

extern "C" {

// scratch: fp32 [rows + cols] (t2 then t1). training != 0: u, v are updated in place (one power iteration).
int gp_sn_sigma(const float* w, int A, int B, int T, int dim, float* u, float* v, float eps, int training,
                float* scratch, float* sigma, void* stream) {
  GP_REQUIRE(w && u && v && scratch && sigma && A > 0 && B > 0 && T > 0 && (dim == 0 || dim == 1), "gp_sn_sigma: bad arguments");
  SnLayout L{A, B, T, dim};
  const int R = L.rows(), Cc = L.cols();
  float* t2 = scratch;
  float* t1 = scratch + R;
  cudaStream_t st = as_stream(stream);
  if (training) {
    GP_CHECK_CUDA(cudaMemsetAsync(t1, 0, sizeof(float) * Cc, st));
    const int rpb = R > 64 ? 64 : R;
    dim3 grid((Cc + 255) / 256, (R + rpb - 1) / rpb);
    gp::launch_pdl(sn_gemv_t_kernel, grid, 256, 0, st, w, L, u, t1, rpb);
    GP_CHECK_LAUNCH();
    gp::launch_pdl(sn_normalize_kernel, 1, 1024, 0, st, t1, Cc, eps, v, nullptr);
    GP_CHECK_LAUNCH();
    gp::launch_pdl(sn_gemv_kernel, R, 256, 0, st, w, L, v, t2);
    GP_CHECK_LAUNCH();
    gp::launch_pdl(sn_normalize_kernel, 1, 1024, 0, st, t2, R, eps, u, sigma);
    GP_CHECK_LAUNCH();
  } else {
    gp::launch_pdl(sn_gemv_kernel, R, 256, 0, st, w, L, v, t2);
    GP_CHECK_LAUNCH();
    gp::launch_pdl(sn_dot_kernel, 1, 1024, 0, st, u, t2, R, sigma);
    GP_CHECK_LAUNCH();
  }
  return GP_OK;
}

int gp_sn_scale(const float* w, const float* sigma, float* out, long long n, void* stream) {
  GP_REQUIRE(w && sigma && out && n > 0, "gp_sn_scale: bad arguments");
  gp::launch_pdl(sn_scale_kernel, grid1d(n), 256, 0, as_stream(stream), w, sigma, out, n);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

// dot: fp32 scalar scratch. out = (g - <g, w_sn> u v^T) / sigma in the layout of w.
int gp_sn_grad(const float* g, const float* w_sn, int A, int B, int T, int dim, const float* u, const float* v,
               const float* sigma, float* dot, float* out, void* stream) {
  GP_REQUIRE(g && w_sn && u && v && sigma && dot && out, "gp_sn_grad: bad arguments");
  SnLayout L{A, B, T, dim};
  const long long n = (long long)A * B * T;
  cudaStream_t st = as_stream(stream);
  GP_CHECK_CUDA(cudaMemsetAsync(dot, 0, sizeof(float), st));
  gp::launch_pdl(sn_grad_dot_kernel, grid1d(n), 256, 0, st, g, w_sn, n, dot);
  GP_CHECK_LAUNCH();
  gp::launch_pdl(sn_grad_kernel, grid1d(n), 256, 0, st, g, L, u, v, sigma, dot, out);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

// Every spectral-norm hook of one forward in five launches (+ one memset). scratch: fp32 [sum over hooks of rows + cols];
// keep (optional): fp32 of the same size, receives the u | v each hook used (what the backward needs).
int gp_sn_batched(const gp_sn_batch_t* p, void* stream) {
  GP_REQUIRE(p != nullptr && p->count > 0 && p->count <= GP_SN_MAX && p->scratch != nullptr, "gp_sn_batched: bad arguments");
  SnBatch b;
  memset(&b, 0, sizeof(b));
  b.count = p->count;
  b.eps = p->eps;
  long long off = 0, max_elems = 0;
  int max_rows = 0, max_tblocks = 0;
  for (int m = 0; m < p->count; ++m) {
    GP_REQUIRE(p->w[m] && p->u[m] && p->v[m] && p->out[m] && p->sigma[m] && p->A[m] > 0 && p->B[m] > 0 && p->T[m] > 0 &&
                   (p->dim[m] == 0 || p->dim[m] == 1),
               "gp_sn_batched: bad hook %d", m);
    const SnLayout L{p->A[m], p->B[m], p->T[m], p->dim[m]};
    b.w[m] = p->w[m], b.u[m] = p->u[m], b.v[m] = p->v[m], b.out[m] = p->out[m], b.sigma[m] = p->sigma[m];
    b.A[m] = L.A, b.B[m] = L.B, b.T[m] = L.T, b.dim[m] = L.dim;
    b.t2[m] = p->scratch + off;
    b.t1[m] = p->scratch + off + L.rows();
    if (p->keep != nullptr) {
      b.u_keep[m] = p->keep + off;
      b.v_keep[m] = p->keep + off + L.rows();
    }
    off += L.rows() + L.cols();
    const long long n = (long long)L.A * L.B * L.T;
    if (n > max_elems) max_elems = n;
    if (L.rows() > max_rows) max_rows = L.rows();
    const int tb = ((L.cols() + 255) / 256) * ((L.rows() + kSnRowChunk - 1) / kSnRowChunk);
    if (tb > max_tblocks) max_tblocks = tb;
  }
  cudaStream_t st = as_stream(stream);
  if (p->training) {
    GP_CHECK_CUDA(cudaMemsetAsync(p->scratch, 0, sizeof(float) * off, st));
    gp::launch_pdl(snb_gemv_t_kernel, dim3(max_tblocks, b.count), 256, 0, st, b);
    GP_CHECK_LAUNCH();
    gp::launch_pdl(snb_normalize_kernel, b.count, 1024, 0, st, b, 0);
    GP_CHECK_LAUNCH();
    gp::launch_pdl(snb_gemv_kernel, dim3(max_rows, b.count), 256, 0, st, b);
    GP_CHECK_LAUNCH();
    gp::launch_pdl(snb_normalize_kernel, b.count, 1024, 0, st, b, 1);
    GP_CHECK_LAUNCH();
  } else {
    gp::launch_pdl(snb_gemv_kernel, dim3(max_rows, b.count), 256, 0, st, b);
    GP_CHECK_LAUNCH();
    gp::launch_pdl(snb_eval_sigma_kernel, b.count, 1024, 0, st, b);
    GP_CHECK_LAUNCH();
  }
  int gx = (int)((max_elems + 255) / 256);
  const int cap = num_sms() * 4;
  if (gx > cap) gx = cap;
  gp::launch_pdl(snb_scale_kernel, dim3(gx, b.count), 256, 0, st, b);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

int gp_sn_grad_batched(const gp_sn_grad_batch_t* p, void* stream) {
  GP_REQUIRE(p != nullptr && p->count > 0 && p->count <= GP_SN_MAX && p->dot != nullptr, "gp_sn_grad_batched: bad arguments");
  SnGradBatch b;
  memset(&b, 0, sizeof(b));
  b.count = p->count;
  b.dot = p->dot;
  long long max_elems = 0;
  for (int m = 0; m < p->count; ++m) {
    b.g[m] = p->g[m];
    if (p->g[m] == nullptr) continue;
    GP_REQUIRE(p->w_sn[m] && p->u[m] && p->v[m] && p->sigma[m] && p->out[m], "gp_sn_grad_batched: bad hook %d", m);
    b.w_sn[m] = p->w_sn[m], b.u[m] = p->u[m], b.v[m] = p->v[m], b.sigma[m] = p->sigma[m], b.out[m] = p->out[m];
    b.A[m] = p->A[m], b.B[m] = p->B[m], b.T[m] = p->T[m], b.dim[m] = p->dim[m];
    const long long n = (long long)p->A[m] * p->B[m] * p->T[m];
    if (n > max_elems) max_elems = n;
  }
  if (max_elems == 0) return GP_OK;
  cudaStream_t st = as_stream(stream);
  GP_CHECK_CUDA(cudaMemsetAsync(p->dot, 0, sizeof(float) * p->count, st));
  int gx = (int)((max_elems + 255) / 256);
  const int cap = num_sms() * 4;
  if (gx > cap) gx = cap;
  gp::launch_pdl(snb_grad_dot_kernel, dim3(gx, b.count), 256, 0, st, b);
  GP_CHECK_LAUNCH();
  gp::launch_pdl(snb_grad_kernel, dim3(gx, b.count), 256, 0, st, b);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

}  // extern "C"
