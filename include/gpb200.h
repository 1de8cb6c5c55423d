/*
 * gpb200.h — C ABI of the B200-native hot path of gan-playground (one adversarial G+D training step of the
 * DCGAN-family conv nets). Plain pointers and sizes only; no torch types. Every entry point enqueues work on
 * the given CUDA stream (a cudaStream_t passed as void*) and returns 0 on success or a negative code;
 * gp_last_error() returns the message for the calling thread. All device pointers are owned by the caller
 * (torch's caching allocator in the Python host); the library never allocates caller-visible memory.
 *
 * Each function cites the reference call site it replaces (paths relative to the reference repo; `torch:` =
 * the third-party PyTorch implementation the reference delegates to).
 *
 * Activation tensors are NHWC bf16 ("pixels x channels"); packed weights are bf16 [Nout][tap][C];
 * weight gradients are fp32 [Cdense][tap][Cgath]. Layout conversions to/from torch's fp32 NCHW / OIHW
 * happen only at network boundaries and in gp_pack_weight / gp_unpack_wgrad.
 */
#ifndef GPB200_H_
#define GPB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GP_OK 0
#define GP_ERR_INVALID (-1)
#define GP_ERR_CUDA (-2)
#define GP_ERR_UNSUPPORTED (-3)

/* activation codes fused into epilogues (models/dcgan.py:39 ReLU, :109 LeakyReLU(0.2), :43 Tanh) */
#define GP_ACT_NONE 0
#define GP_ACT_RELU 1
#define GP_ACT_LRELU 2
#define GP_ACT_TANH 3

/* convolution geometry kinds */
#define GP_KIND_CONV_K4S2 0  /* nn.Conv2d(k=4, s=2, p=1) fprop, or ConvTranspose2d(4,2,1) dgrad  (models/dcgan.py:106) */
#define GP_KIND_CONVT_K4S2 1 /* nn.ConvTranspose2d(k=4, s=2, p=1) fprop, or Conv2d(4,2,1) dgrad  (models/dcgan.py:36)  */
#define GP_KIND_CONV_K3S1 2  /* nn.Conv2d(k=3, s=1, p=1) fprop / dgrad (models/sngan_projection.py:30,33)              */
#define GP_KIND_CONV_K1S1 3  /* nn.Conv2d(k=1) / nn.Linear as a 1-tap GEMM (models/sngan_projection.py:43; dcgan.py:32) */

const char* gp_version(void);
const char* gp_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
uint64_t gp_launch_count(void);

/* ---- implicit-GEMM convolution forward / data-gradient on tcgen05 (replaces aten::convolution reached from
 * models/dcgan.py:53-55,119-120 and the dgrad half of aten::convolution_backward).
 *   out[n,oh,ow,:] = act( sum_taps in[n, ih(tap), iw(tap), :] . w[:, tap, :] + bias )
 * in  : bf16 NHWC (NB, Hin, Win, Cin)          w   : bf16 packed [Nout][taps][Cin]
 * out : bf16 NHWC (NB, Hout, Wout, Nout)       bias: fp32 [Nout] or NULL
 * col_sum/col_sumsq (optional, fp32 [Nout], caller-zeroed): per-channel sum and sum of squares of the fp32
 * pre-activation output, accumulated from the fp32 accumulators (BatchNorm batch statistics,
 * torch: aten::native_batch_norm). */
typedef struct {
  const void* in;
  const void* w;
  const float* bias;
  void* out;
  float* col_sum;
  float* col_sumsq;
  int32_t NB, Hin, Win, Cin;
  int32_t Hout, Wout, Nout;
  int32_t kind;
  int32_t act;
} gp_conv_fwd_t;
int gp_conv_fwd(const gp_conv_fwd_t* p, void* stream);

/* ---- weight gradient (the wgrad half of aten::convolution_backward).
 *   dw[m, tap, n] += sum_pixels dense[pix, m] * gath[gather(pix, tap), n]
 * dense: bf16 NHWC on the small pixel grid (NB, Hs, Ws, Cd);  gath: bf16 NHWC (NB, Hg, Wg, Cg).
 * GP_KIND_CONV_K4S2 : Hg = 2*Hs. For Conv2d: dense = dY, gath = X  -> dw[Cout][kh][kw][Cin].
 *                     For ConvTranspose2d: dense = X, gath = dY    -> dw[Cin][kh][kw][Cout].
 * dw is fp32 [Cd][taps][Cg] and MUST be zeroed by the caller (split-K partial sums are accumulated atomically). */
typedef struct {
  const void* dense;
  const void* gath;
  float* dw;
  int32_t NB, Hs, Ws, Cd;
  int32_t Hg, Wg, Cg;
  int32_t kind;
} gp_conv_wgrad_t;
int gp_conv_wgrad(const gp_conv_wgrad_t* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GPB200_H_ */
