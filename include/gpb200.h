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

/* Companion tensor of an NHWC activation. Every activation has a bf16 tensor — what autograd sees and
 * every backward kernel / GEMM reads — and, depending on the forward precision mode that produced it, a second tensor of
 * the same shape: GP_COMP_LO = bf16(value - bf16(value)) ("bf16x3": value = hi + lo), GP_COMP_F16 = fp16(value) ("fp16":
 * the operand of the next single-MMA forward GEMM). Forward element-wise kernels taking (x, x_comp, ..., comp_fmt) read
 * the most precise view and write both tensors; a NULL companion pointer = that tensor has none. */
#define GP_COMP_NONE 0
#define GP_COMP_LO 1
#define GP_COMP_F16 2

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
  int32_t act;            /* GP_ACT_NONE / RELU / LRELU (tanh lives in gp_col2im_k4s2) */
  const void* residual;   /* optional bf16 tensor with the layout of `out`, added before the activation
                             (residual blocks of models/sngan_projection.py:66,136); NULL = none */
  /* bf16x3 forward precision mode (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, fp32 accumulate): */
  const void* in_lo;      /* low halves of the input (same layout as `in`); then w is [Nout][2][taps][Cin] (hi | lo) */
  void* out_lo;           /* optional: also write bf16(v - bf16(v)) here (hi/lo output pair)                        */
  float* out_f32;         /* optional: write fp32 here instead of bf16 `out` (pre-BatchNorm outputs)                */
  /* "fp16" forward precision mode (ONE MMA on fp16 operands: 11 significant bits instead of bf16's 8): */
  int32_t flags;          /* GP_CONV_IN_F16: `in` and `w` hold fp16 (same layouts; in_lo must be NULL);
                             GP_CONV_LO_F16: out_lo receives fp16(v) — the next layer's operand — instead of the residual */
  /* Backward-side fusions for a GEMM that runs as the DATA GRADIENT of the next layer (its output is dA of the layer that
   * produced its input): work of the PRODUCING layer's backward, which would otherwise start with a pass over dA:
   *   GP_CONV_BWD_MASK   : out = v * act'(a); bwd_src = a, the producer's bf16 activation (layout of `out`); ReLU /
   *                        LeakyReLU by sign, bwd_slope = 0 / 0.2. Replaces gp_act_bwd of a block without BatchNorm.
   *   GP_CONV_BWD_BN_F32 / _BF16: out = v unchanged; col_sum += sum dz, col_sumsq += sum dz * xhat with
   *                        dz = bf16(v) * act'(scale*y + shift), xhat = (y - mean) * rstd; bwd_src = y, the producer's
   *                        pre-BatchNorm tensor (fp32 / bf16, layout of `out`); bwd_fin = fp32 [4][Nout]: mean | rstd |
   *                        scale | shift (the gp_bn_finalize outputs). Replaces gp_bn_bwd_reduce[_f32].
   * Requires a plain bf16 output: no bias, residual, activation, in_lo, out_lo, out_f32. */
  const void* bwd_src;
  const float* bwd_fin;
  float bwd_slope;
  int32_t bwd_mode;
} gp_conv_fwd_t;
#define GP_CONV_BWD_NONE 0
#define GP_CONV_BWD_MASK 1
#define GP_CONV_BWD_BN_F32 2
#define GP_CONV_BWD_BN_BF16 3
#define GP_CONV_IN_F16 1
#define GP_CONV_LO_F16 2
#define GP_CONV_RES_F16 4 /* `residual` holds fp16: the companion of the shortcut activation in the "fp16" mode */
#define GP_CONV_OUT_F16 8 /* `out` receives fp16(v) instead of bf16(v): 2-byte pre-BatchNorm storage of the "fp16" mode */
int gp_conv_fwd(const gp_conv_fwd_t* p, void* stream);
/* Host-only: the tile shape gp_conv_fwd would use for this problem (BN in {64,128,256} output columns, MT in {1,2}
 * 128-row sub-tiles) and the resulting number of output tiles. No pointers are dereferenced, nothing is launched. */
int gp_conv_fwd_plan(const gp_conv_fwd_t* p, int* bn, int* mt, int* tiles);

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

/* ---- weight staging: fp32 torch-layout parameters -> bf16 GEMM operands (and gradients back).
 * gp_pack_conv_weight: src (D0, D1, taps) fp32 [Conv2d: (Cout, Cin, kh*kw); ConvTranspose2d: (Cin, Cout, kh*kw)]
 *   -> dst bf16 [N][tap][C] with (N, C) = (D0, D1) when (n_dim & 1) == 0, (D1, D0) when (n_dim & 1) == 1;
 *   n_dim & 2 reverses the tap order (dgrad of a stride-1 conv = conv with the flipped kernel).
 *   inv_scale (optional device scalar): every element is divided by it — the W / sigma of spectral norm
 *   (torch:nn/utils/spectral_norm.py:112) fused into the staging pass.
 * gp_unpack_conv_wgrad: fp32 [M][tap][N] -> fp32 (M, N, tap), i.e. the torch layout of the parameter gradient.
 * gp_pack_matrix: dst[r][k] = src[map(r)*s_r + k*s_k] / inv_scale for r < R, k < K, zero padding up to [Rpad][ld_dst];
 *   perm > 1 reorders rows from NCHW-flatten (c*perm + hw) to NHWC-flatten (hw*(R/perm) + c) order
 *   (the view(B, C, 4, 4) after the generator's Linear, models/dcgan.py:50-51).
 * gp_unpack_matrix: the inverse mapping for fp32 gradients.
 * GP_UNPACK_ACCUMULATE or-ed into `taps` (gp_unpack_conv_wgrad) / `perm` (gp_unpack_matrix): dst += value instead of
 * dst = value — the destination is the parameter's existing .grad buffer (gradients of the D-real and D-fake passes
 * accumulate there, main_dcgan.py:73,84), so no separate accumulation pass runs afterwards. */
#define GP_UNPACK_ACCUMULATE (1 << 30)
int gp_pack_conv_weight(const float* src, void* dst, int D0, int D1, int taps, int n_dim, const float* inv_scale,
                        void* stream);
/* gp_stage_conv_weights: the batched form for 16-tap (4x4) weights — every listed layer in every listed orientation /
 * format from ONE launch that reads each fp32 source once (the per-tensor calls above cost ~30 launches per DCGAN step
 * whose bytes do not shrink with the batch). Destination rows are [N][tap][C] as gp_pack_conv_weight writes them:
 * GP_STAGE_BF16 bf16, GP_STAGE_F16 fp16 (the single-MMA fp16 operand), GP_STAGE_SPLIT bf16 hi block | lo block per row
 * (gp_split_conv_weight's layout); ld = elements per destination row; D0, D1 multiples of 32. tile0 / total_tiles are
 * filled by the library. */
#define GP_STAGE_MAX_LAYERS 12
#define GP_STAGE_MAX_DST 4
#define GP_STAGE_BF16 0
#define GP_STAGE_F16 1
#define GP_STAGE_SPLIT 2
typedef struct {
  void* ptr;
  long long ld;
  int32_t n_dim;
  int32_t fmt;
} gp_stage_dst_t;
typedef struct {
  const float* src;   /* (D0, D1, 16) fp32, torch layout of Conv2d (Cout, Cin, 4, 4) / ConvTranspose2d (Cin, Cout, 4, 4) */
  int32_t D0, D1;
  int32_t tile0;
  int32_t ndst;
  gp_stage_dst_t dst[GP_STAGE_MAX_DST];
} gp_stage_layer_t;
typedef struct {
  int32_t count;
  int32_t total_tiles;
  gp_stage_layer_t layer[GP_STAGE_MAX_LAYERS];
} gp_stage_table_t;
int gp_stage_conv_weights(const gp_stage_table_t* table, void* stream);
int gp_unpack_conv_wgrad(const float* src, float* dst, int M, int N, int taps, void* stream);
int gp_pack_matrix(const float* src, void* dst, int R, int K, int Rpad, int ld_dst, long long s_r, long long s_k,
                   int perm, const float* inv_scale, void* stream);
int gp_unpack_matrix(const float* src, float* dst, int R, int K, int ld_src, long long s_r, long long s_k, int perm,
                     void* stream);

/* ---- BatchNorm2d in training mode with fused activation (replaces aten::native_batch_norm(+_backward) and the
 * in-place ReLU / LeakyReLU that follow it: models/dcgan.py:37-39,107-109). x, y, out, da, dy: bf16 [P][C].
 * gp_bn_stats     : sum[c] += sum_p x, sumsq[c] += sum_p x^2        (caller zeroes; all-reduce these for SyncBN)
 * gp_bn_finalize  : mean, rstd (biased variance, eps), scale = gamma*rstd, shift = beta - mean*scale;
 *                   running_mean/var (unbiased variance, momentum) and num_batches_tracked updated when non-NULL
 * gp_bn_apply_act : out = act(y*scale + shift)
 * gp_bn_bwd_reduce: sum_dz[c] += sum dz, sum_dzx[c] += sum dz*xhat with dz = da*act'(y*scale+shift)  (= dbeta, dgamma)
 * gp_bn_bwd_apply : dy = scale * (dz - sum_dz/count - xhat*sum_dzx/count)
 * gp_act_bwd      : dy = da * act'(.) given the activation OUTPUT a (layers without BatchNorm)
 * gp_colsum       : out[c] += sum_p x[p][c]  (bias gradients) */
int gp_bn_stats(const void* x, long long P, int C, float* sum, float* sumsq, void* stream);
int gp_bn_finalize(const float* sum, const float* sumsq, double count, int C, float eps, float momentum,
                   const float* gamma, const float* beta, float* mean, float* rstd, float* scale, float* shift,
                   float* running_mean, float* running_var, long long* num_batches_tracked, void* stream);
/* eval mode (module.eval()): mean/rstd/scale/shift from the running statistics */
int gp_bn_eval_params(const float* running_mean, const float* running_var, const float* gamma, const float* beta, int C,
                      float eps, float* mean, float* rstd, float* scale, float* shift, void* stream);
int gp_bn_apply_act(const void* y, void* out, long long P, int C, const float* scale, const float* shift, int act,
                    void* stream);
int gp_bn_bwd_reduce(const void* da, const void* y, long long P, int C, const float* scale, const float* shift,
                     const float* mean, const float* rstd, int act, float* sum_dz, float* sum_dzx, void* stream);
/* acc_dbeta / acc_dgamma (optional, both or neither): the affine parameters' gradient buffers (fp32 [C]);
 * acc_dbeta += sum_dz * acc_scale, acc_dgamma += sum_dzx * acc_scale in the same launch (acc_scale = 1 / world_size) —
 * the gradient lands in `.grad` without an accumulation pass. */
int gp_bn_bwd_apply(const void* da, const void* y, void* dy, long long P, int C, const float* scale,
                    const float* shift, const float* mean, const float* rstd, const float* sum_dz,
                    const float* sum_dzx, double count, int act, float* acc_dbeta, float* acc_dgamma, float acc_scale,
                    void* stream);
int gp_act_bwd(const void* da, const void* a, void* dy, long long n, int act, void* stream);
int gp_colsum(const void* x, long long P, int C, float* out, void* stream);

/* ---- the two image-side layers (Cin = img_dim of D's first Conv2d, models/dcgan.py:106; Cout = img_dim of G's last
 * ConvTranspose2d + Tanh, models/dcgan.py:41-44) as im2col / col2im around a 1-tap tensor-core GEMM.
 * col is bf16 [NB*(Hi/2)*(Wi/2)][64], column (c*4 + kh)*4 + kw, zero for columns >= ch*16; img is fp32 NCHW (NB, ch, Hi, Wi).
 * gp_im2col_k4s2: col = im2col(img * (mul ? 1 - mul^2 : 1))   (mul = tanh output -> fused tanh')
 * gp_col2im_k4s2: img = act(bias[c] + col2im(col)) */
int gp_im2col_k4s2(const float* img, const float* mul, void* col, int NB, int ch, int Hi, int Wi, void* stream);
int gp_col2im_k4s2(const void* col, const float* bias, float* img, int NB, int ch, int Hi, int Wi, int act,
                   void* stream);
/* dbias[c] += sum dout[n,c,:,:] * (mul ? 1 - mul^2 : 1); dbias fp32 [ch], caller-zeroed */
int gp_image_bias_grad(const float* dout, const float* mul, float* dbias, int NB, int ch, int HW, void* stream);

/* ---- the same two layers WITHOUT the column buffer: the im2col tile of 128 pixels is built in shared
 * memory from the fp32 NCHW image rows and consumed by tcgen05.mma in place; ch == 3, Wo = Wi/2 divides 128,
 * (Hi/2)*(Wi/2) % 128 == 0, channel count % 8 == 0 and <= 128 (other shapes: the im2col / col2im entry points above).
 * gp_image_conv_k4s2_fwd  (models/dcgan.py:106-109, Conv2d(img_dim, ndf, 4, 2, 1) + LeakyReLU on the image):
 *   out[px][n] = act(bias[n] + sum_j col[px][j] * w[n][j]), col = im2col(img * (mul ? 1 - mul^2 : 1)) as above,
 *   w fp32 [Cout][ch*16] = the Conv2d weight in torch's layout; out bf16 NHWC; comp_fmt / out_comp: GP_COMP_NONE (bf16
 *   operands), GP_COMP_LO (bf16x3 operands, out_comp = bf16 low halves), GP_COMP_F16 (fp16 operands, out_comp = fp16 copy).
 *   With mul = the tanh output and w = the ConvTranspose2d weight [Cin][ch*16] of models/dcgan.py:41-44 it is that layer's
 *   data gradient (dx[px][ci] from d(image)).
 * gp_image_conv_k4s2_wgrad: dw[m][j] += sum_px dense[px][m] * col[px][j] (fp32, torch layout [M][ch*16], accumulated);
 *   dbias[m] += sum_px dense[px][m] when dbias != NULL. dense bf16 [pixels][M]: dy of D's first conv, or x of G's last
 *   ConvTranspose2d (then mul = tanh output).
 * gp_image_convt_k4s2_fwd (models/dcgan.py:41-44, ConvTranspose2d(ngf, img_dim, 4, 2, 1) + Tanh; also the image gradient
 *   of D's first conv): img = act(bias[c] + col2im(x * w)), x 2-byte NHWC (NB, Hs, Ws, C) in format fmt (GP_COMP_NONE bf16,
 *   GP_COMP_LO bf16 hi + x_lo, GP_COMP_F16 fp16), w fp32 [C][ch*16], img fp32 NCHW (NB, ch, 2Hs, 2Ws); C % 16 == 0, <= 64. */
int gp_image_conv_k4s2_fwd(const float* img, const float* mul, const float* w, const float* bias, void* out,
                           void* out_comp, int comp_fmt, int NB, int ch, int Hi, int Wi, int Cout, int act, void* stream);
int gp_image_conv_k4s2_wgrad(const void* dense, const float* img, const float* mul, float* dw, float* dbias, int NB,
                             int ch, int Hi, int Wi, int M, void* stream);
int gp_image_convt_k4s2_fwd(const void* x, const void* x_lo, int fmt, const float* w, const float* bias, float* img,
                            int NB, int Hs, int Ws, int C, int ch, int act, void* stream);

/* ---- discriminator heads: out[b][o] = bias[o] + sum_{hw,c} a[b,hw,c] * w[o*s_o + c*s_c + hw*s_hw]
 * (global sum pooling + Linear, models/dcgan.py:121-122: s_hw = 0; flatten + Linear, models/dcgan_specnorm.py:125-126:
 * s_c = HW, s_hw = 1; also the projection inner product of models/sngan_projection.py:193-195 with a gathered w).
 * gp_head_bwd: da (bf16, may be NULL), dw (fp32, caller-zeroed, may be NULL), dbias (fp32 [O], may be NULL). */
int gp_head_fwd(const void* a, const float* w, const float* bias, float* out, int NB, int HW, int C, int O,
                long long s_o, long long s_c, long long s_hw, void* stream);
int gp_head_bwd(const float* dout, const void* a, const float* w, void* da, float* dw, float* dbias, int NB, int HW,
                int C, int O, long long s_o, long long s_c, long long s_hw, void* stream);

/* ---- GANLoss (utils/criterion.py:30-41): mean loss and d(loss)/d(pred) in one pass.
 * mode 0: BCE-with-logits vs constant `target`; 1: MSE vs `target`; 2: hinge, real: relu(1-p);
 * 3: hinge, fake: relu(1+p); 4: generator hinge: -p. */
#define GP_LOSS_BCE 0
#define GP_LOSS_MSE 1
#define GP_LOSS_HINGE_REAL 2
#define GP_LOSS_HINGE_FAKE 3
#define GP_LOSS_NEG_MEAN 4
int gp_gan_loss(const float* pred, int n, int mode, float target, float* loss, float* dpred, void* stream);

/* ---- ACGAN objective (main_acgan.py:95-97,114-116,129-131: criterion_adv(outD_adv, ...) + 0.5 * MSELoss(outD_cls, c))
 * on the packed two-head logits fp32 [NB][1 + K] (column 0: out_layer, columns 1..K: out_aux, models/acgan.py:122-126)
 * and the float label vectors fp32 [NB][K]; value and gradient in one pass.
 * out4 = { GANLoss term (mode / target as gp_gan_loss), MSE term (mean over NB*K, unweighted), term0 + aux_weight*term1,
 *          mean(sigmoid(column 0)) — the D(x) / D(G(z)) print of :94,112,127 };  dlogits = d out4[2] / d logits. */
int gp_acgan_loss(const float* logits, const float* labels, int NB, int K, int mode, float target, float aux_weight,
                  float* out4, float* dlogits, void* stream);

/* ---- spectral normalisation (torch.nn.utils.spectral_norm's pre-forward hook, torch:nn/utils/spectral_norm.py:62-114,
 * applied at models/dcgan_specnorm.py:37,42,107 and models/sngan_projection.py:110-181) as GEMV kernels.
 * w: fp32 parameter in torch's layout viewed as [A][B][T]; dim == 0: W_mat[a][b*T+t] (Conv2d/Linear/Embedding),
 * dim == 1: W_mat[b][a*T+t] (ConvTranspose2d). u: fp32 [rows], v: fp32 [cols].
 * gp_sn_sigma : training != 0 -> one power iteration in place (v = normalize(W^T u), u = normalize(W v), eps clamp),
 *               then sigma = u.(W v); training == 0 -> sigma from the stored u, v. scratch: fp32 [rows + cols].
 * gp_sn_scale : out = w / sigma (the normalised weight handed to the conv / linear kernels)
 * gp_sn_grad  : gradient through W / sigma with u, v held constant:
 *               out = (g - <g, w_sn> * u v^T) / sigma, in w's layout; dot: fp32 scalar scratch. */
int gp_sn_sigma(const float* w, int A, int B, int T, int dim, float* u, float* v, float eps, int training,
                float* scratch, float* sigma, void* stream);
int gp_sn_scale(const float* w, const float* sigma, float* out, long long n, void* stream);
int gp_sn_grad(const float* g, const float* w_sn, int A, int B, int T, int dim, const float* u, const float* v,
               const float* sigma, float* dot, float* out, void* stream);

/* Batched over the hooks of one forward (the projection discriminator has 17: models/sngan_projection.py:110-181): the same
 * semantics as gp_sn_sigma + gp_sn_scale for `count` weights in five launches. scratch: fp32 [sum_i rows_i + cols_i];
 * keep (optional, same size): receives the (u | v) each hook used in this forward — what gp_sn_grad_batched needs, since
 * later forwards advance u, v in place. gp_sn_grad_batched: hooks with g[i] == NULL are skipped; dot: fp32 [count]. */
#define GP_SN_MAX 24
typedef struct {
  const float* w[GP_SN_MAX];
  float* u[GP_SN_MAX];
  float* v[GP_SN_MAX];
  float* out[GP_SN_MAX];   /* w / sigma, same layout as w */
  float* sigma[GP_SN_MAX]; /* fp32 scalars */
  int32_t A[GP_SN_MAX], B[GP_SN_MAX], T[GP_SN_MAX], dim[GP_SN_MAX];
  int32_t count;
  int32_t training;
  float eps;
  float* scratch;
  float* keep;
} gp_sn_batch_t;
typedef struct {
  const float* g[GP_SN_MAX];
  const float* w_sn[GP_SN_MAX];
  const float* u[GP_SN_MAX];
  const float* v[GP_SN_MAX];
  const float* sigma[GP_SN_MAX];
  float* out[GP_SN_MAX];
  int32_t A[GP_SN_MAX], B[GP_SN_MAX], T[GP_SN_MAX], dim[GP_SN_MAX];
  int32_t count;
  float* dot;
} gp_sn_grad_batch_t;
int gp_sn_batched(const gp_sn_batch_t* p, void* stream);
int gp_sn_grad_batched(const gp_sn_grad_batch_t* p, void* stream);

/* ---- SNGAN projection networks (models/sngan_projection.py).
 * Conditional BatchNorm (:6-19): BatchNorm2d(affine=False) statistics come from gp_bn_stats / gp_bn_finalize
 * (gamma = beta = NULL); the per-sample scale / shift are gathered from the embedding table
 * emb fp32 [n_classes][2C] (gamma = emb[label][0:C], beta = emb[label][C:2C]) inside these kernels. emb == NULL: plain
 * non-affine normalisation. upsample != 0 fuses F.interpolate(scale_factor=2) (nearest, :53) into the write (forward)
 * and the 2x2 gradient sum into the read (backward): out / da then live on the (2H, 2W) grid.
 *   gp_cbn_apply_act : out = act(xhat * gamma[n] + beta[n])
 *   gp_cbn_bwd_reduce: part fp32 [NB][2][C] scratch; S fp32 [2][C] = (sum gamma*dz, sum gamma*dz*xhat) (all-reduce for
 *                      SyncBN); demb fp32 [n_classes][2C] = embedding gradient (may be NULL)
 *   gp_cbn_bwd_apply : dy = rstd * (gamma[n]*dz - S0/count - xhat*S1/count) */
int gp_cbn_apply_act(const void* y, const void* y_comp, void* out, void* out_comp, int comp_fmt, int NB, int H, int W,
                     int C, const float* mean, const float* rstd, const float* emb, const long long* labels, int act,
                     int upsample, void* stream);
int gp_cbn_bwd_reduce(const void* da, const void* y, const void* y_comp, int comp_fmt, int NB, int H, int W, int C,
                      const float* mean, const float* rstd, const float* emb, const long long* labels, int act,
                      int upsample, float* part, float* S, float* demb, int n_classes, void* stream);
int gp_cbn_bwd_apply(const void* da, const void* y, const void* y_comp, int comp_fmt, void* dy, int NB, int H, int W,
                     int C, const float* mean, const float* rstd, const float* emb, const long long* labels,
                     const float* S, double count, int act, int upsample, void* stream);
/* nearest x2 upsampling / 2x2 sum pooling with a scale (F.interpolate :53,60; F.avg_pool2d :128,132 = pool with 0.25;
 * each is the other's gradient). gp_pool2x: (Hout, Wout) is the pooled size. gp_act_fwd: out = act(in) (F.relu :122). */
int gp_upsample2x(const void* in, void* out, int NB, int H, int W, int C, float scale, void* stream);
int gp_pool2x(const void* in, const void* in_comp, void* out, void* out_comp, int comp_fmt, int NB, int Hout, int Wout,
              int C, float scale, void* stream);
int gp_act_fwd(const void* in, const void* in_comp, void* out, void* out_comp, int comp_fmt, long long n, int act,
               void* stream);
/* 3x3 image-side layers (first conv of the discriminator :141-148, last conv + tanh of the generator :80,95):
 * col bf16 [NB*H*W][32], column (c*3+kh)*3+kw; NHWC-8 bf16 <-> NCHW fp32 image with fused tanh / tanh'.
 * gp_nhwc8_to_image: in_f32 != 0 reads the GEMM's fp32 output (the precise modes keep the pre-tanh image out of bf16). */
int gp_im2col_k3s1(const float* img, void* col, void* col_comp, int comp_fmt, int NB, int ch, int H, int W, void* stream);
int gp_col2im_k3s1(const void* col, float* img, int NB, int ch, int H, int W, void* stream);
int gp_nhwc8_to_image(const void* in, int in_f32, float* img, int NB, int ch, int HW, int tanh_act, void* stream);
int gp_image_to_nhwc8_grad(const float* dout, const float* out, void* dy, int NB, int ch, int HW, int tanh_act,
                           void* stream);
/* projection head (:190-195): h = sum_hw relu(a) (fp32 [NB][C]); out[n] = b + sum_c h[n][c] * (w[c] + E[label[n]][c])
 * — the Linear l6 and the embedding inner product as one warp-level GEMV. Backward: dh, dw [C], db [1], dE [n_classes][C]. */
int gp_relu_sumpool(const void* a, const void* a_comp, int comp_fmt, float* h, int NB, int HW, int C, void* stream);
int gp_relu_sumpool_bwd(const float* dh, const void* a, void* da, int NB, int HW, int C, void* stream);
int gp_proj_head_fwd(const float* h, const float* w, const float* b, const float* E, const long long* labels, float* out,
                     int NB, int C, void* stream);
int gp_proj_head_bwd(const float* dout, const float* h, const float* w, const float* E, const long long* labels,
                     float* dh, float* dw, float* db, float* dE, int NB, int C, int n_classes, void* stream);

/* ---- "bf16x3" forward precision mode: the forward GEMMs consume hi/lo bf16 pairs
 * (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo) and pre-BatchNorm outputs stay fp32; these kernels produce / consume those formats.
 * gp_split_matrix     : like gp_pack_matrix, writing hi at dst[r*ld + k] and lo at dst[lo_off + r*ld + k] (k < width)
 * gp_split_conv_weight: like gp_pack_conv_weight, dst bf16 [N][2][taps][C] (hi block | lo block per row)
 * gp_bn_*_f32 / _split : the BatchNorm kernels above with y in fp32 and the activation written as a hi/lo pair
 * gp_im2col_k4s2_split / gp_col2im_k4s2_f32 / gp_head_fwd_split: image-side layers and head on those formats. */
/* "fp16" mode staging: out[r*ld_out + k] = fp16(hi[r*ld_in + k] + lo[r*ld_in + k]) for r < rows, k < cols — turns any
 * hi/lo bf16 pair produced by the kernels above (weights, im2col columns, latent rows) into the single fp16 operand.
 * gp_bn_apply_act_pair: gp_bn_apply_act_split with the second output = fp16(v) (operand of the next forward GEMM) next to
 * the bf16(v) the backward GEMMs read. */
/* BatchNorm statistics / apply on an EXISTING activation with a companion tensor (generator's b6 of
 * models/sngan_projection.py:92; every BatchNorm of models/dcgan_blur.py, which follows a BlurPool). */
int gp_bn_stats_comp(const void* x, const void* x_comp, int comp_fmt, long long P, int C, float* sum, float* sumsq,
                     void* stream);
int gp_bn_apply_act_comp(const void* y, const void* y_comp, void* out, void* out_comp, int comp_fmt, long long P, int C,
                         const float* scale, const float* shift, int act, void* stream);
/* backward of the same: y is read through its companion, i.e. exactly the value the forward normalised (the
 * activation mask and xhat must not come from the bf16 rounding of y) */
int gp_bn_bwd_reduce_comp(const void* da, const void* y, const void* y_comp, int comp_fmt, long long P, int C,
                          const float* scale, const float* shift, const float* mean, const float* rstd, int act,
                          float* sum_dz, float* sum_dzx, void* stream);
int gp_bn_bwd_apply_comp(const void* da, const void* y, const void* y_comp, int comp_fmt, void* dy, long long P, int C,
                         const float* scale, const float* shift, const float* mean, const float* rstd,
                         const float* sum_dz, const float* sum_dzx, double count, int act, float* acc_dbeta,
                         float* acc_dgamma, float acc_scale, void* stream);
int gp_pair_to_f16(const void* hi, const void* lo, long long ld_in, void* out, long long ld_out, long long rows, int cols,
                   void* stream);
int gp_bn_apply_act_pair(const float* y, void* out_bf16, void* out_f16, long long P, int C, const float* scale,
                         const float* shift, int act, void* stream);
int gp_split_matrix(const float* src, void* dst, int R, int K, int Rpad, int ld, int width, long long s_r, long long s_k,
                    int perm, long long lo_off, void* stream);
int gp_split_conv_weight(const float* src, void* dst, int D0, int D1, int taps, int n_dim, void* stream);
int gp_bn_stats_f32(const float* y, long long P, int C, float* sum, float* sumsq, void* stream);
int gp_bn_apply_act_split(const float* y, void* out_hi, void* out_lo, long long P, int C, const float* scale,
                          const float* shift, int act, void* stream);
int gp_bn_bwd_reduce_f32(const void* da, const float* y, long long P, int C, const float* scale, const float* shift,
                         const float* mean, const float* rstd, int act, float* sum_dz, float* sum_dzx, void* stream);
int gp_bn_bwd_apply_f32(const void* da, const float* y, void* dy, long long P, int C, const float* scale,
                        const float* shift, const float* mean, const float* rstd, const float* sum_dz,
                        const float* sum_dzx, double count, int act, float* acc_dbeta, float* acc_dgamma,
                        float acc_scale, void* stream);
int gp_im2col_k4s2_split(const float* img, void* col_hi, void* col_lo, int NB, int ch, int Hi, int Wi, void* stream);
int gp_col2im_k4s2_f32(const float* col, const float* bias, float* img, int NB, int ch, int Hi, int Wi, int act,
                       void* stream);
int gp_head_fwd_split(const void* a_hi, const void* a_lo, const float* w, const float* bias, float* out, int NB, int HW,
                      int C, int O, long long s_o, long long s_c, long long s_hw, void* stream);
/* gp_head_fwd on features with a companion tensor of either format (GP_COMP_LO == gp_head_fwd_split) */
int gp_head_fwd_comp(const void* a, const void* a_comp, int comp_fmt, const float* w, const float* bias, float* out, int NB,
                     int HW, int C, int O, long long s_o, long long s_c, long long s_hw, void* stream);

/* ---- BlurPool2d(filt_size=3, pad_type='reflect', stride 1 | 2) of models/ops.py:7-47, the anti-aliasing filter of
 * models/dcgan_blur.py:41,116 (the networks main_dcgan.py:52-53 instantiates): reflection pad 1 + depth-wise 3x3
 * outer([1,2,1],[1,2,1])/16 on NHWC bf16 (NB, H, W, C) -> (NB, (H-1)/stride+1, (W-1)/stride+1, C); _bwd is its adjoint
 * (dout on the output grid -> din on the (H, W) grid). */
int gp_blur3x3_fwd(const void* in, const void* in_comp, void* out, void* out_comp, int comp_fmt, int NB, int H, int W,
                   int C, int stride, void* stream);
int gp_blur3x3_bwd(const void* dout, void* din, int NB, int H, int W, int C, int stride, void* stream);

/* ---- SyncBN over NVLink peer memory (the reference has no parallelism; batch-sharded data parallelism
 * needs global-batch BatchNorm statistics, 15 forward + 12 backward reductions of <= 8 KB per DCGAN-64 step).
 * One-shot all-reduce: every rank pushes its partial sums into a slot of every peer's symmetric buffer, waits on
 * sequence-numbered flags and adds the slots in rank order (bit-identical totals on all ranks); one launch, no NCCL.
 * gp_peer_t: bufs[r] = device pointer (valid in THIS process) to rank r's buffer of gp_peer_buffer_bytes() bytes,
 * zero-initialised before first use; epoch = this rank's call counter (uint32 in device memory, zero-initialised).
 * All ranks must issue the same sequence of gp_peer_* calls. n <= 4096 floats.
 * gp_bn_finalize_peer = that exchange on st = (sum[C] | sumsq[C]) fused with gp_bn_finalize. */
typedef struct {
  void* bufs[8];
  int32_t world, rank;
  void* epoch;
} gp_peer_t;
long long gp_peer_buffer_bytes(void);
int gp_peer_allreduce_sum(const gp_peer_t* peer, float* data, int n, void* stream);
int gp_bn_finalize_peer(const gp_peer_t* peer, float* st, double count, int C, float eps, float momentum,
                        const float* gamma, const float* beta, float* mean, float* rstd, float* scale, float* shift,
                        float* running_mean, float* running_var, long long* num_batches_tracked, void* stream);

/* ---- optimiser edge: Adam with the literals of main_dcgan.py:55-56 / main_sngan.py:55-56
 * (torch.optim.Adam, no weight decay, no amsgrad) over FLAT fp32 buffers — one launch per network:
 *   m += (g*grad_scale - m)*(1-beta1); v = beta2*v + (1-beta2)*(g*grad_scale)^2;
 *   p -= lr/(1-beta1^t) * m / (sqrt(v)/sqrt(1-beta2^t) + eps),  t = *step + 1
 * step: device scalar (fp32) = steps taken before this call (the caller increments it; graph-capturable);
 * grad_scale: 1/world_size of data-parallel averaging folded into the gradient read. n % 4 == 0, 16-byte aligned. */
int gp_adam_flat(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1, double beta2,
                 double eps, const float* step, double grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GPB200_H_ */
