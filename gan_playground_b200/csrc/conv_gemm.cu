// Host side of the tcgen05 implicit-GEMM convolution: tap tables, TMA tensor maps, tile-shape choice, launch.
#include "conv_gemm.cuh"

#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>

#include "common.h"

namespace gp {

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

static int gcd_int(int a, int b) {
  while (b) {
    int t = a % b;
    a = b;
    b = t;
  }
  return a;
}

// Factor a block of `total` pixels (a power of two) over the (N, H, W) grid so that every tile is a box.
static void factor_tile(int total, int H, int W, int* Nt, int* Ht, int* Wt) {
  *Wt = gcd_int(W, total);
  *Ht = gcd_int(H, total / *Wt);
  *Nt = total / (*Wt * *Ht);
}

// 4-D bf16 NHWC view (C, W, H, N) with explicit element strides; box = (64 channels, bw, bh, bn), 128B swizzle.
static int make_map_nhwc(CUtensorMap* m, const void* base, int C, int W, int H, int N, long long sW, long long sH,
                         long long sN, int bw, int bh, int bn) {
  auto fn = get_encode_fn();
  if (fn == nullptr) return set_error(GP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sN * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 3; ++i)
    if (strides[i] % 16 != 0) return set_error(GP_ERR_INVALID, "TMA stride %d not a multiple of 16 bytes", i);
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(GP_ERR_INVALID, "TMA base not 16B aligned");
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(GP_ERR_CUDA, "cuTensorMapEncodeTiled(4d) failed: %d (C=%d W=%d H=%d N=%d box=%d,%d,%d)", (int)r,
                     C, W, H, N, bw, bh, bn);
  return GP_OK;
}

static int make_map_2d(CUtensorMap* m, const void* base, long long inner, long long outer, int box_outer) {
  auto fn = get_encode_fn();
  if (fn == nullptr) return set_error(GP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)inner * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  if (strides[0] % 16 != 0) return set_error(GP_ERR_INVALID, "packed weight row (%lld) not a multiple of 8", inner);
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(GP_ERR_INVALID, "TMA base not 16B aligned");
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(GP_ERR_CUDA, "cuTensorMapEncodeTiled(2d) failed: %d", (int)r);
  return GP_OK;
}

// Strided (k4 s2 p1) gather: input row ih = 2*oh - 1 + kh  ==  parity r = (kh+1)&1, half-row oh + floor((kh-1)/2).
static void k4s2_tap(int kh, int* parity, int* d) {
  *parity = (kh + 1) & 1;
  *d = (kh == 0) ? -1 : ((kh == 3) ? 1 : 0);
}

static int pick_bn(int n) {
  if (n >= 256 && n % 256 == 0) return 256;
  if (n >= 128) return 128;
  return 64;
}

// Experiment hook: GP_TILE_FWD / GP_TILE_WGRAD = "<BN>x<MT>" (e.g. "256x2") overrides the tile-shape heuristics.
static bool tile_override(const char* var, int* bn, int* mt) {
  const char* e = getenv(var);
  int b = 0, m = 0;
  if (e == nullptr || sscanf(e, "%dx%d", &b, &m) != 2) return false;
  if ((b != 64 && b != 128 && b != 256) || (m != 1 && m != 2)) return false;
  *bn = b;
  *mt = m;
  return true;
}

template <int MODE, int BNV, int MTV, bool X3V = false>
static int launch_cfg(const ConvGemmParams& prm, int grid, cudaStream_t st) {
  auto kfn = conv_gemm_kernel<MODE, BNV, MTV, X3V>;
  static bool attr_set = false;  // one flag per template instantiation
  if (!attr_set) {
    GP_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BNV, MTV, X3V>::kSmemBytes));
    attr_set = true;
  }
  gp::launch_pdl(kfn, grid, kNumThreads, GemmCfg<BNV, MTV, X3V>::kSmemBytes, st, prm);
  GP_CHECK_LAUNCH();
  return GP_OK;
}

template <int MODE>
static int launch(const ConvGemmParams& prm, int bn, int mt, int grid, cudaStream_t st) {
  if (grid <= 0) return GP_OK;
  if (bn == 256) return mt == 2 ? launch_cfg<MODE, 256, 2>(prm, grid, st) : launch_cfg<MODE, 256, 1>(prm, grid, st);
  if (bn == 128) return mt == 2 ? launch_cfg<MODE, 128, 2>(prm, grid, st) : launch_cfg<MODE, 128, 1>(prm, grid, st);
  return mt == 2 ? launch_cfg<MODE, 64, 2>(prm, grid, st) : launch_cfg<MODE, 64, 1>(prm, grid, st);
}

// CTA-pair forward (tcgen05 cta_group::2): 2 x (128 x 256) tile per cluster of two CTAs (conv_gemm.cuh, GemmCfg C2)
template <int BNV, bool X3V>
static int launch_fwd_pair(const ConvGemmParams& prm, int pair_tiles, cudaStream_t st) {
  auto kfn = conv_gemm_kernel<MODE_FWD, BNV, 1, X3V, true>;
  using Cfg = GemmCfg<BNV, 1, X3V, true>;
  static bool attr_set = false;
  if (!attr_set) {
    GP_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  int pairs = num_sms() / 2;
  if (pair_tiles < pairs) pairs = pair_tiles;
  if (pairs <= 0) return GP_OK;
  GP_CHECK_CUDA(gp::launch_attrs(kfn, dim3(2 * pairs, 1, 1), dim3(kNumThreads, 1, 1), Cfg::kSmemBytes, st, gp::pdl_enabled(),
                                 2u, prm));
  gp::count_launch();
  return GP_OK;
}

// CTA-pair wgrad: a pair owns a 256 x 256 dW tile (128 rows per CTA), each CTA stages its own 128 dense channels and half
// of the gathered columns: 32 KB per stage instead of 64 KB (6 stages instead of 3) and two accumulator buffers per CTA
// where the lone 256 x 256 tile has one.
static int launch_wgrad_pair(const ConvGemmParams& prm, int grid, cudaStream_t st) {
  auto kfn = conv_gemm_kernel<MODE_WGRAD, 256, 1, false, true>;
  using Cfg = GemmCfg<256, 1, false, true>;
  static bool attr_set = false;
  if (!attr_set) {
    GP_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  grid &= ~1;
  if (grid < 2) grid = 2;
  GP_CHECK_CUDA(gp::launch_attrs(kfn, dim3(grid, 1, 1), dim3(kNumThreads, 1, 1), Cfg::kSmemBytes, st, gp::pdl_enabled(), 2u,
                                 prm));
  gp::count_launch();
  return GP_OK;
}

static bool wide_flat_enabled() {
  const char* e = getenv("GP_WGRAD_WIDE_FLAT");
  return e == nullptr || e[0] != '0';
}

static bool wgrad_pair_enabled() {
  const char* e = getenv("GP_WGRAD_2CTA");
  return e != nullptr && e[0] == '1';   // off until validated
}

// GP_FWD_2CTA=0 disables the CTA-pair variant (1 = on, the default once a layer has enough 128-row tiles)
static bool pair_enabled() {
  const char* e = getenv("GP_FWD_2CTA");
  return e == nullptr || e[0] != '0';
}

// bf16x3 forward: hi and lo tiles of both operands per stage (no 256x256 variant: its stage would be 128 KB)
static int launch_fwd_x3(const ConvGemmParams& prm, int bn, int mt, int grid, cudaStream_t st) {
  if (grid <= 0) return GP_OK;
  if (bn == 256) return launch_cfg<MODE_FWD, 256, 1, true>(prm, grid, st);
  if (bn == 128) return mt == 2 ? launch_cfg<MODE_FWD, 128, 2, true>(prm, grid, st) : launch_cfg<MODE_FWD, 128, 1, true>(prm, grid, st);
  return mt == 2 ? launch_cfg<MODE_FWD, 64, 2, true>(prm, grid, st) : launch_cfg<MODE_FWD, 64, 1, true>(prm, grid, st);
}

// Tile shape of a forward / dgrad call: the candidate (BN, MT) with the smallest estimated time = waves x tile MACs x
// relative cost per MAC. Costs are measured ratios on B200 (tools/prof_gemm.py): 128x256 is the reference; 256x256 (no
// epilogue overlap) wins only for long K loops; narrower tiles move more operand bytes per FLOP but fill the 148 SMs at
// small batch.
static void choose_fwd_tile(const gp_conv_fwd_t* a, int* bn_out, int* mt_out) {
  int bn = 0, mt_sub = 1;
  {
    const long long small_px = (a->kind == GP_KIND_CONV_K4S2) ? (long long)a->NB * a->Hout * a->Wout
                                                              : (long long)a->NB * a->Hin * a->Win;
    const int phases = a->kind == GP_KIND_CONVT_K4S2 ? 4 : 1;
    const int taps_tile = a->kind == GP_KIND_CONV_K4S2 ? 16 : (a->kind == GP_KIND_CONVT_K4S2 ? 4 : (a->kind == GP_KIND_CONV_K3S1 ? 9 : 1));
    const long long ksteps = (long long)taps_tile * ((a->Cin + kBlockK - 1) / kBlockK) * (a->in_lo != nullptr ? 3 : 1);
    static const struct { int bn, mt; double cost; } cand[] = {
        {256, 1, 1.00}, {256, 2, 0.92}, {128, 2, 1.15}, {128, 1, 1.35}, {64, 2, 1.7}, {64, 1, 2.2}};
    double best = 0;
    for (const auto& c : cand) {
      if (c.bn > 64 && c.bn / 2 >= a->Nout) continue;            // more than half of the tile would be padding
      if (c.bn == 256 && c.mt == 2 && (ksteps < 128 || a->in_lo != nullptr)) continue;  // exposed epilogue not amortised; no bf16x3 variant
      const long long tiles = phases * ((small_px + c.mt * kBlockM - 1) / (c.mt * kBlockM)) * ((a->Nout + c.bn - 1) / c.bn);
      const long long waves = (tiles + num_sms() - 1) / num_sms();
      // short K loops: the per-tile epilogue and pipeline fill are not hidden; charge them as extra K steps
      const double t = (double)waves * c.mt * c.bn * c.cost * (double)(ksteps + 6);
      if (bn == 0 || t < best) {
        best = t;
        bn = c.bn;
        mt_sub = c.mt;
      }
    }
    tile_override("GP_TILE_FWD", &bn, &mt_sub);
  }
  *bn_out = bn;
  *mt_out = mt_sub;
}

}  // namespace gp

using namespace gp;

extern "C" int gp_conv_fwd(const gp_conv_fwd_t* a, void* stream) {
  GP_REQUIRE(a != nullptr && a->in && a->w && (a->out || a->out_f32), "gp_conv_fwd: null pointer");
  GP_REQUIRE(a->NB > 0 && a->Cin > 0 && a->Nout > 0, "gp_conv_fwd: empty problem");
  GP_REQUIRE(a->Cin % 8 == 0, "gp_conv_fwd: Cin=%d must be a multiple of 8 (16-byte TMA rows)", a->Cin);
  GP_REQUIRE(a->Nout % 8 == 0, "gp_conv_fwd: Nout=%d must be a multiple of 8", a->Nout);
  GP_REQUIRE(a->act != GP_ACT_TANH, "gp_conv_fwd: tanh is fused in gp_col2im_k4s2, not in the GEMM epilogue");
  GP_REQUIRE(a->col_sum == nullptr || a->Nout <= kMaxStatCols, "gp_conv_fwd: fused statistics support Nout <= %d", kMaxStatCols);
  ConvGemmParams prm;
  memset(&prm, 0, sizeof(prm));
  const int Cin = a->Cin;
  int ntaps_total = 0;
  int rc;
  int bn = 0, mt_sub = 1;
  choose_fwd_tile(a, &bn, &mt_sub);
  // CTA pair (cta_group::2): for 256-wide column tiles with at least one full wave of pair tiles. Each CTA of the pair
  // works on its own 128-pixel tile, so the pixel tiling below is the MT = 1 one.
  bool use_pair = false;
  long long pair_tiles = 0;
  // (128-wide column tiles gain nothing from pairing — a lone CTA's 256 x 128 tile already moves the same bytes per FLOP;
  // measured: G block-2 fprop 333 us paired vs 255 us — so only 256-wide tiles are paired; GP_FWD_2CTA=128 forces them)
  const char* pe = getenv("GP_FWD_2CTA");
  const bool pair128 = pe != nullptr && pe[0] == '1' && pe[1] == '2';
  if ((bn == 256 || (bn == 128 && pair128)) && pair_enabled() && a->Nout % bn == 0) {
    const int Hs = a->kind == GP_KIND_CONV_K4S2 ? a->Hout : a->Hin, Ws = a->kind == GP_KIND_CONV_K4S2 ? a->Wout : a->Win;
    int Nt1, Ht1, Wt1;
    factor_tile(kBlockM, Hs, Ws, &Nt1, &Ht1, &Wt1);
    const long long m128 = (long long)((a->NB + Nt1 - 1) / Nt1) * (Hs / Ht1) * (Ws / Wt1);
    pair_tiles = (a->kind == GP_KIND_CONVT_K4S2 ? 4 : 1) * ((m128 + 1) / 2) * (a->Nout / bn);
    use_pair = pair_tiles >= num_sms() / 2;
    if (use_pair) mt_sub = 1;
  }
  const int tile_px = mt_sub * kBlockM;
  const int n_halves = a->in_lo != nullptr ? 2 : 1;  // bf16x3: hi and lo halves of the activation operand
  const long long inW = Cin, inH = (long long)a->Win * Cin, inN = (long long)a->Hin * a->Win * Cin;
  switch (a->kind) {
    case GP_KIND_CONV_K4S2: {
      GP_REQUIRE(a->Hin == 2 * a->Hout && a->Win == 2 * a->Wout, "gp_conv_fwd: k4s2 needs Hin=2*Hout");
      prm.Hs = a->Hout;
      prm.Ws = a->Wout;
      prm.n_phases = 1;
      prm.taps_per_phase = 16;
      factor_tile(tile_px, prm.Hs, prm.Ws, &prm.Nt, &prm.Ht, &prm.Wt);
      for (int half = 0; half < n_halves; ++half)
        for (int r = 0; r < 2; ++r)
          for (int s = 0; s < 2; ++s) {
            const __nv_bfloat16* base =
                static_cast<const __nv_bfloat16*>(half ? a->in_lo : a->in) + ((long long)r * a->Win + s) * Cin;
            rc = make_map_nhwc(&prm.map_g[half * 4 + r * 2 + s], base, Cin, a->Win / 2, a->Hin / 2, a->NB, 2 * inW,
                               2 * inH, inN, prm.Wt, prm.Ht, prm.Nt);
            if (rc) return rc;
          }
      for (int kh = 0; kh < 4; ++kh)
        for (int kw = 0; kw < 4; ++kw) {
          int r, dh, s, dw;
          k4s2_tap(kh, &r, &dh);
          k4s2_tap(kw, &s, &dw);
          Tap& t = prm.taps[ntaps_total++];
          t.map = (int8_t)(r * 2 + s);
          t.dh = (int8_t)dh;
          t.dw = (int8_t)dw;
          t.koff = (kh * 4 + kw) * Cin;
        }
      prm.out_sN = (long long)a->Hout * a->Wout * a->Nout;
      prm.out_sH = (long long)a->Wout * a->Nout;
      prm.out_sW = a->Nout;
      break;
    }
    case GP_KIND_CONVT_K4S2: {
      GP_REQUIRE(a->Hout == 2 * a->Hin && a->Wout == 2 * a->Win, "gp_conv_fwd: convT k4s2 needs Hout=2*Hin");
      prm.Hs = a->Hin;
      prm.Ws = a->Win;
      prm.n_phases = 4;
      prm.taps_per_phase = 4;
      factor_tile(tile_px, prm.Hs, prm.Ws, &prm.Nt, &prm.Ht, &prm.Wt);
      for (int half = 0; half < n_halves; ++half) {
        rc = make_map_nhwc(&prm.map_g[half * 4], half ? a->in_lo : a->in, Cin, a->Win, a->Hin, a->NB, inW, inH, inN,
                           prm.Wt, prm.Ht, prm.Nt);
        if (rc) return rc;
      }
      // output row oh = 2*ih - 1 + kh. Phase ph = oh & 1: ph=0 -> (kh=1, ih=i), (kh=3, ih=i-1); ph=1 -> (kh=0, ih=i+1), (kh=2, ih=i)
      static const int kk[2][2] = {{1, 3}, {0, 2}};
      static const int dd[2][2] = {{0, -1}, {1, 0}};
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
          for (int ta = 0; ta < 2; ++ta)
            for (int tb = 0; tb < 2; ++tb) {
              Tap& t = prm.taps[ntaps_total++];
              t.map = 0;
              t.dh = (int8_t)dd[ph][ta];
              t.dw = (int8_t)dd[pw][tb];
              t.koff = (kk[ph][ta] * 4 + kk[pw][tb]) * Cin;
            }
          prm.phase_off[ph * 2 + pw] = ((long long)ph * a->Wout + pw) * a->Nout;
        }
      prm.out_sN = (long long)a->Hout * a->Wout * a->Nout;
      prm.out_sH = 2LL * a->Wout * a->Nout;
      prm.out_sW = 2LL * a->Nout;
      break;
    }
    case GP_KIND_CONV_K3S1:
    case GP_KIND_CONV_K1S1: {
      GP_REQUIRE(a->Hin == a->Hout && a->Win == a->Wout, "gp_conv_fwd: stride-1 conv needs Hin=Hout");
      const int k = (a->kind == GP_KIND_CONV_K3S1) ? 3 : 1;
      prm.Hs = a->Hin;
      prm.Ws = a->Win;
      prm.n_phases = 1;
      prm.taps_per_phase = k * k;
      factor_tile(tile_px, prm.Hs, prm.Ws, &prm.Nt, &prm.Ht, &prm.Wt);
      for (int half = 0; half < n_halves; ++half) {
        rc = make_map_nhwc(&prm.map_g[half * 4], half ? a->in_lo : a->in, Cin, a->Win, a->Hin, a->NB, inW, inH, inN,
                           prm.Wt, prm.Ht, prm.Nt);
        if (rc) return rc;
      }
      for (int kh = 0; kh < k; ++kh)
        for (int kw = 0; kw < k; ++kw) {
          Tap& t = prm.taps[ntaps_total++];
          t.map = 0;
          t.dh = (int8_t)(kh - k / 2);
          t.dw = (int8_t)(kw - k / 2);
          t.koff = (kh * k + kw) * Cin;
        }
      prm.out_sN = (long long)a->Hout * a->Wout * a->Nout;
      prm.out_sH = (long long)a->Wout * a->Nout;
      prm.out_sW = a->Nout;
      break;
    }
    default:
      return set_error(GP_ERR_UNSUPPORTED, "gp_conv_fwd: unknown kind %d", a->kind);
  }
  const long long ktot = (long long)ntaps_total * Cin;  // packed row = every tap of the kernel window
  if (n_halves == 2) {
    // bf16x3 (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo): the lo activation maps are map_g[4..7], the lo weights are the second
    // half [ktot, 2*ktot) of each packed row; the kernel loads hi and lo tiles of both operands into one stage
    prm.lo_koff = (int)ktot;
  }
  rc = make_map_2d(&prm.map_w, a->w, n_halves * ktot, a->Nout, use_pair ? bn / 2 : bn);  // a pair CTA loads half the columns
  if (rc) return rc;
  prm.map_d = prm.map_g[0];
  prm.NB = a->NB;
  prm.C = Cin;
  prm.N = a->Nout;
  prm.out = static_cast<__nv_bfloat16*>(a->out);
  prm.out_lo = static_cast<__nv_bfloat16*>(a->out_lo);
  prm.out_f32 = a->out_f32;
  prm.residual = static_cast<const __nv_bfloat16*>(a->residual);
  prm.bias = a->bias;
  prm.act_slope = a->act == GP_ACT_RELU ? 0.f : (a->act == GP_ACT_LRELU ? 0.2f : 1.f);
  prm.col_sum = a->col_sum;
  prm.col_sumsq = a->col_sumsq;
  GP_REQUIRE((a->flags & ~(GP_CONV_IN_F16 | GP_CONV_LO_F16 | GP_CONV_RES_F16 | GP_CONV_OUT_F16)) == 0,
             "gp_conv_fwd: unknown flags 0x%x", a->flags);
  GP_REQUIRE(!(a->flags & GP_CONV_OUT_F16) || (a->out != nullptr && a->out_lo == nullptr && a->out_f32 == nullptr),
             "gp_conv_fwd: GP_CONV_OUT_F16 writes a single fp16 tensor to `out`");
  GP_REQUIRE(!(a->flags & GP_CONV_RES_F16) || a->residual != nullptr, "gp_conv_fwd: GP_CONV_RES_F16 needs a residual");
  GP_REQUIRE(!(a->flags & GP_CONV_IN_F16) || a->in_lo == nullptr, "gp_conv_fwd: fp16 operands run as one MMA (in_lo must be NULL)");
  GP_REQUIRE(!(a->flags & GP_CONV_LO_F16) || (a->out_lo != nullptr && a->out != nullptr), "gp_conv_fwd: GP_CONV_LO_F16 needs out and out_lo");
  prm.fmt_flags = ((a->flags & GP_CONV_IN_F16) ? kFmtInF16 : 0) | ((a->flags & GP_CONV_LO_F16) ? kFmtLoF16 : 0) |
                  ((a->flags & GP_CONV_RES_F16) ? kFmtResF16 : 0) | ((a->flags & GP_CONV_OUT_F16) ? kFmtOutF16 : 0);
  GP_REQUIRE((a->col_sum == nullptr) == (a->col_sumsq == nullptr), "gp_conv_fwd: col_sum and col_sumsq go together");
  GP_REQUIRE(a->bwd_mode >= GP_CONV_BWD_NONE && a->bwd_mode <= GP_CONV_BWD_BN_BF16, "gp_conv_fwd: unknown bwd_mode %d", a->bwd_mode);
  if (a->bwd_mode != GP_CONV_BWD_NONE) {
    // this GEMM is a data gradient: plain bf16 output, nothing else fused
    GP_REQUIRE(a->bwd_src != nullptr && a->out != nullptr && a->out_lo == nullptr && a->out_f32 == nullptr &&
                   a->bias == nullptr && a->residual == nullptr && a->act == GP_ACT_NONE && a->in_lo == nullptr &&
                   (a->flags & (GP_CONV_LO_F16 | GP_CONV_OUT_F16)) == 0,
               "gp_conv_fwd: bwd_mode %d needs bwd_src and a plain bf16 output (no bias / residual / activation / pair)", a->bwd_mode);
    GP_REQUIRE(a->bwd_slope >= 0.f && a->bwd_slope <= 1.f, "gp_conv_fwd: bwd_slope %g outside [0, 1]", (double)a->bwd_slope);
    if (a->bwd_mode == GP_CONV_BWD_MASK)
      GP_REQUIRE(a->col_sum == nullptr, "gp_conv_fwd: GP_CONV_BWD_MASK does not produce statistics");
    else
      GP_REQUIRE(a->col_sum != nullptr && a->bwd_fin != nullptr, "gp_conv_fwd: the fused BatchNorm-backward reduction needs col_sum / col_sumsq / bwd_fin");
    prm.bwd_src = a->bwd_src;
    prm.bwd_fin = a->bwd_fin;
    prm.bwd_slope = a->bwd_slope;
    prm.bwd_mode = a->bwd_mode;
  }
  const int mtiles = ((a->NB + prm.Nt - 1) / prm.Nt) * (prm.Hs / prm.Ht) * (prm.Ws / prm.Wt);
  const int ntn = (a->Nout + bn - 1) / bn;
  const int num_tiles = prm.n_phases * mtiles * ntn;
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  if (use_pair) {
    if (bn == 256)
      return n_halves == 2 ? launch_fwd_pair<256, true>(prm, (int)pair_tiles, as_stream(stream))
                           : launch_fwd_pair<256, false>(prm, (int)pair_tiles, as_stream(stream));
    return n_halves == 2 ? launch_fwd_pair<128, true>(prm, (int)pair_tiles, as_stream(stream))
                         : launch_fwd_pair<128, false>(prm, (int)pair_tiles, as_stream(stream));
  }
  if (n_halves == 2) return launch_fwd_x3(prm, bn, mt_sub, grid, as_stream(stream));
  return launch<MODE_FWD>(prm, bn, mt_sub, grid, as_stream(stream));
}

extern "C" int gp_conv_fwd_plan(const gp_conv_fwd_t* a, int* bn, int* mt, int* tiles) {
  GP_REQUIRE(a != nullptr && bn && mt && tiles && a->NB > 0 && a->Cin > 0 && a->Nout > 0, "gp_conv_fwd_plan: bad arguments");
  choose_fwd_tile(a, bn, mt);
  const long long small_px = (a->kind == GP_KIND_CONV_K4S2) ? (long long)a->NB * a->Hout * a->Wout : (long long)a->NB * a->Hin * a->Win;
  const int phases = a->kind == GP_KIND_CONVT_K4S2 ? 4 : 1;
  *tiles = (int)(phases * ((small_px + *mt * kBlockM - 1) / (*mt * kBlockM)) * ((a->Nout + *bn - 1) / *bn));
  return GP_OK;
}

extern "C" int gp_conv_wgrad(const gp_conv_wgrad_t* a, void* stream) {
  GP_REQUIRE(a != nullptr && a->dense && a->gath && a->dw, "gp_conv_wgrad: null pointer");
  GP_REQUIRE(a->NB > 0 && a->Cd > 0 && a->Cg > 0, "gp_conv_wgrad: empty problem");
  GP_REQUIRE(a->Cd % 8 == 0 && a->Cg % 8 == 0, "gp_conv_wgrad: channel counts must be multiples of 8");
  ConvGemmParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.NB = a->NB;
  prm.Hs = a->Hs;
  prm.Ws = a->Ws;
  factor_tile(kBlockK, prm.Hs, prm.Ws, &prm.Nt, &prm.Ht, &prm.Wt);
  int rc = make_map_nhwc(&prm.map_d, a->dense, a->Cd, a->Ws, a->Hs, a->NB, a->Cd, (long long)a->Ws * a->Cd,
                         (long long)a->Hs * a->Ws * a->Cd, prm.Wt, prm.Ht, prm.Nt);
  if (rc) return rc;
  const int Cg = a->Cg;
  const long long gW = Cg, gH = (long long)a->Wg * Cg, gN = (long long)a->Hg * a->Wg * Cg;
  int ntaps = 0;
  switch (a->kind) {
    case GP_KIND_CONV_K4S2: {
      GP_REQUIRE(a->Hg == 2 * a->Hs && a->Wg == 2 * a->Ws, "gp_conv_wgrad: k4s2 needs Hg=2*Hs");
      for (int r = 0; r < 2; ++r)
        for (int s = 0; s < 2; ++s) {
          const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(a->gath) + ((long long)r * a->Wg + s) * Cg;
          rc = make_map_nhwc(&prm.map_g[r * 2 + s], base, Cg, a->Wg / 2, a->Hg / 2, a->NB, 2 * gW, 2 * gH, gN, prm.Wt,
                             prm.Ht, prm.Nt);
          if (rc) return rc;
        }
      for (int kh = 0; kh < 4; ++kh)
        for (int kw = 0; kw < 4; ++kw) {
          int r, dh, s, dw;
          k4s2_tap(kh, &r, &dh);
          k4s2_tap(kw, &s, &dw);
          Tap& t = prm.taps[ntaps++];
          t.map = (int8_t)(r * 2 + s);
          t.dh = (int8_t)dh;
          t.dw = (int8_t)dw;
          t.koff = (kh * 4 + kw) * Cg;
        }
      break;
    }
    case GP_KIND_CONV_K3S1:
    case GP_KIND_CONV_K1S1: {
      GP_REQUIRE(a->Hg == a->Hs && a->Wg == a->Ws, "gp_conv_wgrad: stride-1 needs Hg=Hs");
      const int k = (a->kind == GP_KIND_CONV_K3S1) ? 3 : 1;
      rc = make_map_nhwc(&prm.map_g[0], a->gath, Cg, a->Wg, a->Hg, a->NB, gW, gH, gN, prm.Wt, prm.Ht, prm.Nt);
      if (rc) return rc;
      for (int i = 1; i < 4; ++i) prm.map_g[i] = prm.map_g[0];
      for (int kh = 0; kh < k; ++kh)
        for (int kw = 0; kw < k; ++kw) {
          Tap& t = prm.taps[ntaps++];
          t.map = 0;
          t.dh = (int8_t)(kh - k / 2);
          t.dw = (int8_t)(kw - k / 2);
          t.koff = (kh * k + kw) * Cg;
        }
      break;
    }
    default:
      return set_error(GP_ERR_UNSUPPORTED, "gp_conv_wgrad: unsupported kind %d", a->kind);
  }
  prm.map_w = prm.map_d;
  prm.n_phases = 1;
  prm.taps_per_phase = ntaps;
  prm.M = a->Cd;
  prm.N = Cg;
  prm.dw = a->dw;
  prm.ldw = ntaps * Cg;
  // narrow gathered operand with several taps: tile the flattened (tap, channel) columns so that a 256-wide tile
  // spans 256 / Cg taps (dW rows are [tap][channel], i.e. contiguous in that index)
  prm.wg_flat = (Cg % 64 == 0 && ntaps > 1 && Cg < 256) ? 1 : 0;
  int bn = pick_bn(prm.wg_flat ? ntaps * Cg : Cg);
  // flattened 3x3 columns (9 * Cg) are never a multiple of 256: a partly empty last 256-wide tile (the producer clamps
  // its taps, the epilogue masks its columns) still moves fewer operand bytes per FLOP than 128-wide tiles when the
  // padding stays under ~15 % (Cg = 128: 1152 -> 1280)
  if (prm.wg_flat && bn == 128 && wide_flat_enabled()) {
    const int n = ntaps * Cg, padded = (n + 255) / 256 * 256;
    if (n >= 512 && (padded - n) * 100 <= 15 * n) bn = 256;
  }
  // two M=128 sub-tiles share the gathered tile whenever dW has >= 256 rows: the wgrad mainloop is bound by operand
  // traffic from L2 (48 KB per 128x256x64 MMA block), a 256-row tile moves 1.5x fewer bytes per FLOP
  int mt_sub = a->Cd >= 2 * kBlockM ? 2 : 1;
  tile_override("GP_TILE_WGRAD", &bn, &mt_sub);
  const int mtiles = (a->Cd + mt_sub * kBlockM - 1) / (mt_sub * kBlockM);
  const int ntn = ((prm.wg_flat ? ntaps * Cg : Cg) + bn - 1) / bn;
  const int base_tiles = mtiles * (prm.wg_flat ? 1 : ntaps) * ntn;
  prm.kblocks_total = ((a->NB + prm.Nt - 1) / prm.Nt) * (prm.Hs / prm.Ht) * (prm.Ws / prm.Wt);
  // split-K over pixel blocks. The hybrid stream-K schedule balances any tile count, so splits are only there to keep
  // the pixel range swept by one wave of CTAs small enough for its operand tiles to be shared in L2 (~192 K blocks
  // per split); every extra split costs one more fp32 atomic pass over dW.
  int splits = (prm.kblocks_total + 191) / 192;
  const int max_splits = prm.kblocks_total / 8 > 0 ? prm.kblocks_total / 8 : 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  // drop empty trailing splits
  const int per = (prm.kblocks_total + splits - 1) / splits;
  splits = (prm.kblocks_total + per - 1) / per;
  prm.splits = splits;
  // Whole tiles round-robin, the K ranges of the last partial wave cut evenly over all CTAs (WorkPlan in conv_gemm.cuh):
  // every SM gets the same number of K steps, so the grid is always the full machine unless the problem is tiny.
  const long long total_ksteps = (long long)base_tiles * prm.kblocks_total;
  int grid = num_sms();
  if (total_ksteps < 4LL * grid) grid = (int)((total_ksteps + 3) / 4);
  // Less than one wave of tiles, but most of one (e.g. 128 tiles of 256 x 256 for dW[1024][16][512] on 148 SMs): cutting
  // every tile's K range over all CTAs would add each tile to dW twice (two partial accumulators per CTA, two atomic
  // passes over a 33 MB gradient that a short K loop cannot hide). One whole tile per CTA on fewer CTAs moves half the
  // atomic traffic for 1 / 0.75 more K steps at worst. GP_WGRAD_WHOLE=0 keeps the even cut.
  {
    static const bool whole = [] { const char* e = getenv("GP_WGRAD_WHOLE"); return e == nullptr || e[0] != '0'; }();
    const int T = splits * base_tiles;
    if (whole && grid == num_sms() && T < grid && 4 * T >= 3 * grid) grid = T;
  }
  if (bn == 256 && mt_sub == 2 && grid == num_sms() && wgrad_pair_enabled())
    return launch_wgrad_pair(prm, grid, as_stream(stream));   // same 256 x 256 tiles, one per CTA pair
  return launch<MODE_WGRAD>(prm, bn, mt_sub, grid, as_stream(stream));
}
