// extern "C" entry points of the conv / transposed-conv family: tap-table construction + dispatch.
#include <stdlib.h>
#include <string.h>

#include "uad_conv.cuh"

static inline int pad_lo_of(int k) { return (k - 2) / 2; }   // TF SAME, stride 2, even input: total pad k-2, low half first

// all k*k taps, stride-2 gather (Form F / Form W)
static void taps_full(TapSet* ts, int k) {
  const int lo = pad_lo_of(k);
  ts->n = k * k;
  ts->oh0 = ts->ow0 = 0;
  for (int kh = 0; kh < k; ++kh)
    for (int kw = 0; kw < k; ++kw) {
      int t = kh * k + kw;
      ts->dh[t] = (signed char)(kh - lo);
      ts->dw[t] = (signed char)(kw - lo);
      ts->wt[t] = (signed char)t;
    }
}

// output-parity class (p, q) of the transposed form: fine pixel (2r+p, 2s+q) <- coarse pixel (r+dh, s+dw)
static void taps_parity(TapSet* ts, int k, int p, int q) {
  const int lo = pad_lo_of(k);
  ts->n = 0;
  ts->oh0 = p;
  ts->ow0 = q;
  for (int kh = 0; kh < k; ++kh) {
    if ((p - kh + lo) & 1) continue;
    for (int kw = 0; kw < k; ++kw) {
      if ((q - kw + lo) & 1) continue;
      int i = ts->n++;
      ts->dh[i] = (signed char)((p - kh + lo) / 2);
      ts->dw[i] = (signed char)((q - kw + lo) / 2);
      ts->wt[i] = (signed char)(kh * k + kw);
    }
  }
}

static int check_geom(const char* op, int B, int H, int W, int Cin, int Cout, int k) {
  UAD_REQUIRE(B > 0 && uad_is_pow2(H) && uad_is_pow2(W) && H >= 2 && W >= 2, "%s: H=%d W=%d must be powers of two", op, H, W);
  UAD_REQUIRE(k >= 2 && k <= 5, "%s: ksize=%d unsupported", op, k);
  UAD_REQUIRE(Cin > 0 && Cout > 0, "%s: bad channels", op);
  UAD_REQUIRE((long long)B * H * W * 4 < (1LL << 31), "%s: B*H*W too large", op);
  return 0;
}

static bool want_tc(int math_mode) { return math_mode == UAD_MATH_TC_3XTF32 || math_mode == UAD_MATH_TC_1XTF32; }
// the modes that exist: exact-fp32 SIMT, fp32-accurate 3xTF32 tensor cores, and UAD_MATH_TC_1XTF32 - one tf32 MMA per K-step on
// operands rounded to nearest tf32 (fp32 storage, fp32 accumulation: the arithmetic class of a bf16 / TF32 training step, ~1e-3
// relative, NOT the 1e-4 parity mode).  The 1x form exists in conv_halo_ss / wgrad_ss; a shape only the first-generation kernels
// cover runs 3xTF32 in that mode too (MORE accurate than asked for, never less).  Any other value is an error, never an alias.
static int check_mode(const char* op, int math_mode) {
  UAD_REQUIRE(math_mode == UAD_MATH_FP32_SIMT || math_mode == UAD_MATH_TC_3XTF32 || math_mode == UAD_MATH_TC_1XTF32,
              "%s: math_mode %d is not built (UAD_MATH_FP32_SIMT, UAD_MATH_TC_3XTF32, UAD_MATH_TC_1XTF32 only)", op, math_mode);
  return 0;
}

// Form F / Form T on tensor cores: the halo-resident SS kernel (uad_conv_hs.cu) wherever the M-grid is at least 16 x 8, the
// converter-warp kernels (uad_conv_tc.cu) for the 8 x 8 grids at the bottleneck.  UAD_HS=0 (developer switch) forces the latter.
static int launch_gather_tensor(const GatherParams& p, int nclasses, int ksize, bool weights_transposed, const float* w_raw,
                                int math_mode, void* ws, size_t ws_bytes, cudaStream_t st) {
  static int use_hs = -1;
  if (use_hs < 0) { const char* e = getenv("UAD_HS"); use_hs = e ? atoi(e) : 1; }
  if (use_hs && ksize == 5 && uad_hs_gather_supported(p.Cin, p.N, p.lgMH, p.lgMW, nclasses))
    return uad_launch_gather_hs(p, nclasses, ksize, weights_transposed, w_raw, math_mode == UAD_MATH_TC_1XTF32, ws, ws_bytes, st);
  return uad_launch_gather_tc(p, nclasses, ksize, weights_transposed, w_raw, UAD_MATH_TC_3XTF32, ws, ws_bytes, st);
}

// Form W on tensor cores: the MN-major SS kernel (uad_conv_ws.cu) where the M-grid is wide enough for its pixel blocks, the
// converter-warp kernel (uad_conv_tc.cu) otherwise.  UAD_WGRAD_SS=0 (developer switch) forces the latter.
static int launch_wgrad_tensor(const WgradParams& p, float* dw, int accumulate, int math_mode, void* ws, size_t ws_bytes, cudaStream_t st) {
  static int use_ss = -1;
  if (use_ss < 0) { const char* e = getenv("UAD_WGRAD_SS"); use_ss = e ? atoi(e) : 1; }
  if (use_ss && uad_ws_wgrad_supported(p.Cg, p.Co, p.lgMH, p.lgMW)) return uad_launch_wgrad_ss(p, dw, accumulate, math_mode == UAD_MATH_TC_1XTF32, ws, ws_bytes, st);
  return uad_launch_wgrad_tc(p, dw, accumulate, ws, ws_bytes, st);
}

extern "C" int uad_conv_tc_supported(int op, int B, int H, int W, int Cin, int Cout, int ksize) {
  (void)B;
  if (ksize != 5) return 0;
  switch (op) {
    case UAD_OP_CONV_FWD:    return uad_tc_gather_supported(Cin, Cout, uad_ilog2(H / 2), uad_ilog2(W / 2));
    case UAD_OP_CONV_DGRAD:  return uad_tc_gather_supported(Cout, Cin, uad_ilog2(H / 2), uad_ilog2(W / 2));
    case UAD_OP_CONVT_FWD:   return uad_tc_gather_supported(Cin, Cout, uad_ilog2(H), uad_ilog2(W));
    case UAD_OP_CONVT_DGRAD: return uad_tc_gather_supported(Cout, Cin, uad_ilog2(H), uad_ilog2(W));
    case UAD_OP_CONV_WGRAD:  return Cin > 1 && uad_tc_wgrad_supported(Cin, Cout, uad_ilog2(H / 2), uad_ilog2(W / 2));
    case UAD_OP_CONVT_WGRAD: return uad_tc_wgrad_supported(Cout, Cin, uad_ilog2(H), uad_ilog2(W));
    default: return 0;
  }
}

extern "C" size_t uad_conv_workspace_bytes(int op, int B, int H, int W, int Cin, int Cout, int ksize, int math_mode) {
  (void)math_mode;
  const size_t wbytes = (size_t)ksize * ksize * Cin * Cout * sizeof(float);
  switch (op) {
    case UAD_OP_CONV_FWD:
    case UAD_OP_CONVT_DGRAD:
    case UAD_OP_CONV_DGRAD:
    case UAD_OP_CONVT_FWD: {
      size_t tc = uad_tc_gather_ws_bytes(ksize, Cin, Cout);
      return (tc > wbytes ? tc : wbytes) + 256;
    }
    case UAD_OP_CONV_WGRAD: {
      if (Cin == 1) return (size_t)uad_conv_c1_wgrad_blocks(B, H) * ksize * ksize * Cout * sizeof(float) + 256;
      int splits, chunk;
      uad_wgrad_plan(ksize * ksize * Cin, Cout, B * (H / 2) * (W / 2), &splits, &chunk);
      size_t simt = (size_t)splits * wbytes + 256;
      size_t tc = (ksize == 5 && uad_tc_wgrad_supported(Cin, Cout, uad_ilog2(H / 2), uad_ilog2(W / 2)))
                      ? uad_tc_wgrad_ws_bytes(Cin, Cout, B * (H / 2) * (W / 2)) : 0;
      if (ksize == 5 && Cin % 32 == 0 && uad_ws_wgrad_supported(Cin, Cout, uad_ilog2(H / 2), uad_ilog2(W / 2))) {
        const size_t ss = uad_ws_wgrad_ws_bytes(Cin, Cout, B, H / 2, W / 2);
        if (ss > tc) tc = ss;
      }
      return simt > tc ? simt : tc;
    }
    case UAD_OP_CONVT_WGRAD: {
      int splits, chunk;
      uad_wgrad_plan(ksize * ksize * Cout, Cin, B * H * W, &splits, &chunk);
      size_t simt = (size_t)splits * wbytes + 256;
      size_t tc = (ksize == 5 && uad_tc_wgrad_supported(Cout, Cin, uad_ilog2(H), uad_ilog2(W)))
                      ? uad_tc_wgrad_ws_bytes(Cout, Cin, B * H * W) : 0;
      if (ksize == 5 && Cout % 32 == 0 && uad_ws_wgrad_supported(Cout, Cin, uad_ilog2(H), uad_ilog2(W))) {
        const size_t ss = uad_ws_wgrad_ws_bytes(Cout, Cin, B, H, W);
        if (ss > tc) tc = ss;
      }
      return simt > tc ? simt : tc;
    }
    default: return 0;
  }
}

// -------------------------------------------------------------------------------------------- conv (strided)
extern "C" int uad_conv2d_fwd(const float* x, const float* w, const float* bias, const float* gamma, const float* beta,
                              float* z_out, float* a_out, int B, int H, int W, int Cin, int Cout, int ksize, int act,
                              float alpha, float bn_c, int math_mode, void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_geom("uad_conv2d_fwd", B, H, W, Cin, Cout, ksize)) return e;
  if (int e = check_mode("uad_conv2d_fwd", math_mode)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 1)
    return uad_launch_conv_c1_fwd(x, w, bias, gamma, beta, z_out, a_out, B, H, W, Cout, ksize, act, alpha, bn_c, st);
  GatherParams p;
  memset(&p, 0, sizeof(p));
  p.in = x; p.wmat = w; p.z_out = z_out; p.a_out = a_out; p.bias = bias; p.gamma = gamma; p.beta = beta;
  p.B = B; p.IH = H; p.IW = W; p.Cin = Cin;
  p.lgMH = uad_ilog2(H / 2); p.lgMW = uad_ilog2(W / 2); p.sh = 2;
  p.OH = H / 2; p.OW = W / 2; p.N = Cout; p.osh = 1;
  p.M = B << (p.lgMH + p.lgMW);
  p.act = act; p.alpha = alpha; p.bn_c = bn_c;
  taps_full(&p.taps[0], ksize);
  if (want_tc(math_mode) && uad_conv_tc_supported(UAD_OP_CONV_FWD, B, H, W, Cin, Cout, ksize))
    return launch_gather_tensor(p, 1, ksize, false, w, math_mode, ws, ws_bytes, st);
  return uad_launch_gather_simt(p, 1, st);
}

extern "C" int uad_conv2d_dgrad(const float* dz, const float* w, float* dx, int B, int H, int W, int Cin, int Cout,
                                int ksize, int math_mode, void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_geom("uad_conv2d_dgrad", B, H, W, Cin, Cout, ksize)) return e;
  if (int e = check_mode("uad_conv2d_dgrad", math_mode)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 1) return uad_launch_conv_c1_dgrad(dz, w, dx, B, H, W, Cout, ksize, st);
  GatherParams p;
  memset(&p, 0, sizeof(p));
  p.in = dz; p.a_out = dx;
  p.B = B; p.IH = H / 2; p.IW = W / 2; p.Cin = Cout;
  p.lgMH = uad_ilog2(H / 2); p.lgMW = uad_ilog2(W / 2); p.sh = 1;
  p.OH = H; p.OW = W; p.N = Cin; p.osh = 2;
  p.M = B << (p.lgMH + p.lgMW);
  p.act = UAD_ACT_NONE; p.alpha = 0.f; p.bn_c = 1.f;
  for (int c = 0; c < 4; ++c) taps_parity(&p.taps[c], ksize, c >> 1, c & 1);
  if (want_tc(math_mode) && uad_conv_tc_supported(UAD_OP_CONV_DGRAD, B, H, W, Cin, Cout, ksize))
    return launch_gather_tensor(p, 4, ksize, true, w, math_mode, ws, ws_bytes, st);
  // SIMT: needs wmat[t][Cout][Cin] = transpose of HWIO w[t][Cin][Cout]
  size_t need = (size_t)ksize * ksize * Cin * Cout * sizeof(float);
  UAD_REQUIRE(ws && ws_bytes >= need, "uad_conv2d_dgrad: workspace too small (%zu < %zu)", ws_bytes, need);
  if (int e = uad_launch_transpose_taps(w, (float*)ws, ksize * ksize, Cin, Cout, st)) return e;
  p.wmat = (const float*)ws;
  return uad_launch_gather_simt(p, 4, st);
}

extern "C" int uad_conv2d_wgrad(const float* x, const float* dz, float* dw, int B, int H, int W, int Cin, int Cout,
                                int ksize, int accumulate, int math_mode, void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_geom("uad_conv2d_wgrad", B, H, W, Cin, Cout, ksize)) return e;
  if (int e = check_mode("uad_conv2d_wgrad", math_mode)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 1) return uad_launch_conv_c1_wgrad(x, dz, dw, B, H, W, Cout, ksize, accumulate, ws, ws_bytes, st);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.g = x; p.o = dz;
  p.B = B; p.GH = H; p.GW = W; p.Cg = Cin;
  p.lgMH = uad_ilog2(H / 2); p.lgMW = uad_ilog2(W / 2); p.sh = 2; p.Co = Cout;
  p.Mp = ksize * ksize * Cin;
  p.P = B << (p.lgMH + p.lgMW);
  taps_full(&p.taps, ksize);
  if (want_tc(math_mode) && uad_conv_tc_supported(UAD_OP_CONV_WGRAD, B, H, W, Cin, Cout, ksize))
    return launch_wgrad_tensor(p, dw, accumulate, math_mode, ws, ws_bytes, st);
  return uad_launch_wgrad_simt(p, dw, accumulate, ws, ws_bytes, st);
}

// -------------------------------------------------------------------------------------------- transposed conv
extern "C" int uad_convT2d_fwd(const float* x, const float* w, const float* bias, const float* gamma, const float* beta,
                               float* z_out, float* a_out, int B, int H, int W, int Cin, int Cout, int ksize, int act,
                               float alpha, float bn_c, int math_mode, void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_geom("uad_convT2d_fwd", B, H, W, Cin, Cout, ksize)) return e;
  if (int e = check_mode("uad_convT2d_fwd", math_mode)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  GatherParams p;
  memset(&p, 0, sizeof(p));
  p.in = x; p.z_out = z_out; p.a_out = a_out; p.bias = bias; p.gamma = gamma; p.beta = beta;
  p.B = B; p.IH = H; p.IW = W; p.Cin = Cin;
  p.lgMH = uad_ilog2(H); p.lgMW = uad_ilog2(W); p.sh = 1;
  p.OH = 2 * H; p.OW = 2 * W; p.N = Cout; p.osh = 2;
  p.M = B << (p.lgMH + p.lgMW);
  p.act = act; p.alpha = alpha; p.bn_c = bn_c;
  for (int c = 0; c < 4; ++c) taps_parity(&p.taps[c], ksize, c >> 1, c & 1);
  if (want_tc(math_mode) && uad_conv_tc_supported(UAD_OP_CONVT_FWD, B, H, W, Cin, Cout, ksize))
    return launch_gather_tensor(p, 4, ksize, true, w, math_mode, ws, ws_bytes, st);
  // SIMT: needs wmat[t][Cin][Cout] = transpose of TF layout w[t][Cout][Cin]
  size_t need = (size_t)ksize * ksize * Cin * Cout * sizeof(float);
  UAD_REQUIRE(ws && ws_bytes >= need, "uad_convT2d_fwd: workspace too small (%zu < %zu)", ws_bytes, need);
  if (int e = uad_launch_transpose_taps(w, (float*)ws, ksize * ksize, Cout, Cin, st)) return e;
  p.wmat = (const float*)ws;
  return uad_launch_gather_simt(p, 4, st);
}

// transposed conv block + the 1x1 conv that follows it (Cout -> 1 channel), fused into the block's epilogue: the tensor-core path
// only, Cout = 32 (one thread of the epilogue holds a pixel's 32 channels).  Callers ask uad_convT2d_fwd_head_supported first.
extern "C" int uad_convT2d_fwd_head_supported(int B, int H, int W, int Cin, int Cout, int ksize, int math_mode) {
  static int use_hs = -1;
  if (use_hs < 0) { const char* e = getenv("UAD_HS"); use_hs = e ? atoi(e) : 1; }
  return use_hs && want_tc(math_mode) && ksize == 5 && Cout == 32 && uad_is_pow2(H) && uad_is_pow2(W) && B > 0 &&
         uad_hs_gather_supported(Cin, Cout, uad_ilog2(H), uad_ilog2(W), 4);
}

extern "C" int uad_convT2d_fwd_head(const float* x, const float* w, const float* bias, const float* gamma, const float* beta,
                                    float* a_out, const float* head_w, const float* head_b, float* head_out, int B, int H, int W,
                                    int Cin, int Cout, int ksize, int act, float alpha, float bn_c, int math_mode, void* ws,
                                    size_t ws_bytes, void* stream) {
  if (int e = check_geom("uad_convT2d_fwd_head", B, H, W, Cin, Cout, ksize)) return e;
  if (int e = check_mode("uad_convT2d_fwd_head", math_mode)) return e;
  UAD_REQUIRE(uad_convT2d_fwd_head_supported(B, H, W, Cin, Cout, ksize, math_mode), "uad_convT2d_fwd_head: unsupported shape / math mode");
  UAD_REQUIRE(a_out && head_w && head_b && head_out, "uad_convT2d_fwd_head: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  GatherParams p;
  memset(&p, 0, sizeof(p));
  p.in = x; p.a_out = a_out; p.bias = bias; p.gamma = gamma; p.beta = beta;
  p.B = B; p.IH = H; p.IW = W; p.Cin = Cin;
  p.lgMH = uad_ilog2(H); p.lgMW = uad_ilog2(W); p.sh = 1;
  p.OH = 2 * H; p.OW = 2 * W; p.N = Cout; p.osh = 2;
  p.M = B << (p.lgMH + p.lgMW);
  p.act = act; p.alpha = alpha; p.bn_c = bn_c;
  p.head_w = head_w; p.head_b = head_b; p.head_out = head_out;
  for (int c = 0; c < 4; ++c) taps_parity(&p.taps[c], ksize, c >> 1, c & 1);
  return uad_launch_gather_hs(p, 4, ksize, true, w, math_mode == UAD_MATH_TC_1XTF32, ws, ws_bytes, st);
}

extern "C" int uad_convT2d_dgrad(const float* dz, const float* w, float* dx, int B, int H, int W, int Cin, int Cout,
                                 int ksize, int math_mode, void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_geom("uad_convT2d_dgrad", B, H, W, Cin, Cout, ksize)) return e;
  if (int e = check_mode("uad_convT2d_dgrad", math_mode)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  GatherParams p;
  memset(&p, 0, sizeof(p));
  p.in = dz; p.wmat = w; p.a_out = dx;          // w[t][Cout][Cin] is already [t][gather-channel][N]
  p.B = B; p.IH = 2 * H; p.IW = 2 * W; p.Cin = Cout;
  p.lgMH = uad_ilog2(H); p.lgMW = uad_ilog2(W); p.sh = 2;
  p.OH = H; p.OW = W; p.N = Cin; p.osh = 1;
  p.M = B << (p.lgMH + p.lgMW);
  p.act = UAD_ACT_NONE; p.alpha = 0.f; p.bn_c = 1.f;
  taps_full(&p.taps[0], ksize);
  if (want_tc(math_mode) && uad_conv_tc_supported(UAD_OP_CONVT_DGRAD, B, H, W, Cin, Cout, ksize))
    return launch_gather_tensor(p, 1, ksize, false, w, math_mode, ws, ws_bytes, st);
  return uad_launch_gather_simt(p, 1, st);
}

extern "C" int uad_convT2d_wgrad(const float* x, const float* dz, float* dw, int B, int H, int W, int Cin, int Cout,
                                 int ksize, int accumulate, int math_mode, void* ws, size_t ws_bytes, void* stream) {
  if (int e = check_geom("uad_convT2d_wgrad", B, H, W, Cin, Cout, ksize)) return e;
  if (int e = check_mode("uad_convT2d_wgrad", math_mode)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.g = dz; p.o = x;
  p.B = B; p.GH = 2 * H; p.GW = 2 * W; p.Cg = Cout;
  p.lgMH = uad_ilog2(H); p.lgMW = uad_ilog2(W); p.sh = 2; p.Co = Cin;
  p.Mp = ksize * ksize * Cout;
  p.P = B << (p.lgMH + p.lgMW);
  taps_full(&p.taps, ksize);
  if (want_tc(math_mode) && uad_conv_tc_supported(UAD_OP_CONVT_WGRAD, B, H, W, Cin, Cout, ksize))
    return launch_wgrad_tensor(p, dw, accumulate, math_mode, ws, ws_bytes, st);
  return uad_launch_wgrad_simt(p, dw, accumulate, ws, ws_bytes, st);
}
