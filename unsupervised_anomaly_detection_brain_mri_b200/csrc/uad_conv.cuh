// Shared parameter blocks of the conv kernel families (SIMT and tcgen05).
#pragma once
#include "uad_common.cuh"

#define UAD_MAX_TAPS 25

// One tap list: input offset (dh, dw) in gathered-tensor pixels, weight tap index wt = kh*k + kw, and the output-pixel
// offset (oh0, ow0) of the output-parity class this list belongs to.
struct TapSet {
  int n;
  int oh0, ow0;
  signed char dh[UAD_MAX_TAPS], dw[UAD_MAX_TAPS], wt[UAD_MAX_TAPS];
};

// Form F / Form T: out[pix(m), n] = epi( sum_{t in taps} sum_ci in[gather(m, t), ci] * wmat[wt(t)][ci][n] )
struct GatherParams {
  const float* in;      // [B, IH, IW, Cin]
  const float* wmat;    // [k*k][Cin][N] row-major
  float* z_out;         // [B, OH, OW, N] or NULL
  float* a_out;         // [B, OH, OW, N] or NULL
  const float* bias;    // [N] or NULL
  const float* gamma;   // [N] or NULL
  const float* beta;    // [N] or NULL
  int B, IH, IW, Cin;
  int lgMH, lgMW;       // log2 of the M-grid (rows of the GEMM are (b, r, s) with s fastest)
  int sh;               // gather stride: input pixel = (r*sh + dh, s*sh + dw)
  int OH, OW, N;
  int osh;              // output pixel = (r*osh + oh0, s*osh + ow0)
  int M;                // B << (lgMH + lgMW)
  int act;
  float alpha, bn_c;
  TapSet taps[4];
  // optional fused 1x1 head (tensor-core stride-1 form with N = 32 only): head_out[pixel] = sum_n a[pixel][n] * head_w[n] + head_b[0]
  const float* head_w;
  const float* head_b;
  float* head_out;
};

// Form W: partial[z][(t, cg)][co] = sum_{pix in chunk z} g[gather(pix, t), cg] * o[pix, co]
struct WgradParams {
  const float* g;       // gathered tensor [B, GH, GW, Cg]
  const float* o;       // M-grid tensor   [B, MH, MW, Co]
  float* partial;
  int B, GH, GW, Cg;
  int lgMH, lgMW, sh, Co;
  int Mp;               // ntaps * Cg
  int P;                // B << (lgMH + lgMW)
  int chunk;
  TapSet taps;
};

// ---- SIMT launchers (uad_conv_simt.cu)
int uad_launch_gather_simt(const GatherParams& p, int nclasses, cudaStream_t st);
int uad_wgrad_plan(int Mp, int Co, int P, int* splits, int* chunk);
int uad_launch_wgrad_simt(WgradParams p, float* out, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st);
int uad_launch_splitk_reduce(const float* partial, int splits, size_t n, float* out, int accumulate, cudaStream_t st);
int uad_launch_transpose_taps(const float* w, float* wT, int T, int A, int Bd, cudaStream_t st);
int uad_launch_conv_c1_fwd(const float* x, const float* w, const float* bias, const float* gamma, const float* beta,
                           float* z_out, float* a_out, int B, int H, int W, int Cout, int ksize, int act, float alpha,
                           float bn_c, cudaStream_t st);
int uad_conv_c1_wgrad_blocks(int B, int H);
int uad_launch_conv_c1_wgrad(const float* x, const float* dz, float* dw, int B, int H, int W, int Cout, int ksize,
                             int accumulate, void* ws, size_t ws_bytes, cudaStream_t st);
int uad_launch_conv_c1_dgrad(const float* dz, const float* w, float* dx, int B, int H, int W, int Cout, int ksize,
                             cudaStream_t st);

// ---- tcgen05 launchers (uad_conv_tc.cu)
int uad_tc_gather_supported(int Cin, int N, int lgMH, int lgMW);
size_t uad_tc_gather_ws_bytes(int ksize, int Cin, int N);
int uad_launch_gather_tc(const GatherParams& p, int nclasses, int ksize, bool weights_transposed, const float* w_raw,
                         int math_mode, void* ws, size_t ws_bytes, cudaStream_t st);
// ---- halo-resident SS-form tcgen05 launcher (uad_conv_hs.cu; round 2): M-grids of at least 16 x 8
int uad_hs_gather_supported(int Cin, int N, int lgMH, int lgMW, int nclasses);
size_t uad_hs_gather_ws_bytes(int ksize, int Cin, int N);
int uad_launch_gather_hs(const GatherParams& p, int nclasses, int ksize, bool weights_transposed, const float* w_raw, bool fast,
                         void* ws, size_t ws_bytes, cudaStream_t st);
// ---- MN-major SS-form Form-W kernel (uad_conv_ws.cu; round 2)
int uad_ws_wgrad_supported(int Cg, int Co, int lgMH, int lgMW);
size_t uad_ws_wgrad_ws_bytes(int Cg, int Co, int B, int MH, int MW);
int uad_launch_wgrad_ss(const WgradParams& w, float* out, int accumulate, bool fast, void* ws, size_t ws_bytes, cudaStream_t st);
int uad_tc_wgrad_supported(int Cg, int Co, int lgMH, int lgMW);
size_t uad_tc_wgrad_ws_bytes(int Cg, int Co, int P);
int uad_launch_wgrad_tc(const WgradParams& w, float* out, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st);
