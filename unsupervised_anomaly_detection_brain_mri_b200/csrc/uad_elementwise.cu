// HBM-bound elementwise / reduction kernels of the hot path: BN+activation backward, reparameterise+KL, fused final
// 1x1 conv + L1, loss scalars, TF-form Adam, Philox RNG, small helpers.  All reductions are two-stage and deterministic.
#include "uad_common.cuh"

// ------------------------------------------------------------------------------------------------ error plumbing
static thread_local char g_uad_err[512] = "";

int uad_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_uad_err, sizeof(g_uad_err), fmt, ap);
  va_end(ap);
  return 1;
}
extern "C" const char* uad_last_error(void) { return g_uad_err; }
extern "C" int uad_abi_version(void) { return UAD_ABI_VERSION; }
long long g_uad_launches = 0;
extern "C" long long uad_launch_count(void) { return g_uad_launches; }

// ------------------------------------------------------------------------------------------------ act + frozen-BN backward
// stage 1: dz = gamma*bn_c * da * act'(u), per-block partial sums of du and du*z per channel.
// FROM_A (act | UAD_ACT_FROM_OUTPUT): the second operand is the block's OUTPUT a = act(u) instead of its pre-BN input z
// (the forward then never writes z).  For the piecewise-linear activations u is recovered exactly up to one rounding
// (LeakyReLU: u = a > 0 ? a : a/alpha, same sign as a; ReLU: du = 0 wherever u is unknown); the second partial sum
// becomes sum du*(u - beta) = gamma*bn_c * sum du*z, and stage 2 divides by gamma instead of multiplying by bn_c.
template <bool FROM_A>
__global__ void __launch_bounds__(256) act_bn_bwd_kernel(const float* __restrict__ da, const float* __restrict__ z,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float* __restrict__ dz, float* __restrict__ partial, long long rows,
                                                         int C, int act, float alpha, float bn_c) {
  __shared__ float red[2][256][4];
  const int tpr = C / 4;                 // threads per row
  const int cg = threadIdx.x % tpr;      // channel group (4 channels)
  const int rl = threadIdx.x / tpr;
  const int rpp = 256 / tpr;             // rows per pass
  float sc[4], sf[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j] = gamma ? gamma[cg * 4 + j] * bn_c : 1.f;
    sf[j] = beta ? beta[cg * 4 + j] : 0.f;
  }
  float s_du[4] = {0.f, 0.f, 0.f, 0.f}, s_duz[4] = {0.f, 0.f, 0.f, 0.f};
  const float inv_alpha = alpha != 0.f ? 1.f / alpha : 0.f;
  for (long long r = (long long)blockIdx.x * rpp + rl; r < rows; r += (long long)gridDim.x * rpp) {
    const size_t o = (size_t)r * C + cg * 4;
    const float4 g4 = *reinterpret_cast<const float4*>(da + o);
    const float4 z4 = *reinterpret_cast<const float4*>(z + o);
    const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, zv[4] = {z4.x, z4.y, z4.z, z4.w};
    float out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float du;
      if (FROM_A) {
        const float u = uad_preact_from_output(zv[j], act, inv_alpha);
        du = gv[j] * uad_act_grad(u, act, alpha);
        s_duz[j] += du * (u - sf[j]);
      } else {
        const float u = sc[j] * zv[j] + sf[j];
        du = gv[j] * uad_act_grad(u, act, alpha);
        s_duz[j] += du * zv[j];
      }
      s_du[j] += du;
      out[j] = sc[j] * du;
    }
    *reinterpret_cast<float4*>(dz + o) = make_float4(out[0], out[1], out[2], out[3]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[0][threadIdx.x][j] = s_du[j];
    red[1][threadIdx.x][j] = s_duz[j];
  }
  __syncthreads();
  if (threadIdx.x < 2 * C) {
    const int which = threadIdx.x / C, c = threadIdx.x % C;
    float s = 0.f;
    for (int k = 0; k < rpp; ++k) s += red[which][k * tpr + c / 4][c % 4];
    partial[(size_t)blockIdx.x * 2 * C + threadIdx.x] = s;
  }
}

// stage 2: reduce block partials, finalise dgamma / dbeta / dbias.  One BLOCK per channel: 128 threads stride over the block
// partials (<= 5 independent loads each instead of 19 dependent iterations of one warp: 19 us -> ~3 us per call), then a fixed-order
// tree -> deterministic.
__global__ void __launch_bounds__(128) act_bn_bwd_final_kernel(const float* __restrict__ partial, int nblocks, const float* __restrict__ gamma,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
                                                               int C, float bn_c, int accumulate, int from_a) {
  __shared__ float sh[2][128];
  const int c = blockIdx.x, t = threadIdx.x;
  float s_du = 0.f, s_duz = 0.f;
  for (int b = t; b < nblocks; b += 128) {
    s_du += partial[(size_t)b * 2 * C + c];
    s_duz += partial[(size_t)b * 2 * C + C + c];
  }
  sh[0][t] = s_du;
  sh[1][t] = s_duz;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (t < o) { sh[0][t] += sh[0][t + o]; sh[1][t] += sh[1][t + o]; }
    __syncthreads();
  }
  if (t != 0) return;
  s_du = sh[0][0];
  s_duz = sh[1][0];
  const float sc = gamma ? gamma[c] * bn_c : 1.f;
  // from_a: s_duz = sum du*(u - beta) = gamma*bn_c*sum du*z  ->  dgamma = s_duz / gamma
  const float dg = from_a ? (gamma ? s_duz / gamma[c] : 0.f) : bn_c * s_duz;
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + dg;
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + s_du;
  if (dbias) dbias[c] = (accumulate ? dbias[c] : 0.f) + sc * s_du;
}

static int rowreduce_blocks(long long rows, int C) {
  const int rpp = 256 / (C / 4);
  long long b = (rows + rpp - 1) / rpp;
  const long long cap = 4 * UAD_NUM_SMS;
  return (int)(b < cap ? b : cap);
}

extern "C" size_t uad_rowreduce_workspace_bytes(long long rows, int C) {
  if (C < 4 || C % 4) return 0;
  return (size_t)rowreduce_blocks(rows, C) * 2 * C * sizeof(float) + 256;
}

extern "C" int uad_act_bn_bwd(const float* da, const float* z, const float* gamma, const float* beta, float* dz,
                              float* dgamma, float* dbeta, float* dbias, long long rows, int C, int act, float alpha,
                              float bn_c, int accumulate, void* ws, size_t ws_bytes, void* stream) {
  UAD_REQUIRE(C % 4 == 0 && C >= 4 && C <= 128 && 256 % (C / 4) == 0, "uad_act_bn_bwd: unsupported C=%d", C);
  UAD_REQUIRE((gamma == nullptr) == (beta == nullptr), "uad_act_bn_bwd: gamma/beta must both be set or both NULL");
  const int from_a = (act & UAD_ACT_FROM_OUTPUT) ? 1 : 0;
  act &= ~UAD_ACT_FROM_OUTPUT;
  UAD_REQUIRE(!from_a || act == UAD_ACT_NONE || act == UAD_ACT_RELU || (act == UAD_ACT_LEAKY && alpha > 0.f),
              "uad_act_bn_bwd: UAD_ACT_FROM_OUTPUT needs a piecewise-linear activation (act=%d alpha=%g)", act, (double)alpha);
  const int nb = rowreduce_blocks(rows, C);
  UAD_REQUIRE(ws && ws_bytes >= (size_t)nb * 2 * C * sizeof(float), "uad_act_bn_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (from_a)
    act_bn_bwd_kernel<true><<<nb, 256, 0, st>>>(da, z, gamma, beta, dz, (float*)ws, rows, C, act, alpha, bn_c);
  else
    act_bn_bwd_kernel<false><<<nb, 256, 0, st>>>(da, z, gamma, beta, dz, (float*)ws, rows, C, act, alpha, bn_c);
  UAD_LAUNCH_CHECK("act_bn_bwd");
  act_bn_bwd_final_kernel<<<C, 128, 0, st>>>((const float*)ws, nb, gamma, dgamma, dbeta, dbias, C, bn_c,
                                                            accumulate, from_a);
  UAD_LAUNCH_CHECK("act_bn_bwd_final");
  return 0;
}

// ------------------------------------------------------------------------------------------------ dropout -> frozen BN -> act
// z = x * (mask ? mask*keep : 1);  a = act(gamma*bn_c*z + beta).  The spatial-bottleneck models (reference
// models/autoencoder_spatial.py:16-23) feed the dropped-out encoder output straight into the decoder's BN + ReLU.
__global__ void mask_bn_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mask, float keep,
                                       const float* __restrict__ gamma, const float* __restrict__ beta, float bn_c, int act,
                                       float alpha, float* __restrict__ z_out, float* __restrict__ a_out, size_t n, int C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  float z = x[i];
  if (mask) z *= mask[i] * keep;
  if (z_out) z_out[i] = z;
  if (a_out) {
    const float u = gamma ? gamma[c] * bn_c * z + beta[c] : z;
    a_out[i] = uad_act(u, act, alpha);
  }
}

extern "C" int uad_mask_bn_act_fwd(const float* x, const float* mask, float keep, const float* gamma, const float* beta, float bn_c,
                                   int act, float alpha, float* z_out, float* a_out, long long rows, int C, void* stream) {
  UAD_REQUIRE(rows > 0 && C > 0, "uad_mask_bn_act_fwd: bad dims");
  UAD_REQUIRE((gamma == nullptr) == (beta == nullptr), "uad_mask_bn_act_fwd: gamma/beta must both be set or both NULL");
  const size_t n = (size_t)rows * C;
  mask_bn_act_fwd_kernel<<<uad_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, mask, keep, gamma, beta, bn_c, act, alpha, z_out,
                                                                            a_out, n, C);
  UAD_LAUNCH_CHECK("mask_bn_act_fwd");
  return 0;
}

// y = x * (mask ? mask : 1) * scale   (dropout backward; y may alias x)
__global__ void mask_scale_kernel(const float* __restrict__ x, const float* __restrict__ mask, float scale, float* __restrict__ y,
                                  size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = x[i] * (mask ? mask[i] * scale : scale);
}

extern "C" int uad_mask_scale(const float* x, const float* mask, float scale, float* y, size_t n, void* stream) {
  mask_scale_kernel<<<uad_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, mask, scale, y, n);
  UAD_LAUNCH_CHECK("mask_scale");
  return 0;
}

// ------------------------------------------------------------------------------------------------ reparameterise + KL
__global__ void reparam_kl_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ ls, const float* __restrict__ eps,
                                      float* __restrict__ sigma, float* __restrict__ z, float* __restrict__ kl, int Z) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  float s = 0.f;
  for (int j = threadIdx.x; j < Z; j += blockDim.x) {
    const size_t o = (size_t)b * Z + j;
    const float m = mu[o], sg = expf(ls[o]);
    if (sigma) sigma[o] = sg;
    if (z) z[o] = eps ? m + eps[o] * sg : m;
    const float s2 = sg * sg;
    s += m * m + s2 - logf(s2) - 1.f;     // trainers/VAE.py:38 literally
  }
  s = uad_warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += red[w];
    if (kl) kl[b] = 0.5f * t;
  }
}

extern "C" int uad_reparam_kl_fwd(const float* mu, const float* log_sigma, const float* eps, float* sigma, float* z,
                                  float* kl, int B, int Z, void* stream) {
  UAD_REQUIRE(B > 0 && Z > 0, "uad_reparam_kl_fwd: bad dims");
  reparam_kl_fwd_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(mu, log_sigma, eps, sigma, z, kl, Z);
  UAD_LAUNCH_CHECK("reparam_kl_fwd");
  return 0;
}

__global__ void reparam_kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ ls, const float* __restrict__ eps,
                                      const float* __restrict__ dz, float kl_scale, float* __restrict__ dmu,
                                      float* __restrict__ dls, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float sg = expf(ls[i]);
  const float g = dz ? dz[i] : 0.f;
  dmu[i] = g + kl_scale * mu[i];
  // d/dls of 0.5*(sigma^2 - log(sigma^2) - 1) = sigma^2 - 1 ; z = mu + eps*sigma -> dz/dls = eps*sigma
  dls[i] = (eps ? g * eps[i] * sg : 0.f) + kl_scale * (sg * sg - 1.f);
}

extern "C" int uad_reparam_kl_bwd(const float* mu, const float* log_sigma, const float* eps, const float* dz,
                                  float kl_scale, float* dmu, float* dls, int B, int Z, void* stream) {
  size_t n = (size_t)B * Z;
  reparam_kl_bwd_kernel<<<uad_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(mu, log_sigma, eps, dz, kl_scale, dmu, dls, n);
  UAD_LAUNCH_CHECK("reparam_kl_bwd");
  return 0;
}

// ------------------------------------------------------------------------------------------------ final 1x1 + L1
// lanes-per-pixel LP = Cin/4; each lane holds a float4 of channels; a block covers `ppb` pixels of ONE sample.
__global__ void __launch_bounds__(256) final1x1_l1_fwd_kernel(const float* __restrict__ a, const float* __restrict__ w,
                                                              const float* __restrict__ bias, const float* __restrict__ x,
                                                              float* __restrict__ xhat, float* __restrict__ l1,
                                                              float* __restrict__ partial, int HW, int Cin, int ppb) {
  __shared__ float red[8];
  const int LP = Cin / 4;
  const int lp = threadIdx.x % LP;
  const int pl = threadIdx.x / LP;
  const int ppp = 256 / LP;                              // pixels per pass
  const int bps = HW / ppb;                              // blocks per sample
  const int b = blockIdx.x / bps;
  const size_t pix0 = (size_t)b * HW + (size_t)(blockIdx.x % bps) * ppb;
  const float4 w4 = *reinterpret_cast<const float4*>(w + lp * 4);
  const float bv = bias ? bias[0] : 0.f;
  float s = 0.f;
  for (int q = pl; q < ppb; q += ppp) {
    const size_t pix = pix0 + q;
    const float4 a4 = __ldg(reinterpret_cast<const float4*>(a + pix * Cin + lp * 4));
    float d = a4.x * w4.x + a4.y * w4.y + a4.z * w4.z + a4.w * w4.w;
    for (int o = LP >> 1; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lp == 0) {
      const float xh = d + bv;
      const float e = fabsf(xh - x[pix]);
      xhat[pix] = xh;
      if (l1) l1[pix] = e;
      s += e;
    }
  }
  s = uad_warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0 && partial) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += red[wv];
    partial[blockIdx.x] = t;
  }
}

__global__ void rec_final_kernel(const float* __restrict__ partial, int bps, float* __restrict__ rec, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s = 0.f;
  for (int k = 0; k < bps; ++k) s += partial[(size_t)b * bps + k];
  rec[b] = s;
}

static int final_ppb(int HW) { return HW < 2048 ? HW : 2048; }

extern "C" int uad_final1x1_l1_fwd(const float* a, const float* w, const float* bias, const float* x, float* xhat,
                                   float* l1, float* rec, int B, int HW, int Cin, void* ws, size_t ws_bytes, void* stream) {
  UAD_REQUIRE(Cin % 4 == 0 && uad_is_pow2(Cin / 4) && Cin <= 128, "uad_final1x1_l1_fwd: unsupported Cin=%d", Cin);
  UAD_REQUIRE(uad_is_pow2(HW), "uad_final1x1_l1_fwd: HW=%d must be a power of two", HW);
  const int ppb = final_ppb(HW), bps = HW / ppb;
  UAD_REQUIRE(ppb % (256 / (Cin / 4)) == 0, "uad_final1x1_l1_fwd: HW=%d too small for Cin=%d", HW, Cin);
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = nullptr;
  if (rec) {
    UAD_REQUIRE(ws && ws_bytes >= (size_t)B * bps * sizeof(float), "uad_final1x1_l1_fwd: workspace too small");
    partial = (float*)ws;
  }
  final1x1_l1_fwd_kernel<<<B * bps, 256, 0, st>>>(a, w, bias, x, xhat, l1, partial, HW, Cin, ppb);
  UAD_LAUNCH_CHECK("final1x1_l1_fwd");
  if (rec) {
    rec_final_kernel<<<uad_cdiv(B, 128), 128, 0, st>>>(partial, bps, rec, B);
    UAD_LAUNCH_CHECK("rec_final");
  }
  return 0;
}

__global__ void __launch_bounds__(256) final1x1_l1_bwd_kernel(const float* __restrict__ a, const float* __restrict__ w,
                                                              const float* __restrict__ x, const float* __restrict__ xhat,
                                                              float scale, float* __restrict__ da, float* __restrict__ partial,
                                                              size_t npix, int Cin) {
  __shared__ float red[256][5];
  const int LP = Cin / 4;
  const int lp = threadIdx.x % LP;
  const int pl = threadIdx.x / LP;
  const int ppp = 256 / LP;
  const float4 w4 = *reinterpret_cast<const float4*>(w + lp * 4);
  float sw[4] = {0.f, 0.f, 0.f, 0.f}, sb = 0.f;
  for (size_t pix = (size_t)blockIdx.x * ppp + pl; pix < npix; pix += (size_t)gridDim.x * ppp) {
    float g;
    if (xhat) {
      const float e = xhat[pix] - x[pix];
      g = (e > 0.f ? scale : (e < 0.f ? -scale : 0.f));
    } else {
      g = x[pix] * scale;                      // direct mode: x holds the incoming gradient d/dxhat
    }
    const float4 a4 = __ldg(reinterpret_cast<const float4*>(a + pix * Cin + lp * 4));
    sw[0] += g * a4.x; sw[1] += g * a4.y; sw[2] += g * a4.z; sw[3] += g * a4.w;
    if (lp == 0) sb += g;
    if (da) *reinterpret_cast<float4*>(da + pix * Cin + lp * 4) = make_float4(g * w4.x, g * w4.y, g * w4.z, g * w4.w);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[threadIdx.x][j] = sw[j];
  red[threadIdx.x][4] = sb;
  __syncthreads();
  if (threadIdx.x <= Cin) {
    float s = 0.f;
    if (threadIdx.x < Cin) {
      const int c = threadIdx.x;
      for (int k = 0; k < ppp; ++k) s += red[k * LP + c / 4][c % 4];
    } else {
      for (int k = 0; k < ppp; ++k) s += red[k * LP][4];
    }
    partial[(size_t)blockIdx.x * (Cin + 1) + threadIdx.x] = s;
  }
}

__global__ void final_bwd_reduce_kernel(const float* __restrict__ partial, int nblocks, int Cin, float* __restrict__ dw,
                                        float* __restrict__ dbias, int accumulate) {
  const int c = threadIdx.x;
  if (c > Cin) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * (Cin + 1) + c];
  if (c < Cin) { if (dw) dw[c] = (accumulate ? dw[c] : 0.f) + s; }
  else { if (dbias) dbias[0] = (accumulate ? dbias[0] : 0.f) + s; }
}

extern "C" int uad_final1x1_l1_bwd(const float* a, const float* w, const float* x, const float* xhat, float scale,
                                   float* da, float* dw, float* dbias, int B, int HW, int Cin, int accumulate, void* ws,
                                   size_t ws_bytes, void* stream) {
  UAD_REQUIRE(Cin % 4 == 0 && uad_is_pow2(Cin / 4) && Cin <= 128, "uad_final1x1_l1_bwd: unsupported Cin=%d", Cin);
  const size_t npix = (size_t)B * HW;
  const int ppp = 256 / (Cin / 4);
  long long nb = (npix + ppp - 1) / ppp;
  if (nb > 4 * UAD_NUM_SMS) nb = 4 * UAD_NUM_SMS;
  UAD_REQUIRE(ws && ws_bytes >= (size_t)nb * (Cin + 1) * sizeof(float), "uad_final1x1_l1_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  final1x1_l1_bwd_kernel<<<(int)nb, 256, 0, st>>>(a, w, x, xhat, scale, da, (float*)ws, npix, Cin);
  UAD_LAUNCH_CHECK("final1x1_l1_bwd");
  final_bwd_reduce_kernel<<<1, 256, 0, st>>>((const float*)ws, (int)nb, Cin, dw, dbias, accumulate);
  UAD_LAUNCH_CHECK("final_bwd_reduce");
  return 0;
}

// backward of the final 1x1 conv alone for an arbitrary incoming gradient dxhat [B*HW] (f-AnoGAN generator, whose head is
// sigmoid(conv1x1) rather than an L1 residual: models/fanogan.py:41,46)
extern "C" int uad_final1x1_bwd(const float* a, const float* w, const float* dxhat, float* da, float* dw, float* dbias, int B,
                                int HW, int Cin, int accumulate, void* ws, size_t ws_bytes, void* stream) {
  UAD_REQUIRE(Cin % 4 == 0 && uad_is_pow2(Cin / 4) && Cin <= 128, "uad_final1x1_bwd: unsupported Cin=%d", Cin);
  const size_t npix = (size_t)B * HW;
  const int ppp = 256 / (Cin / 4);
  long long nb = (npix + ppp - 1) / ppp;
  if (nb > 4 * UAD_NUM_SMS) nb = 4 * UAD_NUM_SMS;
  UAD_REQUIRE(ws && ws_bytes >= (size_t)nb * (Cin + 1) * sizeof(float), "uad_final1x1_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  final1x1_l1_bwd_kernel<<<(int)nb, 256, 0, st>>>(a, w, dxhat, nullptr, 1.f, da, (float*)ws, npix, Cin);
  UAD_LAUNCH_CHECK("final1x1_bwd");
  final_bwd_reduce_kernel<<<1, 256, 0, st>>>((const float*)ws, (int)nb, Cin, dw, dbias, accumulate);
  UAD_LAUNCH_CHECK("final_bwd_reduce");
  return 0;
}

// ---- fused: backward of (final 1x1 conv + L1) AND of the preceding "z -> frozen BN -> activation" block.
// Reads z (pre-BN output of the last transposed conv), x, xhat; writes dz; never materialises da or re-reads a.
// FROM_A: reads the block's output a instead of z (see act_bn_bwd_kernel).
template <bool FROM_A>
__global__ void __launch_bounds__(256) final_bwd_fused_kernel(const float* __restrict__ z, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ w,
                                                              const float* __restrict__ x, const float* __restrict__ xhat,
                                                              float scale, float* __restrict__ dz, float* __restrict__ partial,
                                                              size_t npix, int Cin, int act, float alpha, float bn_c) {
  __shared__ float red[256][13];
  const int LP = Cin / 4;
  const int lp = threadIdx.x % LP;
  const int pl = threadIdx.x / LP;
  const int ppp = 256 / LP;
  const float4 w4 = *reinterpret_cast<const float4*>(w + lp * 4);
  const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
  float sc[4], sf[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j] = gamma ? gamma[lp * 4 + j] * bn_c : 1.f;
    sf[j] = beta ? beta[lp * 4 + j] : 0.f;
  }
  float s_du[4] = {0.f, 0.f, 0.f, 0.f}, s_duz[4] = {0.f, 0.f, 0.f, 0.f}, s_w[4] = {0.f, 0.f, 0.f, 0.f}, s_b = 0.f;
  const float inv_alpha = alpha != 0.f ? 1.f / alpha : 0.f;
  for (size_t pix = (size_t)blockIdx.x * ppp + pl; pix < npix; pix += (size_t)gridDim.x * ppp) {
    const float e = xhat[pix] - x[pix];
    const float g = (e > 0.f ? scale : (e < 0.f ? -scale : 0.f));
    const float4 z4 = __ldg(reinterpret_cast<const float4*>(z + pix * Cin + lp * 4));
    const float zv[4] = {z4.x, z4.y, z4.z, z4.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a, du;
      if (FROM_A) {
        a = zv[j];
        const float u = uad_preact_from_output(a, act, inv_alpha);
        du = g * wv[j] * uad_act_grad(u, act, alpha);
        s_duz[j] += du * (u - sf[j]);
      } else {
        const float u = sc[j] * zv[j] + sf[j];
        a = uad_act(u, act, alpha);
        du = g * wv[j] * uad_act_grad(u, act, alpha);
        s_duz[j] += du * zv[j];
      }
      s_w[j] += g * a;
      s_du[j] += du;
      o[j] = sc[j] * du;
    }
    if (lp == 0) s_b += g;
    *reinterpret_cast<float4*>(dz + pix * Cin + lp * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[threadIdx.x][j] = s_du[j];
    red[threadIdx.x][4 + j] = s_duz[j];
    red[threadIdx.x][8 + j] = s_w[j];
  }
  red[threadIdx.x][12] = s_b;
  __syncthreads();
  // partial layout per block: [s_du(Cin) | s_duz(Cin) | s_w(Cin) | s_b]
  for (int idx = threadIdx.x; idx < 3 * Cin + 1; idx += blockDim.x) {
    float s = 0.f;
    if (idx < 3 * Cin) {
      const int which = idx / Cin, c = idx % Cin;
      for (int k = 0; k < ppp; ++k) s += red[k * LP + c / 4][which * 4 + (c % 4)];
    } else {
      for (int k = 0; k < ppp; ++k) s += red[k * LP][12];
    }
    partial[(size_t)blockIdx.x * (3 * Cin + 1) + idx] = s;
  }
}

__global__ void final_bwd_fused_reduce_kernel(const float* __restrict__ partial, int nblocks, int Cin, const float* __restrict__ gamma,
                                              float bn_c, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                              float* __restrict__ dbias_prev, float* __restrict__ dw, float* __restrict__ dbias,
                                              int accumulate, int from_a) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;       // one warp per channel (+1 for the bias)
  const int lane = threadIdx.x & 31;
  const int stride = 3 * Cin + 1;
  if (c < Cin) {
    float a = 0.f, b = 0.f, d = 0.f;
    for (int k = lane; k < nblocks; k += 32) {
      a += partial[(size_t)k * stride + c];
      b += partial[(size_t)k * stride + Cin + c];
      d += partial[(size_t)k * stride + 2 * Cin + c];
    }
    a = uad_warp_sum(a); b = uad_warp_sum(b); d = uad_warp_sum(d);
    if (lane != 0) return;
    const float scv = gamma ? gamma[c] * bn_c : 1.f;
    const float dg = from_a ? (gamma ? b / gamma[c] : 0.f) : bn_c * b;
    if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + dg;
    if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + a;
    if (dbias_prev) dbias_prev[c] = (accumulate ? dbias_prev[c] : 0.f) + scv * a;
    if (dw) dw[c] = (accumulate ? dw[c] : 0.f) + d;
  } else if (c == Cin) {
    float sb = 0.f;
    for (int k = lane; k < nblocks; k += 32) sb += partial[(size_t)k * stride + 3 * Cin];
    sb = uad_warp_sum(sb);
    if (lane == 0 && dbias) dbias[0] = (accumulate ? dbias[0] : 0.f) + sb;
  }
}

extern "C" int uad_final1x1_l1_bwd_fused(const float* z, const float* gamma, const float* beta, const float* w, const float* x,
                                         const float* xhat, float scale, float* dz, float* dgamma, float* dbeta,
                                         float* dbias_prev, float* dw, float* dbias, int B, int HW, int Cin, int act, float alpha,
                                         float bn_c, int accumulate, void* ws, size_t ws_bytes, void* stream) {
  UAD_REQUIRE(Cin % 4 == 0 && uad_is_pow2(Cin / 4) && Cin <= 128, "uad_final1x1_l1_bwd_fused: unsupported Cin=%d", Cin);
  UAD_REQUIRE((gamma == nullptr) == (beta == nullptr), "uad_final1x1_l1_bwd_fused: gamma/beta must both be set or both NULL");
  const int from_a = (act & UAD_ACT_FROM_OUTPUT) ? 1 : 0;
  act &= ~UAD_ACT_FROM_OUTPUT;
  UAD_REQUIRE(!from_a || act == UAD_ACT_NONE || act == UAD_ACT_RELU || (act == UAD_ACT_LEAKY && alpha > 0.f),
              "uad_final1x1_l1_bwd_fused: UAD_ACT_FROM_OUTPUT needs a piecewise-linear activation (act=%d alpha=%g)", act,
              (double)alpha);
  const size_t npix = (size_t)B * HW;
  const int ppp = 256 / (Cin / 4);
  long long nb = (npix + ppp - 1) / ppp;
  if (nb > 4 * UAD_NUM_SMS) nb = 4 * UAD_NUM_SMS;
  UAD_REQUIRE(ws && ws_bytes >= (size_t)nb * (3 * Cin + 1) * sizeof(float), "uad_final1x1_l1_bwd_fused: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (from_a)
    final_bwd_fused_kernel<true><<<(int)nb, 256, 0, st>>>(z, gamma, beta, w, x, xhat, scale, dz, (float*)ws, npix, Cin, act, alpha, bn_c);
  else
    final_bwd_fused_kernel<false><<<(int)nb, 256, 0, st>>>(z, gamma, beta, w, x, xhat, scale, dz, (float*)ws, npix, Cin, act, alpha, bn_c);
  UAD_LAUNCH_CHECK("final_bwd_fused");
  final_bwd_fused_reduce_kernel<<<uad_cdiv((Cin + 1) * 32, 256), 256, 0, st>>>((const float*)ws, (int)nb, Cin, gamma, bn_c, dgamma, dbeta,
                                                                       dbias_prev, dw, dbias, accumulate, from_a);
  UAD_LAUNCH_CHECK("final_bwd_fused_reduce");
  return 0;
}

// ------------------------------------------------------------------------------------------------ loss scalars
__global__ void loss_scalars_kernel(const float* __restrict__ rec, const float* __restrict__ kl, float* __restrict__ out, int B) {
  __shared__ float red[3][32];
  float a = 0.f, b = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float r = rec[i], k = kl ? kl[i] : 0.f;
    a += r; b += k; c += r + k;
  }
  a = uad_warp_sum(a); b = uad_warp_sum(b); c = uad_warp_sum(c);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; red[2][threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int w = 0; w < blockDim.x / 32; ++w) s += red[threadIdx.x][w];
    out[threadIdx.x] = s / (float)B;
  }
}

extern "C" int uad_loss_scalars(const float* rec, const float* kl, float* out3, int B, void* stream) {
  loss_scalars_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(rec, kl, out3, B);
  UAD_LAUNCH_CHECK("loss_scalars");
  return 0;
}

// ------------------------------------------------------------------------------------------------ TF-form Adam
__global__ void adam_tf_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                               size_t n, float lr_t, float b1, float b2, float eps, float gs,
                               const long long* __restrict__ step_dev) {
  if (step_dev) {
    const double t = (double)(*step_dev);
    lr_t = (float)((double)lr_t * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
  }
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 p4 = *reinterpret_cast<float4*>(p + i), g4 = *reinterpret_cast<const float4*>(g + i);
    float4 m4 = *reinterpret_cast<float4*>(m + i), v4 = *reinterpret_cast<float4*>(v + i);
    float pp[4] = {p4.x, p4.y, p4.z, p4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
    float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = gg[j] * gs;
      mm[j] = b1 * mm[j] + (1.f - b1) * gj;
      vv[j] = b2 * vv[j] + (1.f - b2) * gj * gj;
      pp[j] = pp[j] - lr_t * mm[j] / (sqrtf(vv[j]) + eps);
    }
    *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
  } else {
    for (; i < n; ++i) {
      const float gj = g[i] * gs;
      const float mj = b1 * m[i] + (1.f - b1) * gj;
      const float vj = b2 * v[i] + (1.f - b2) * gj * gj;
      m[i] = mj; v[i] = vj;
      p[i] = p[i] - lr_t * mj / (sqrtf(vj) + eps);
    }
  }
}

extern "C" int uad_adam_tf_step(float* params, const float* grads, float* m, float* v, size_t n, float lr_t, float b1,
                                float b2, float eps, float grad_scale, const int64_t* step_dev, void* stream) {
  UAD_REQUIRE(((uintptr_t)params % 16 == 0) && ((uintptr_t)grads % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
              ((uintptr_t)v % 16 == 0), "uad_adam_tf_step: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  adam_tf_kernel<<<uad_cdiv((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(params, grads, m, v, n, lr_t, b1, b2, eps,
                                                                              grad_scale, (const long long*)step_dev);
  UAD_LAUNCH_CHECK("adam_tf");
  return 0;
}

// ------------------------------------------------------------------------------------------------ Philox-4x32-10
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__global__ void randn_kernel(float* __restrict__ out, size_t n, uint64_t seed, uint64_t offset,
                             const uint64_t* __restrict__ offset_dev) {
  if (offset_dev) offset += *offset_dev;
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q * 4 >= n) return;
  const uint64_t ctr = offset + q;
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x5eedu, 0u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  float r[4];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float u1 = ((float)c[2 * h] + 1.0f) * 2.3283064365386963e-10f;      // (0,1]
    const float u2 = (float)c[2 * h + 1] * 2.3283064365386963e-10f;
    const float rad = sqrtf(-2.f * logf(u1));
    float sn, cs;
    sincosf(6.283185307179586f * u2, &sn, &cs);
    r[2 * h] = rad * cs;
    r[2 * h + 1] = rad * sn;
  }
  for (int j = 0; j < 4 && q * 4 + j < n; ++j) out[q * 4 + j] = r[j];
}

extern "C" int uad_randn(float* out, size_t n, uint64_t seed, uint64_t offset, const uint64_t* offset_dev, void* stream) {
  if (n == 0) return 0;
  randn_kernel<<<uad_cdiv((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(out, n, seed, offset, offset_dev);
  UAD_LAUNCH_CHECK("randn");
  return 0;
}

__global__ void dropout_mask_kernel(float* __restrict__ mask, size_t n, float rate, uint64_t seed, uint64_t offset,
                                    const uint64_t* __restrict__ offset_dev) {
  if (offset_dev) offset += *offset_dev;
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q * 4 >= n) return;
  const uint64_t ctr = offset + q;
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0xd509u, 0u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  for (int j = 0; j < 4 && q * 4 + j < n; ++j) {
    const float u = (float)(c[j] >> 8) * 5.9604644775390625e-08f;             // [0,1)
    mask[q * 4 + j] = (u >= rate) ? 1.f : 0.f;                                // Keras: keep where uniform >= rate
  }
}

extern "C" int uad_dropout_mask(float* mask, size_t n, float rate, uint64_t seed, uint64_t offset,
                                const uint64_t* offset_dev, void* stream) {
  if (n == 0) return 0;
  dropout_mask_kernel<<<uad_cdiv((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(mask, n, rate, seed, offset, offset_dev);
  UAD_LAUNCH_CHECK("dropout_mask");
  return 0;
}

// uniform [0,1) stream (tf.random_uniform of the WGAN-GP interpolation, models/fanogan.py:67)
__global__ void uniform_kernel(float* __restrict__ out, size_t n, uint64_t seed, uint64_t offset,
                               const uint64_t* __restrict__ offset_dev) {
  if (offset_dev) offset += *offset_dev;
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q * 4 >= n) return;
  const uint64_t ctr = offset + q;
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0xa1fau, 0u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  for (int j = 0; j < 4 && q * 4 + j < n; ++j) out[q * 4 + j] = (float)(c[j] >> 8) * 5.9604644775390625e-08f;
}

extern "C" int uad_uniform(float* out, size_t n, uint64_t seed, uint64_t offset, const uint64_t* offset_dev, void* stream) {
  if (n == 0) return 0;
  uniform_kernel<<<uad_cdiv((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(out, n, seed, offset, offset_dev);
  UAD_LAUNCH_CHECK("uniform");
  return 0;
}

// ------------------------------------------------------------------------------------------------ helpers
__global__ void counter_add_kernel(uint64_t* c, uint64_t inc) { *c += inc; }
extern "C" int uad_counter_add(uint64_t* counter_dev, uint64_t inc, void* stream) {
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter_dev, inc);
  UAD_LAUNCH_CHECK("counter_add");
  return 0;
}
__global__ void mul_abs_kernel(const float* __restrict__ l1, const float* __restrict__ gx, float* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = l1[i] * fabsf(gx[i]);
}
extern "C" int uad_mul_abs(const float* l1, const float* gx, float* out, size_t n, void* stream) {
  if (n == 0) return 0;
  mul_abs_kernel<<<uad_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(l1, gx, out, n);
  UAD_LAUNCH_CHECK("mul_abs");
  return 0;
}

__global__ void axpby_kernel(float a, const float* __restrict__ x, float b, float* __restrict__ y, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a * x[i] + (b != 0.f ? b * y[i] : 0.f);
}
extern "C" int uad_axpby(float a, const float* x, float b, float* y, size_t n, void* stream) {
  if (n == 0) return 0;
  axpby_kernel<<<uad_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(a, x, b, y, n);
  UAD_LAUNCH_CHECK("axpby");
  return 0;
}

__global__ void l1_direct_term_kernel(const float* __restrict__ x, const float* __restrict__ xhat, float scale,
                                      float* __restrict__ gx, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float e = xhat[i] - x[i];
  gx[i] += (e > 0.f ? -scale : (e < 0.f ? scale : 0.f));
}
extern "C" int uad_l1_direct_term(const float* x, const float* xhat, float scale, float* gx, size_t n, void* stream) {
  if (n == 0) return 0;
  l1_direct_term_kernel<<<uad_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, xhat, scale, gx, n);
  UAD_LAUNCH_CHECK("l1_direct_term");
  return 0;
}

// ------------------------------------------------------------------------------------------------ LayerNormalization([1,2])
// tf.keras.layers.LayerNormalization(axis=[1,2]) on NHWC (models/customlayers.py:22,30,35 with use_batchnorm=False; f-AnoGAN):
// mean / variance over (H, W) per (sample, channel); gamma / beta of shape [H, W]; epsilon 1e-3.
// stage 1: per (b, split) partial sum / sum of squares per channel (fp32, <= 256 terms per thread-level chain)
__global__ void __launch_bounds__(256) ln_hw_partial_kernel(const float* __restrict__ x, double* __restrict__ partial, int HW, int C,
                                                            int rows_per_split) {
  __shared__ float red[2][256];
  const int tpr = C;                                   // one thread per channel, 256 / C pixel lanes
  const int c = threadIdx.x % tpr, pl = threadIdx.x / tpr, ppp = 256 / tpr;
  const int b = blockIdx.x, sp = blockIdx.y;
  const int r0 = sp * rows_per_split, r1 = min(HW, r0 + rows_per_split);
  float s = 0.f, ss = 0.f;
  for (int r = r0 + pl; r < r1; r += ppp) {
    const float v = x[((size_t)b * HW + r) * C + c];
    s += v;
    ss = fmaf(v, v, ss);
  }
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.x < C) {
    double a = 0.0, q = 0.0;
    for (int k = 0; k < ppp; ++k) { a += (double)red[0][k * tpr + threadIdx.x]; q += (double)red[1][k * tpr + threadIdx.x]; }
    double* dst = partial + (((size_t)b * gridDim.y + sp) * C + threadIdx.x) * 2;
    dst[0] = a;
    dst[1] = q;
  }
}

// stage 2: mean / rstd per (b, c) in float64, stored as float
__global__ void ln_hw_stats_kernel(const double* __restrict__ partial, int splits, int HW, int BC_C, int C, float eps,
                                   float* __restrict__ mean, float* __restrict__ rstd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;       // (b, c)
  if (i >= BC_C) return;
  const int b = i / C, c = i % C;
  double a = 0.0, q = 0.0;
  for (int s = 0; s < splits; ++s) {
    const double* src = partial + (((size_t)b * splits + s) * C + c) * 2;
    a += src[0];
    q += src[1];
  }
  const double m = a / HW;
  double var = q / HW - m * m;
  if (var < 0.0) var = 0.0;
  mean[i] = (float)m;
  rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

// stage 3: y = act((x - mean) * rstd * gamma[h,w] + beta[h,w])
__global__ void ln_hw_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y,
                                   size_t n4, int HW, int C, int act, float alpha) {
  const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= n4) return;
  const size_t e = i4 * 4;
  const int c = (int)(e % C);
  const size_t pix = e / C;
  const int hw = (int)(pix % HW);
  const int b = (int)(pix / HW);
  const float4 v = *reinterpret_cast<const float4*>(x + e);
  const float4 m = *reinterpret_cast<const float4*>(mean + (size_t)b * C + c);
  const float4 r = *reinterpret_cast<const float4*>(rstd + (size_t)b * C + c);
  const float g = gamma[hw], bt = beta[hw];
  float4 o;
  o.x = uad_act((v.x - m.x) * r.x * g + bt, act, alpha);
  o.y = uad_act((v.y - m.y) * r.y * g + bt, act, alpha);
  o.z = uad_act((v.z - m.z) * r.z * g + bt, act, alpha);
  o.w = uad_act((v.w - m.w) * r.w * g + bt, act, alpha);
  *reinterpret_cast<float4*>(y + e) = o;
}

int uad_layernorm_hw_apply(const float* x, const float* mean, const float* rstd, const float* gamma_hw, const float* beta_hw,
                           float* y, int B, int HW, int C, int act, float alpha, cudaStream_t st) {
  const size_t n4 = (size_t)B * HW * C / 4;
  ln_hw_apply_kernel<<<uad_cdiv(n4, 256), 256, 0, st>>>(x, mean, rstd, gamma_hw, beta_hw, y, n4, HW, C, act, alpha);
  UAD_LAUNCH_CHECK("ln_hw_apply");
  return 0;
}

static int ln_splits(int HW) { int s = HW / 256; return s < 1 ? 1 : (s > 64 ? 64 : s); }

extern "C" size_t uad_layernorm_hw_workspace_bytes(int B, int HW, int C) {
  return (size_t)B * ln_splits(HW) * C * 2 * sizeof(double) + 2 * (size_t)B * C * sizeof(float) + 512;
}

extern "C" int uad_layernorm_hw_fwd(const float* x, const float* gamma_hw, const float* beta_hw, float* y, int B, int HW,
                                    int C, float eps, int act, float alpha, void* ws, size_t ws_bytes, void* stream) {
  UAD_REQUIRE(C % 4 == 0 && C <= 256 && 256 % C == 0, "uad_layernorm_hw_fwd: unsupported C=%d", C);
  UAD_REQUIRE(ws && ws_bytes >= uad_layernorm_hw_workspace_bytes(B, HW, C), "uad_layernorm_hw_fwd: workspace too small");
  UAD_REQUIRE(((uintptr_t)ws % 16) == 0, "uad_layernorm_hw_fwd: unaligned workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const int splits = ln_splits(HW);
  const int rps = uad_cdiv(HW, splits);
  double* partial = reinterpret_cast<double*>(ws);
  size_t poff = ((size_t)B * splits * C * 2 * sizeof(double) + 255) / 256 * 256;
  float* mean = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + poff);
  float* rstd = mean + (size_t)B * C;
  ln_hw_partial_kernel<<<dim3(B, splits), 256, 0, st>>>(x, partial, HW, C, rps);
  UAD_LAUNCH_CHECK("ln_hw_partial");
  ln_hw_stats_kernel<<<uad_cdiv(B * C, 128), 128, 0, st>>>(partial, splits, HW, B * C, C, eps, mean, rstd);
  UAD_LAUNCH_CHECK("ln_hw_stats");
  const size_t n4 = (size_t)B * HW * C / 4;
  ln_hw_apply_kernel<<<uad_cdiv(n4, 256), 256, 0, st>>>(x, mean, rstd, gamma_hw, beta_hw, y, n4, HW, C, act, alpha);
  UAD_LAUNCH_CHECK("ln_hw_apply");
  return 0;
}

// y = act(x) elementwise (sigmoid / tanh heads of f-AnoGAN: models/fanogan.py:29,41,46)
__global__ void activation_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n, int act, float alpha) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = uad_act(x[i], act, alpha);
}
extern "C" int uad_activation(const float* x, float* y, size_t n, int act, float alpha, void* stream) {
  if (n == 0) return 0;
  activation_kernel<<<uad_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n, act, alpha);
  UAD_LAUNCH_CHECK("activation");
  return 0;
}
