// f-AnoGAN training kernels (reference trainers/fAnoGAN.py:50-77, models/fanogan.py:67-69, models/customlayers.py:22,35
// with use_batchnorm=False): LayerNormalization([1,2]) forward-with-statistics, its backward, its forward-mode derivative
// (JVP) and the joint backward of (primal, tangent) - the three pieces the WGAN-GP critic step needs - plus the small
// element-wise / reduction kernels of the WGAN losses.
//
// Gradient penalty without a tape: with u = dGP/d(ddx) held constant, grad_theta GP = grad_theta <u, J(theta)^T 1>
// = grad_theta <J(theta) u, 1>, i.e. the ordinary reverse pass of the directional derivative of sum(D(x_hat)) along u.
// So the critic runs: forward, reverse to x_hat (ddx), forward tangent pass with seed u, joint reverse.
#include "uad_common.cuh"

namespace {

struct LnArgs {
  const float* x;       // pre-normalisation tensor [B, HW, C]
  const float* p1;      // BWD: dy ; JVP: xdot ; BWD2: dydot (adjoint of the tangent output)
  const float* p2;      // BWD2: xdot
  const float* p3;      // BWD2: dy (adjoint of the primal output; nullable)
  const float* mean;    // [B*C]
  const float* rstd;    // [B*C]
  const float* gamma;   // [HW]
  const float* beta;    // [HW]
  int HW, C, act;
  float alpha;
};

enum { LN_STATS = 0, LN_BWD = 1, LN_JVP = 2, LN_BWD2 = 3 };

// stage 1: per (b, split) partial sums per channel over the split's pixels
//   STATS: {x, x^2}   BWD: {a, a*xh}   JVP: {xd, xd*xh}   BWD2: {a, a*xh, a*xd, a2, a2*xh}
//   with xh = (x-mean)*rstd, n = gamma*xh+beta, a = gamma*p1*act'(n), a2 = gamma*p3*act'(n)
template <int MODE, int NS>
__global__ void __launch_bounds__(256) ln_sums_kernel(LnArgs A, double* __restrict__ partial, int rows_per_split) {
  __shared__ float red[NS][256];
  const int C = A.C, HW = A.HW;
  const int c = threadIdx.x % C, pl = threadIdx.x / C, ppp = 256 / C;
  const int b = blockIdx.x, sp = blockIdx.y;
  const int r0 = sp * rows_per_split, r1 = min(HW, r0 + rows_per_split);
  float s[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) s[k] = 0.f;
  float mu = 0.f, rs = 0.f;
  if (MODE != LN_STATS) { mu = A.mean[(size_t)b * C + c]; rs = A.rstd[(size_t)b * C + c]; }
  for (int r = r0 + pl; r < r1; r += ppp) {
    const size_t idx = ((size_t)b * HW + r) * C + c;
    const float z = A.x[idx];
    if (MODE == LN_STATS) {
      s[0] += z;
      s[1] = fmaf(z, z, s[1]);
    } else {
      const float xh = (z - mu) * rs;
      if (MODE == LN_JVP) {
        const float zd = A.p1[idx];
        s[0] += zd;
        s[1] = fmaf(zd, xh, s[1]);
      } else {
        const float g = A.gamma[r];
        const float ap = uad_act_grad(fmaf(g, xh, A.beta[r]), A.act, A.alpha);
        const float a = g * A.p1[idx] * ap;
        s[0] += a;
        s[1] = fmaf(a, xh, s[1]);
        if (MODE == LN_BWD2) {
          s[2] = fmaf(a, A.p2[idx], s[2]);
          if (A.p3) {
            const float a2 = g * A.p3[idx] * ap;
            s[3] += a2;
            s[4] = fmaf(a2, xh, s[4]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NS; ++k) red[k][threadIdx.x] = s[k];
  __syncthreads();
  if (threadIdx.x < C) {
    double* dst = partial + (((size_t)b * gridDim.y + sp) * C + threadIdx.x) * NS;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      double a = 0.0;
      for (int q = 0; q < ppp; ++q) a += (double)red[k][q * C + threadIdx.x];
      dst[k] = a;
    }
  }
}

// stage 2: out[k][b*C+c] = (sum over splits) / HW ; STATS: out[0] = mean, out[1] = 1/sqrt(var+eps)
__global__ void ln_finish_kernel(const double* __restrict__ partial, int splits, int HW, int BC, int C, int NS, int stats,
                                 float eps, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BC) return;
  const int b = i / C, c = i % C;
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int s = 0; s < splits; ++s) {
    const double* src = partial + (((size_t)b * splits + s) * C + c) * NS;
    for (int k = 0; k < NS; ++k) acc[k] += src[k];
  }
  if (stats) {
    const double m = acc[0] / HW;
    double var = acc[1] / HW - m * m;
    if (var < 0.0) var = 0.0;
    out[i] = (float)m;
    out[(size_t)BC + i] = (float)(1.0 / sqrt(var + (double)eps));
  } else {
    for (int k = 0; k < NS; ++k) out[(size_t)k * BC + i] = (float)(acc[k] / HW);
  }
}

__device__ __forceinline__ float group_sum(float v, int lanes) {
  for (int o = lanes >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// stage 3 (element-wise, float4 over channels).  S = stage-2 means [NS][BC]; J = saved JVP means [2][BC] (BWD2).
//   BWD : o1 = dx = rstd*(a - S0 - xh*S1)
//   JVP : o1 = ydot = act'(n)*gamma*rstd*(xd - S0 - xh*S1)
//   BWD2: o1 = dxdot = rstd*(a - S0 - xh*S1)
//         o2 = dx    = rstd*(a2 - S3 - xh*S4) - rstd^2*(v - mean(v) - xh*mean(xh v)) - xh*rstd^2*(S2 - S0*J0 - S1*J1)
//              with v = a*J1 + xd*S1, mean(v) = J1*S0 + S1*J0, mean(xh v) = 2*S1*J1
//   per-pixel parameter partials (summed over the pixel's channels): part_g[b*HW+hw], part_b[b*HW+hw]
template <int MODE>
__global__ void __launch_bounds__(256) ln_apply_kernel(LnArgs A, const float* __restrict__ S, const float* __restrict__ J,
                                                       float* __restrict__ o1, float* __restrict__ o2,
                                                       float* __restrict__ part_g, float* __restrict__ part_b, size_t n4,
                                                       int BC) {
  const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // n4 % 256 == 0 (checked by the launcher)
  const int C = A.C, HW = A.HW;
  const size_t e = i4 * 4;
  const int c = (int)(e % C);
  const size_t pix = e / C;
  const int hw = (int)(pix % HW);
  const int b = (int)(pix / HW);
  const size_t sc = (size_t)b * C + c;
  float z[4], mu[4], rs[4], q1[4], q2[4], q3[4], s0[4], s1[4], s2[4], s3[4], s4[4], j0[4], j1[4], r1[4], r2[4];
  *reinterpret_cast<float4*>(z) = *reinterpret_cast<const float4*>(A.x + e);
  *reinterpret_cast<float4*>(mu) = *reinterpret_cast<const float4*>(A.mean + sc);
  *reinterpret_cast<float4*>(rs) = *reinterpret_cast<const float4*>(A.rstd + sc);
  *reinterpret_cast<float4*>(q1) = *reinterpret_cast<const float4*>(A.p1 + e);
  *reinterpret_cast<float4*>(s0) = *reinterpret_cast<const float4*>(S + sc);
  *reinterpret_cast<float4*>(s1) = *reinterpret_cast<const float4*>(S + (size_t)BC + sc);
  if (MODE == LN_BWD2) {
    *reinterpret_cast<float4*>(q2) = *reinterpret_cast<const float4*>(A.p2 + e);
    if (A.p3) *reinterpret_cast<float4*>(q3) = *reinterpret_cast<const float4*>(A.p3 + e);
    *reinterpret_cast<float4*>(s2) = *reinterpret_cast<const float4*>(S + 2 * (size_t)BC + sc);
    *reinterpret_cast<float4*>(s3) = *reinterpret_cast<const float4*>(S + 3 * (size_t)BC + sc);
    *reinterpret_cast<float4*>(s4) = *reinterpret_cast<const float4*>(S + 4 * (size_t)BC + sc);
    *reinterpret_cast<float4*>(j0) = *reinterpret_cast<const float4*>(J + sc);
    *reinterpret_cast<float4*>(j1) = *reinterpret_cast<const float4*>(J + (size_t)BC + sc);
  }
  const float g = A.gamma[hw], bt = A.beta[hw];
  float pg = 0.f, pb = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float xh = (z[j] - mu[j]) * rs[j];
    const float ap = uad_act_grad(fmaf(g, xh, bt), A.act, A.alpha);
    if (MODE == LN_JVP) {
      r1[j] = ap * g * rs[j] * (q1[j] - s0[j] - xh * s1[j]);
    } else {
      const float e1 = q1[j] * ap;
      const float a = g * e1;
      r1[j] = rs[j] * (a - s0[j] - xh * s1[j]);
      if (MODE == LN_BWD) {
        pg = fmaf(e1, xh, pg);
        pb += e1;
      } else {
        const float xd = q2[j];
        const float T = rs[j] * (xd - j0[j] - xh * j1[j]);
        const float v = a * j1[j] + xd * s1[j];
        const float Pv = v - (j1[j] * s0[j] + s1[j] * j0[j]) - xh * (2.f * s1[j] * j1[j]);
        const float rr = rs[j] * rs[j];
        float dx = -rr * Pv - xh * rr * (s2[j] - s0[j] * j0[j] - s1[j] * j1[j]);
        pg = fmaf(e1, T, pg);
        if (A.p3) {
          const float e2 = q3[j] * ap;
          const float a2 = g * e2;
          dx += rs[j] * (a2 - s3[j] - xh * s4[j]);
          pg = fmaf(e2, xh, pg);
          pb += e2;
        }
        r2[j] = dx;
      }
    }
  }
  *reinterpret_cast<float4*>(o1 + e) = *reinterpret_cast<const float4*>(r1);
  if (MODE == LN_BWD2) *reinterpret_cast<float4*>(o2 + e) = *reinterpret_cast<const float4*>(r2);
  if (MODE != LN_JVP && part_g) {
    const int LP = C >> 2;
    pg = group_sum(pg, LP);
    pb = group_sum(pb, LP);
    if ((threadIdx.x & (LP - 1)) == 0) { part_g[pix] = pg; part_b[pix] = pb; }
  }
}

// dgamma[hw] (+)= sum_b part_g[b*HW+hw] ; same for beta
__global__ void ln_param_reduce_kernel(const float* __restrict__ part_g, const float* __restrict__ part_b, int B, int HW,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
  const int hw = blockIdx.x * blockDim.x + threadIdx.x;
  if (hw >= HW) return;
  float sg = 0.f, sb = 0.f;
  for (int b = 0; b < B; ++b) { sg += part_g[(size_t)b * HW + hw]; sb += part_b[(size_t)b * HW + hw]; }
  dgamma[hw] = (accumulate ? dgamma[hw] : 0.f) + sg;
  dbeta[hw] = (accumulate ? dbeta[hw] : 0.f) + sb;
}

int ln_splits(int HW) { int s = HW / 256; return s < 1 ? 1 : (s > 64 ? 64 : s); }

struct LnWs { double* partial; float* means; float* part_g; float* part_b; };

size_t ln_ws_layout(void* ws, int B, int HW, int C, LnWs* L) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_partial = take((size_t)B * ln_splits(HW) * C * 5 * sizeof(double));
  const size_t o_means = take((size_t)5 * B * C * sizeof(float));
  const size_t o_pg = take((size_t)B * HW * sizeof(float));
  const size_t o_pb = take((size_t)B * HW * sizeof(float));
  if (L) {
    uint8_t* base = reinterpret_cast<uint8_t*>(ws);
    L->partial = reinterpret_cast<double*>(base + o_partial);
    L->means = reinterpret_cast<float*>(base + o_means);
    L->part_g = reinterpret_cast<float*>(base + o_pg);
    L->part_b = reinterpret_cast<float*>(base + o_pb);
  }
  return off;
}

template <int MODE, int NS>
int ln_run_sums(const LnArgs& A, int B, const LnWs& L, int stats, float eps, float* out, cudaStream_t st) {
  const int splits = ln_splits(A.HW);
  ln_sums_kernel<MODE, NS><<<dim3(B, splits), 256, 0, st>>>(A, L.partial, uad_cdiv(A.HW, splits));
  UAD_LAUNCH_CHECK("ln_sums");
  ln_finish_kernel<<<uad_cdiv(B * A.C, 128), 128, 0, st>>>(L.partial, splits, A.HW, B * A.C, A.C, NS, stats, eps, out);
  UAD_LAUNCH_CHECK("ln_finish");
  return 0;
}

int ln_check(const char* who, int B, int HW, int C, void* ws, size_t ws_bytes) {
  UAD_REQUIRE(C % 4 == 0 && C >= 8 && C <= 128 && uad_is_pow2(C), "%s: unsupported C=%d", who, C);
  UAD_REQUIRE(((size_t)B * HW * C / 4) % 256 == 0, "%s: B*HW*C must be a multiple of 1024", who);
  UAD_REQUIRE(ws && ((uintptr_t)ws % 16) == 0 && ws_bytes >= ln_ws_layout(nullptr, B, HW, C, nullptr),
              "%s: workspace too small or unaligned", who);
  return 0;
}

}  // namespace

extern "C" size_t uad_layernorm_hw_train_workspace_bytes(int B, int HW, int C) {
  return ln_ws_layout(nullptr, B, HW, C, nullptr);
}

// forward that keeps the statistics: stats[0..BC) = mean, stats[BC..2BC) = rstd
extern "C" int uad_layernorm_hw_fwd_train(const float* x, const float* gamma_hw, const float* beta_hw, float* y, float* stats,
                                          int B, int HW, int C, float eps, int act, float alpha, void* ws, size_t ws_bytes,
                                          void* stream) {
  if (int rc = ln_check("uad_layernorm_hw_fwd_train", B, HW, C, ws, ws_bytes)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  LnWs L;
  ln_ws_layout(ws, B, HW, C, &L);
  LnArgs A{x, nullptr, nullptr, nullptr, nullptr, nullptr, gamma_hw, beta_hw, HW, C, act, alpha};
  if (int rc = ln_run_sums<LN_STATS, 2>(A, B, L, 1, eps, stats, st)) return rc;
  return uad_layernorm_hw_apply(x, stats, stats + (size_t)B * C, gamma_hw, beta_hw, y, B, HW, C, act, alpha, st);
}

extern "C" int uad_layernorm_hw_bwd(const float* dy, const float* x, const float* stats, const float* gamma_hw,
                                    const float* beta_hw, float* dx, float* dgamma, float* dbeta, int B, int HW, int C,
                                    int act, float alpha, int accumulate, void* ws, size_t ws_bytes, void* stream) {
  if (int rc = ln_check("uad_layernorm_hw_bwd", B, HW, C, ws, ws_bytes)) return rc;
  UAD_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "uad_layernorm_hw_bwd: dgamma/dbeta must both be set or both NULL");
  cudaStream_t st = (cudaStream_t)stream;
  LnWs L;
  ln_ws_layout(ws, B, HW, C, &L);
  const int BC = B * C;
  LnArgs A{x, dy, nullptr, nullptr, stats, stats + BC, gamma_hw, beta_hw, HW, C, act, alpha};
  if (int rc = ln_run_sums<LN_BWD, 2>(A, B, L, 0, 0.f, L.means, st)) return rc;
  const size_t n4 = (size_t)B * HW * C / 4;
  ln_apply_kernel<LN_BWD><<<(unsigned)(n4 / 256), 256, 0, st>>>(A, L.means, nullptr, dx, nullptr, dgamma ? L.part_g : nullptr,
                                                               L.part_b, n4, BC);
  UAD_LAUNCH_CHECK("ln_apply_bwd");
  if (dgamma) {
    ln_param_reduce_kernel<<<uad_cdiv(HW, 128), 128, 0, st>>>(L.part_g, L.part_b, B, HW, dgamma, dbeta, accumulate);
    UAD_LAUNCH_CHECK("ln_param_reduce");
  }
  return 0;
}

// forward-mode derivative: ydot = d/de act(LN(x + e*xdot)) ; jstats[0..BC) = mean(xdot), jstats[BC..2BC) = mean(xdot*xh)
extern "C" int uad_layernorm_hw_jvp(const float* xdot, const float* x, const float* stats, const float* gamma_hw,
                                    const float* beta_hw, float* ydot, float* jstats, int B, int HW, int C, int act,
                                    float alpha, void* ws, size_t ws_bytes, void* stream) {
  if (int rc = ln_check("uad_layernorm_hw_jvp", B, HW, C, ws, ws_bytes)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  LnWs L;
  ln_ws_layout(ws, B, HW, C, &L);
  const int BC = B * C;
  LnArgs A{x, xdot, nullptr, nullptr, stats, stats + BC, gamma_hw, beta_hw, HW, C, act, alpha};
  if (int rc = ln_run_sums<LN_JVP, 2>(A, B, L, 0, 0.f, jstats, st)) return rc;
  const size_t n4 = (size_t)B * HW * C / 4;
  ln_apply_kernel<LN_JVP><<<(unsigned)(n4 / 256), 256, 0, st>>>(A, jstats, nullptr, ydot, nullptr, nullptr, nullptr, n4, BC);
  UAD_LAUNCH_CHECK("ln_apply_jvp");
  return 0;
}

// joint reverse of (y, ydot) = (act(LN(x)), JVP): given the adjoints dydot (and dy, nullable) produce dxdot, dx and the
// gamma / beta gradients.
extern "C" int uad_layernorm_hw_bwd2(const float* dydot, const float* dy, const float* x, const float* xdot,
                                     const float* stats, const float* jstats, const float* gamma_hw, const float* beta_hw,
                                     float* dxdot, float* dx, float* dgamma, float* dbeta, int B, int HW, int C, int act,
                                     float alpha, int accumulate, void* ws, size_t ws_bytes, void* stream) {
  if (int rc = ln_check("uad_layernorm_hw_bwd2", B, HW, C, ws, ws_bytes)) return rc;
  UAD_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "uad_layernorm_hw_bwd2: dgamma/dbeta must both be set or both NULL");
  cudaStream_t st = (cudaStream_t)stream;
  LnWs L;
  ln_ws_layout(ws, B, HW, C, &L);
  const int BC = B * C;
  LnArgs A{x, dydot, xdot, dy, stats, stats + BC, gamma_hw, beta_hw, HW, C, act, alpha};
  if (int rc = ln_run_sums<LN_BWD2, 5>(A, B, L, 0, 0.f, L.means, st)) return rc;
  const size_t n4 = (size_t)B * HW * C / 4;
  ln_apply_kernel<LN_BWD2><<<(unsigned)(n4 / 256), 256, 0, st>>>(A, L.means, jstats, dxdot, dx, dgamma ? L.part_g : nullptr,
                                                                L.part_b, n4, BC);
  UAD_LAUNCH_CHECK("ln_apply_bwd2");
  if (dgamma) {
    ln_param_reduce_kernel<<<uad_cdiv(HW, 128), 128, 0, st>>>(L.part_g, L.part_b, B, HW, dgamma, dbeta, accumulate);
    UAD_LAUNCH_CHECK("ln_param_reduce");
  }
  return 0;
}

// ================================================================================================ WGAN-GP element-wise
namespace {

__global__ void fill_kernel(float* __restrict__ y, float v, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = v;
}

// x_hat = x + alpha[b]*(x_ - x)   (models/fanogan.py:67-69)
__global__ void interpolate_kernel(const float* __restrict__ x, const float* __restrict__ xg, const float* __restrict__ alpha,
                                   float* __restrict__ out, size_t n, size_t per) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[i] + alpha[i / per] * (xg[i] - x[i]);
}

__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ u, float* __restrict__ dx, size_t n,
                               int act, float alpha) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = dy[i] * uad_act_grad(u[i], act, alpha);
}

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
  v = uad_warp_sum_d(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    t = uad_warp_sum_d(t);
  }
  __syncthreads();
  return t;     // valid in warp 0
}

// two-stage deterministic reductions.  MODE 0: sum x ; MODE 1: sum (a-b)^2 with optional grad = gscale*(a-b)
template <int MODE>
__global__ void __launch_bounds__(256) reduce_stage1_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            float gscale, float* __restrict__ grad, size_t n,
                                                            double* __restrict__ partial) {
  __shared__ double sh[8];
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    if (MODE == 0) {
      s += a[i];
    } else {
      const float d = a[i] - b[i];
      s = fmaf(d, d, s);
      if (grad) grad[i] = gscale * d;
    }
  }
  const double t = block_sum_d((double)s, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(256) reduce_stage2_kernel(const double* __restrict__ partial, int nb, double scale,
                                                            float* __restrict__ out) {
  __shared__ double sh[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) s += partial[i];
  const double t = block_sum_d(s, sh);
  if (threadIdx.x == 0) out[0] = (float)(t * scale);
}

int reduce_blocks(size_t n) {
  long long nb = (long long)((n + 256 * 16 - 1) / (256 * 16));
  if (nb < 1) nb = 1;
  if (nb > 4 * UAD_NUM_SMS) nb = 4 * UAD_NUM_SMS;
  return (int)nb;
}

// slope[b, j] = sqrt(sum_h ddx[b, h, j]^2) : the reference reduces over axis 1 ONLY (trainers/fAnoGAN.py:56)
__global__ void gp_slope_kernel(const float* __restrict__ ddx, float* __restrict__ slope, int B, int H, int WC) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * WC) return;
  const int b = i / WC, j = i % WC;
  const float* p = ddx + (size_t)b * H * WC + j;
  float s = 0.f;
  for (int h = 0; h < H; ++h) { const float v = p[(size_t)h * WC]; s = fmaf(v, v, s); }
  slope[i] = sqrtf(s);
}

// u = dGP/d(ddx) = scale * 2*(slope-1)/(B*WC) * ddx/slope   (trainers/fAnoGAN.py:56-57)
__global__ void gp_seed_kernel(const float* __restrict__ ddx, const float* __restrict__ slope, float* __restrict__ u, int B,
                               int H, int WC, float coef) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)B * H * WC;
  if (i >= n) return;
  const int j = (int)(i % WC);
  const int b = (int)(i / ((size_t)H * WC));
  const float sl = slope[(size_t)b * WC + j];
  u[i] = sl > 0.f ? coef * (sl - 1.f) / sl * ddx[i] : 0.f;
}

__global__ void __launch_bounds__(256) gp_value_kernel(const float* __restrict__ slope, int n, double scale,
                                                       float* __restrict__ out) {
  __shared__ double sh[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) { const double d = (double)slope[i] - 1.0; s += d * d; }
  const double t = block_sum_d(s, sh);
  if (threadIdx.x == 0) out[0] = (float)(t * scale);
}

// l1 = |xhat - x| ; rec[b] = sum_hw l1   (trainers/fAnoGAN.py:65-66; trainers/AE.py:28-29 behind the fused 1x1 head).
// One block of 1024 threads per sample, 16-byte accesses when HW % 4 == 0 (64 blocks of 256 scalar threads ran at 1.3 TB/s).
__global__ void __launch_bounds__(1024) l1_map_kernel(const float* __restrict__ x, const float* __restrict__ xhat,
                                                      float* __restrict__ l1, float* __restrict__ rec, int HW) {
  __shared__ double sh[32];
  const int b = blockIdx.x;
  const size_t base = (size_t)b * HW;
  float s = 0.f;
  if ((HW & 3) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(x + base);
    const float4* h4 = reinterpret_cast<const float4*>(xhat + base);
    float4* l4 = l1 ? reinterpret_cast<float4*>(l1 + base) : nullptr;
    for (int i = threadIdx.x; i < HW / 4; i += blockDim.x) {
      const float4 a = h4[i], c = x4[i];
      const float4 d = make_float4(fabsf(a.x - c.x), fabsf(a.y - c.y), fabsf(a.z - c.z), fabsf(a.w - c.w));
      if (l4) l4[i] = d;
      s += (d.x + d.y) + (d.z + d.w);
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const float d = fabsf(xhat[base + i] - x[base + i]);
      if (l1) l1[base + i] = d;
      s += d;
    }
  }
  const double t = block_sum_d((double)s, sh);
  if (threadIdx.x == 0 && rec) rec[b] = (float)t;
}

}  // namespace

extern "C" int uad_fill(float* y, float v, size_t n, void* stream) {
  if (n == 0) return 0;
  fill_kernel<<<uad_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(y, v, n);
  UAD_LAUNCH_CHECK("fill");
  return 0;
}

extern "C" int uad_interpolate(const float* x, const float* x_gen, const float* alpha, float* out, int B, size_t per_sample,
                               void* stream) {
  const size_t n = (size_t)B * per_sample;
  if (n == 0) return 0;
  interpolate_kernel<<<uad_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, x_gen, alpha, out, n, per_sample);
  UAD_LAUNCH_CHECK("interpolate");
  return 0;
}

extern "C" int uad_activation_bwd(const float* dy, const float* u, float* dx, size_t n, int act, float alpha, void* stream) {
  if (n == 0) return 0;
  act_bwd_kernel<<<uad_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, u, dx, n, act, alpha);
  UAD_LAUNCH_CHECK("activation_bwd");
  return 0;
}

extern "C" size_t uad_reduce_workspace_bytes(void) { return (size_t)4 * UAD_NUM_SMS * sizeof(double); }

extern "C" int uad_sum_scaled(const float* x, size_t n, double scale, float* out_dev, void* ws, size_t ws_bytes, void* stream) {
  UAD_REQUIRE(ws && ws_bytes >= uad_reduce_workspace_bytes(), "uad_sum_scaled: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = reduce_blocks(n);
  reduce_stage1_kernel<0><<<nb, 256, 0, st>>>(x, nullptr, 0.f, nullptr, n, (double*)ws);
  UAD_LAUNCH_CHECK("sum_stage1");
  reduce_stage2_kernel<<<1, 256, 0, st>>>((const double*)ws, nb, scale, out_dev);
  UAD_LAUNCH_CHECK("sum_stage2");
  return 0;
}

extern "C" int uad_mse(const float* a, const float* b, size_t n, float grad_scale, float* grad_a, double loss_scale,
                       float* loss_out_dev, void* ws, size_t ws_bytes, void* stream) {
  UAD_REQUIRE(ws && ws_bytes >= uad_reduce_workspace_bytes(), "uad_mse: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = reduce_blocks(n);
  reduce_stage1_kernel<1><<<nb, 256, 0, st>>>(a, b, grad_scale, grad_a, n, (double*)ws);
  UAD_LAUNCH_CHECK("mse_stage1");
  if (loss_out_dev) {
    reduce_stage2_kernel<<<1, 256, 0, st>>>((const double*)ws, nb, loss_scale, loss_out_dev);
    UAD_LAUNCH_CHECK("mse_stage2");
  }
  return 0;
}

extern "C" int uad_gradient_penalty(const float* ddx, int B, int H, int WC, float scale, float* u_out, float* gp_out_dev,
                                    void* ws, size_t ws_bytes, void* stream) {
  UAD_REQUIRE(ws && ws_bytes >= (size_t)B * WC * sizeof(float), "uad_gradient_penalty: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* slope = reinterpret_cast<float*>(ws);
  gp_slope_kernel<<<uad_cdiv((long long)B * WC, 128), 128, 0, st>>>(ddx, slope, B, H, WC);
  UAD_LAUNCH_CHECK("gp_slope");
  const double inv = 1.0 / ((double)B * WC);
  if (u_out) {
    gp_seed_kernel<<<uad_cdiv((long long)B * H * WC, 256), 256, 0, st>>>(ddx, slope, u_out, B, H, WC, (float)(2.0 * scale * inv));
    UAD_LAUNCH_CHECK("gp_seed");
  }
  if (gp_out_dev) {
    gp_value_kernel<<<1, 256, 0, st>>>(slope, B * WC, (double)scale * inv, gp_out_dev);
    UAD_LAUNCH_CHECK("gp_value");
  }
  return 0;
}

extern "C" int uad_l1_map(const float* x, const float* xhat, float* l1, float* rec, int B, int HW, void* stream) {
  if (B == 0) return 0;
  l1_map_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(x, xhat, l1, rec, HW);
  UAD_LAUNCH_CHECK("l1_map");
  return 0;
}
