// fp32 SIMT (FFMA) implicit-GEMM kernels for the k x k, stride-2, SAME conv / transposed-conv family.
//
// Three GEMM forms cover all six ops (see DESIGN.md "conv forms"):
//   Form F (gather, stride 2):   conv fwd, convT dgrad       M = B*Ho*Wo pixels, N = out channels, K = taps*Cin
//   Form T (parity-decomposed):  convT fwd, conv dgrad       4 output-parity classes, each a stride-1 gather GEMM
//   Form W (pixel reduction):    conv wgrad, convT wgrad     M' = taps*Cg, N' = Co, K' = B*Ho*Wo pixels, split-K
// plus direct kernels for the Cin == 1 first layer.
//
// This is the exact-fp32 path (math_mode UAD_MATH_FP32_SIMT).  It is also the on-device cross-check for the tcgen05
// path in uad_conv_tc.cu.
#include "uad_conv.cuh"

// ------------------------------------------------------------------------------------------------ Form F / Form T
template <int BM, int BN>
__global__ void __launch_bounds__(BM* BN / 64) gather_gemm_simt(const __grid_constant__ GatherParams p) {
  constexpr int BK = 16;
  constexpr int NT = BM * BN / 64;
  constexpr int LA = BM * 4 / NT;
  constexpr int LB = BK * BN / 4 / NT;
  static_assert(NT % 4 == 0 && LB >= 1, "tile config");
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const TapSet& ts = p.taps[blockIdx.z];
  const int MWm = (1 << p.lgMW) - 1, MHm = (1 << p.lgMH) - 1;

  // --- per-thread A slots (fixed rows across the K loop)
  const int k4 = tid & 3;
  int pb[LA], ih0[LA], iw0[LA];
#pragma unroll
  for (int j = 0; j < LA; ++j) {
    int m = m0 + (tid >> 2) + j * (NT / 4);
    if (m < p.M) {
      int s = m & MWm, r = (m >> p.lgMW) & MHm, b = m >> (p.lgMW + p.lgMH);
      pb[j] = b * p.IH * p.IW;
      ih0[j] = r * p.sh;
      iw0[j] = s * p.sh;
    } else {
      pb[j] = 0; ih0[j] = -100000; iw0[j] = -100000;
    }
  }
  const int cpb = p.Cin / BK;          // K blocks per tap
  const int nkb = ts.n * cpb;

  float4 ra[LA], rb[LB];
  auto load_regs = [&](int tap, int ci0) {
    const int dh = ts.dh[tap], dw = ts.dw[tap], wt = ts.wt[tap];
#pragma unroll
    for (int j = 0; j < LA; ++j) {
      int ih = ih0[j] + dh, iw = iw0[j] + dw;
      bool ok = (unsigned)ih < (unsigned)p.IH && (unsigned)iw < (unsigned)p.IW;
      if (ok) {
        size_t off = ((size_t)(pb[j] + ih * p.IW + iw)) * p.Cin + ci0 + k4 * 4;
        ra[j] = __ldg(reinterpret_cast<const float4*>(p.in + off));
      } else {
        ra[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int j = 0; j < LB; ++j) {
      int idx = tid + j * NT;
      int kr = idx / (BN / 4), n4 = idx % (BN / 4);
      size_t off = ((size_t)(wt * p.Cin + ci0 + kr)) * p.N + n0 + n4 * 4;
      rb[j] = __ldg(reinterpret_cast<const float4*>(p.wmat + off));
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int j = 0; j < LA; ++j) {
      int row = (tid >> 2) + j * (NT / 4);
      As[buf][k4 * 4 + 0][row] = ra[j].x;
      As[buf][k4 * 4 + 1][row] = ra[j].y;
      As[buf][k4 * 4 + 2][row] = ra[j].z;
      As[buf][k4 * 4 + 3][row] = ra[j].w;
    }
#pragma unroll
    for (int j = 0; j < LB; ++j) {
      int idx = tid + j * NT;
      int kr = idx / (BN / 4), n4 = idx % (BN / 4);
      *reinterpret_cast<float4*>(&Bs[buf][kr][n4 * 4]) = rb[j];
    }
  };

  const int tx = tid % (BN / 8), ty = tid / (BN / 8);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  int tap = 0, cb = 0;
  load_regs(0, 0);
  store_smem(0);
  __syncthreads();
  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    int ntap = tap, ncb = cb + 1;
    if (ncb == cpb) { ncb = 0; ++ntap; }
    const bool more = (kb + 1 < nkb);
    if (more) load_regs(ntap, ncb * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][BM / 2 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][BN / 2 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) store_smem(buf ^ 1);
    __syncthreads();
    tap = ntap; cb = ncb;
  }

  // --- epilogue: z = acc + bias ; a = act(gamma*bn_c*z + beta)
  float bia[8], sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int n = n0 + (j < 4 ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4));
    bia[j] = p.bias ? __ldg(p.bias + n) : 0.f;
    sc[j] = p.gamma ? __ldg(p.gamma + n) * p.bn_c : 1.f;
    sf[j] = p.beta ? __ldg(p.beta + n) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
    int s = m & MWm, r = (m >> p.lgMW) & MHm, b = m >> (p.lgMW + p.lgMH);
    size_t opix = ((size_t)b * p.OH + (r * p.osh + ts.oh0)) * p.OW + (s * p.osh + ts.ow0);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int n = n0 + (h == 0 ? tx * 4 : BN / 2 + tx * 4);
      float z[4], a[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        z[j] = acc[i][h * 4 + j] + bia[h * 4 + j];
        a[j] = uad_act(sc[h * 4 + j] * z[j] + sf[h * 4 + j], p.act, p.alpha);
      }
      if (p.z_out) *reinterpret_cast<float4*>(p.z_out + opix * p.N + n) = make_float4(z[0], z[1], z[2], z[3]);
      if (p.a_out) *reinterpret_cast<float4*>(p.a_out + opix * p.N + n) = make_float4(a[0], a[1], a[2], a[3]);
    }
  }
}

int uad_launch_gather_simt(const GatherParams& p, int nclasses, cudaStream_t st) {
  UAD_REQUIRE(p.Cin % 16 == 0, "gather_gemm_simt: Cin=%d must be a multiple of 16", p.Cin);
  if (p.N % 64 == 0) {
    dim3 g(uad_cdiv(p.M, 128), p.N / 64, nclasses);
    gather_gemm_simt<128, 64><<<g, 128, 0, st>>>(p);
  } else if (p.N % 32 == 0) {
    dim3 g(uad_cdiv(p.M, 256), p.N / 32, nclasses);
    gather_gemm_simt<256, 32><<<g, 128, 0, st>>>(p);
  } else {
    return uad_set_error("gather_gemm_simt: N=%d must be a multiple of 32", p.N);
  }
  UAD_LAUNCH_CHECK("gather_gemm_simt");
  return 0;
}

// ------------------------------------------------------------------------------------------------ Form W (wgrad)
template <int BM, int BN>
__global__ void __launch_bounds__(BM* BN / 64) wgrad_simt(const __grid_constant__ WgradParams p) {
  constexpr int BK = 16;
  constexpr int NT = BM * BN / 64;
  constexpr int LA = BK * BM / 4 / NT;
  constexpr int LB = BK * BN / 4 / NT;
  static_assert(NT % (BM / 4) == 0 && NT % (BN / 4) == 0 && LB >= 1, "tile config");
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int MWm = (1 << p.lgMW) - 1, MHm = (1 << p.lgMH) - 1;
  const int pix_begin = blockIdx.z * p.chunk;
  const int pix_end = min(p.P, pix_begin + p.chunk);
  const int nkb = (pix_end - pix_begin + BK - 1) / BK;

  // fixed m' column group of this thread
  const int m4 = tid % (BM / 4);
  const int mp = m0 + m4 * 4;
  const bool mvalid = mp < p.Mp;
  int tdh = 0, tdw = 0, cg = 0;
  if (mvalid) {
    int t = mp / p.Cg;
    cg = mp - t * p.Cg;
    tdh = p.taps.dh[t];
    tdw = p.taps.dw[t];
  }
  const int ka = tid / (BM / 4);             // first pixel row of this thread in the A tile
  const int n4 = tid % (BN / 4);
  const int kb0 = tid / (BN / 4);

  float4 ra[LA], rb[LB];
  auto load_regs = [&](int kb) {
    const int pbase = pix_begin + kb * BK;
#pragma unroll
    for (int j = 0; j < LA; ++j) {
      int pix = pbase + ka + j * (NT / (BM / 4));
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mvalid && pix < pix_end) {
        int s = pix & MWm, r = (pix >> p.lgMW) & MHm, b = pix >> (p.lgMW + p.lgMH);
        int ih = r * p.sh + tdh, iw = s * p.sh + tdw;
        if ((unsigned)ih < (unsigned)p.GH && (unsigned)iw < (unsigned)p.GW) {
          size_t off = ((size_t)(b * p.GH + ih) * p.GW + iw) * p.Cg + cg;
          v = __ldg(reinterpret_cast<const float4*>(p.g + off));
        }
      }
      ra[j] = v;
    }
#pragma unroll
    for (int j = 0; j < LB; ++j) {
      int pix = pbase + kb0 + j * (NT / (BN / 4));
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pix < pix_end) v = __ldg(reinterpret_cast<const float4*>(p.o + (size_t)pix * p.Co + n0 + n4 * 4));
      rb[j] = v;
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int j = 0; j < LA; ++j)
      *reinterpret_cast<float4*>(&As[buf][ka + j * (NT / (BM / 4))][m4 * 4]) = ra[j];
#pragma unroll
    for (int j = 0; j < LB; ++j)
      *reinterpret_cast<float4*>(&Bs[buf][kb0 + j * (NT / (BN / 4))][n4 * 4]) = rb[j];
  };

  const int tx = tid % (BN / 8), ty = tid / (BN / 8);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  if (nkb > 0) {
    load_regs(0);
    store_smem(0);
  }
  __syncthreads();
  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    const bool more = (kb + 1 < nkb);
    if (more) load_regs(kb + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][BM / 2 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][BN / 2 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) store_smem(buf ^ 1);
    __syncthreads();
  }

  float* out = p.partial + (size_t)blockIdx.z * p.Mp * p.Co;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4));
    if (m >= p.Mp) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int n = n0 + (h == 0 ? tx * 4 : BN / 2 + tx * 4);
      *reinterpret_cast<float4*>(out + (size_t)m * p.Co + n) =
          make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
    }
  }
}

// deterministic split-K reduction.  A thread per output walking `splits` partials serially is latency bound when there are many
// (292 dependent iterations for the largest layer: 26 us for 30 MB): from 32 splits on, a block = 32 outputs x 8 split groups -
// every thread sums every 8th partial (coalesced 128-byte rows), the 8 group sums are added in a fixed order.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, int splits, size_t n,
                                                            float* __restrict__ out, int accumulate) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = accumulate ? out[i] : 0.f;
  for (int k = 0; k < splits; ++k) s += partial[(size_t)k * n + i];
  out[i] = s;
}

__global__ void __launch_bounds__(256) splitk_reduce8_kernel(const float* __restrict__ partial, int splits, size_t n,
                                                             float* __restrict__ out, int accumulate) {
  __shared__ float sh[8][32];
  const int o = threadIdx.x & 31, g = threadIdx.x >> 5;
  const size_t i = (size_t)blockIdx.x * 32 + o;
  float s = 0.f;
  if (i < n)
    for (int k = g; k < splits; k += 8) s += partial[(size_t)k * n + i];
  sh[g][o] = s;
  __syncthreads();
  if (g == 0 && i < n) {
    float t = accumulate ? out[i] : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][o];
    out[i] = t;
  }
}

static void launch_splitk_reduce(const float* partial, int splits, size_t n, float* out, int accumulate, cudaStream_t st) {
  if (splits >= 32) splitk_reduce8_kernel<<<uad_cdiv(n, 32), 256, 0, st>>>(partial, splits, n, out, accumulate);
  else splitk_reduce_kernel<<<uad_cdiv(n, 256), 256, 0, st>>>(partial, splits, n, out, accumulate);
}

int uad_wgrad_plan(int Mp, int Co, int P, int* splits, int* chunk) {
  const int BN = (Co % 64 == 0) ? 64 : 32;
  const int BM = (BN == 64) ? 128 : 256;
  int tiles = uad_cdiv(Mp, BM) * (Co / BN);
  int s = uad_cdiv(2 * UAD_NUM_SMS, tiles);
  int maxs = P / 64 > 0 ? P / 64 : 1;
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  int c = uad_cdiv(uad_cdiv(P, s), 16) * 16;
  *splits = uad_cdiv(P, c);
  *chunk = c;
  return 0;
}

int uad_launch_wgrad_simt(WgradParams p, float* out, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  UAD_REQUIRE(p.Cg % 4 == 0 && p.Co % 32 == 0, "wgrad_simt: Cg=%d %%4, Co=%d %%32 required", p.Cg, p.Co);
  int splits, chunk;
  uad_wgrad_plan(p.Mp, p.Co, p.P, &splits, &chunk);
  size_t need = (size_t)splits * p.Mp * p.Co * sizeof(float);
  UAD_REQUIRE(ws && ws_bytes >= need, "wgrad_simt: workspace too small (%zu < %zu)", ws_bytes, need);
  p.partial = reinterpret_cast<float*>(ws);
  p.chunk = chunk;
  if (p.Co % 64 == 0) {
    dim3 g(uad_cdiv(p.Mp, 128), p.Co / 64, splits);
    wgrad_simt<128, 64><<<g, 128, 0, st>>>(p);
  } else {
    dim3 g(uad_cdiv(p.Mp, 256), p.Co / 32, splits);
    wgrad_simt<256, 32><<<g, 128, 0, st>>>(p);
  }
  UAD_LAUNCH_CHECK("wgrad_simt");
  size_t n = (size_t)p.Mp * p.Co;
  launch_splitk_reduce(p.partial, splits, n, out, accumulate, st);
  UAD_LAUNCH_CHECK("splitk_reduce");
  return 0;
}

int uad_launch_splitk_reduce(const float* partial, int splits, size_t n, float* out, int accumulate, cudaStream_t st) {
  launch_splitk_reduce(partial, splits, n, out, accumulate, st);
  UAD_LAUNCH_CHECK("splitk_reduce");
  return 0;
}

// ------------------------------------------------------------------------------------------------ weight prep
// w[t][A][Bd] -> wT[t][Bd][A]
__global__ void transpose_taps_kernel(const float* __restrict__ w, float* __restrict__ wT, int T, int A, int Bd) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)T * A * Bd;
  if (i >= n) return;
  int b = i % Bd;
  int a = (i / Bd) % A;
  int t = i / ((size_t)A * Bd);
  wT[((size_t)t * Bd + b) * A + a] = w[i];
}

int uad_launch_transpose_taps(const float* w, float* wT, int T, int A, int Bd, cudaStream_t st) {
  size_t n = (size_t)T * A * Bd;
  transpose_taps_kernel<<<uad_cdiv(n, 256), 256, 0, st>>>(w, wT, T, A, Bd);
  UAD_LAUNCH_CHECK("transpose_taps");
  return 0;
}

// ------------------------------------------------------------------------------------------------ Cin == 1 first layer
// x [B,H,W,1] -> [B,H/2,W/2,Cout]; block = R consecutive output rows of one image; thread = (output row, 4-pixel strip, 8-channel
// group).  The filter and the 2 R + 3 input rows the block needs are staged once (row-wise, no per-element div / mod); a filter
// row's 11 inputs of a strip are three 16-byte shared-memory reads (the strip starts at a multiple of 8 floats of a row whose pitch
// is a multiple of 4), a tap's eight weights two.  One output row per block (the first version) paid the staging, two barriers and
// the tail once per 128 pixels: 0.10 ms for 0.15 GB of HBM traffic.
template <int K, int R>
__global__ void __launch_bounds__(256) conv_c1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ z_out,
                                                          float* __restrict__ a_out, int B, int H, int W, int Cout, int pad_lo,
                                                          int act, float alpha, float bn_c) {
  extern __shared__ __align__(16) float smem[];
  const int Ho = H / 2, Wo = W / 2;
  const int XW = W + 16;                     // row pitch: pad_lo zeros, W inputs, zeros up to a multiple of 4 (>= 2 Wo + 11 - 1 reads)
  constexpr int NR = 2 * R + K - 2;          // input rows of R output rows
  float* ws = smem;                          // [K*K][Cout]
  float* xs = smem + K * K * Cout;           // [NR][XW]
  const int blocks_per_img = Ho / R;
  const int b = blockIdx.x / blocks_per_img, oh0 = (blockIdx.x % blocks_per_img) * R;
  for (int i = threadIdx.x; i < K * K * Cout; i += blockDim.x) ws[i] = w[i];
  for (int r = threadIdx.x / 64; r < NR; r += blockDim.x / 64) {          // 64 threads per input row
    const int ih = 2 * oh0 + r - pad_lo;
    const bool in = (unsigned)ih < (unsigned)H;
    const float* src = x + ((size_t)b * H + (in ? ih : 0)) * W;
    for (int c = threadIdx.x % 64; c < XW; c += 64) {
      const int iw = c - pad_lo;
      xs[r * XW + c] = (in && (unsigned)iw < (unsigned)W) ? __ldg(src + iw) : 0.f;
    }
  }
  __syncthreads();
  const int ngrp = Cout / 8;
  const int g = threadIdx.x % ngrp;
  const int strips = Wo / 4;                 // 4-pixel strips per output row
  float bia[8], sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int n = g * 8 + j;
    bia[j] = bias ? bias[n] : 0.f;
    sc[j] = gamma ? gamma[n] * bn_c : 1.f;
    sf[j] = beta ? beta[n] : 0.f;
  }
  for (int item = threadIdx.x / ngrp; item < R * strips; item += blockDim.x / ngrp) {
    const int rr = item / strips, ow0 = 4 * (item % strips);
    float acc[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
#pragma unroll 1
    for (int kh = 0; kh < K; ++kh) {
      float xr[12];
      const float4* xrow = reinterpret_cast<const float4*>(xs + (2 * rr + kh) * XW + 2 * ow0);
#pragma unroll
      for (int i = 0; i < 3; ++i) { const float4 t = xrow[i]; xr[4 * i] = t.x; xr[4 * i + 1] = t.y; xr[4 * i + 2] = t.z; xr[4 * i + 3] = t.w; }
#pragma unroll
      for (int kw = 0; kw < K; ++kw) {
        const float4 w0 = *reinterpret_cast<const float4*>(&ws[(kh * K + kw) * Cout + g * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&ws[(kh * K + kw) * Cout + g * 8 + 4]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float xv = xr[2 * q + kw];
          acc[q][0] = fmaf(xv, w0.x, acc[q][0]); acc[q][1] = fmaf(xv, w0.y, acc[q][1]);
          acc[q][2] = fmaf(xv, w0.z, acc[q][2]); acc[q][3] = fmaf(xv, w0.w, acc[q][3]);
          acc[q][4] = fmaf(xv, w1.x, acc[q][4]); acc[q][5] = fmaf(xv, w1.y, acc[q][5]);
          acc[q][6] = fmaf(xv, w1.z, acc[q][6]); acc[q][7] = fmaf(xv, w1.w, acc[q][7]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      size_t o = (((size_t)b * Ho + oh0 + rr) * Wo + ow0 + q) * Cout + g * 8;
      float z[8], a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        z[j] = acc[q][j] + bia[j];
        a[j] = uad_act(sc[j] * z[j] + sf[j], act, alpha);
      }
      if (z_out) {
        *reinterpret_cast<float4*>(z_out + o) = make_float4(z[0], z[1], z[2], z[3]);
        *reinterpret_cast<float4*>(z_out + o + 4) = make_float4(z[4], z[5], z[6], z[7]);
      }
      if (a_out) {
        *reinterpret_cast<float4*>(a_out + o) = make_float4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<float4*>(a_out + o + 4) = make_float4(a[4], a[5], a[6], a[7]);
      }
    }
  }
}

int uad_launch_conv_c1_fwd(const float* x, const float* w, const float* bias, const float* gamma, const float* beta,
                           float* z_out, float* a_out, int B, int H, int W, int Cout, int ksize, int act, float alpha,
                           float bn_c, cudaStream_t st) {
  UAD_REQUIRE(ksize == 5, "conv_c1_fwd: only k=5 (got %d)", ksize);
  UAD_REQUIRE(Cout % 8 == 0 && Cout <= 128 && 256 % (Cout / 8) == 0, "conv_c1_fwd: unsupported Cout=%d", Cout);
  UAD_REQUIRE(W % 8 == 0 && H % 2 == 0, "conv_c1_fwd: W=%d must be a multiple of 8, H=%d even", W, H);
  const int pad_lo = (ksize - 2) / 2;
  const int Ho = H / 2;
  const int R = Ho % 4 == 0 ? 4 : (Ho % 2 == 0 ? 2 : 1);
  size_t smem = ((size_t)ksize * ksize * Cout + (size_t)(2 * R + ksize - 2) * (W + 16)) * sizeof(float);
  UAD_REQUIRE(smem <= 48 * 1024, "conv_c1_fwd: W=%d too large", W);
  const int grid = B * (Ho / R);
  if (R == 4) conv_c1_fwd_kernel<5, 4><<<grid, 256, smem, st>>>(x, w, bias, gamma, beta, z_out, a_out, B, H, W, Cout, pad_lo, act, alpha, bn_c);
  else if (R == 2) conv_c1_fwd_kernel<5, 2><<<grid, 256, smem, st>>>(x, w, bias, gamma, beta, z_out, a_out, B, H, W, Cout, pad_lo, act, alpha, bn_c);
  else conv_c1_fwd_kernel<5, 1><<<grid, 256, smem, st>>>(x, w, bias, gamma, beta, z_out, a_out, B, H, W, Cout, pad_lo, act, alpha, bn_c);
  UAD_LAUNCH_CHECK("conv_c1_fwd");
  return 0;
}

// dw[t][co] = sum_pix x[pix + t] * dz[pix][co];  grid-stride over output rows, per-block partials
template <int K, int Q>
__global__ void __launch_bounds__(256) conv_c1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                            float* __restrict__ partial, int B, int H, int W, int Cout,
                                                            int pad_lo) {
  extern __shared__ __align__(16) float smem[];
  const int Ho = H / 2, Wo = W / 2;
  const int XW = W + 24;                      // zero padded, pitch a multiple of 4 floats: a filter row's inputs of a strip are 16-byte reads
  float* xs = smem;                           // [K][XW]
  float* red = smem + K * XW;                 // [nwarps][K*K*Cout]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float acc[K * K][Q];
#pragma unroll
  for (int t = 0; t < K * K; ++t)
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[t][q] = 0.f;

  for (int row = blockIdx.x; row < B * Ho; row += gridDim.x) {
    const int b = row / Ho, oh = row % Ho;
    __syncthreads();
    for (int kh = warp; kh < K; kh += nwarps) {                   // one warp per input row: no per-element div / mod
      const int ih = 2 * oh + kh - pad_lo;
      const bool in = (unsigned)ih < (unsigned)H;
      const float* src = x + ((size_t)b * H + (in ? ih : 0)) * W;
      for (int c = lane; c < XW; c += 32) {
        const int iw = c - pad_lo;
        xs[kh * XW + c] = (in && (unsigned)iw < (unsigned)W) ? __ldg(src + iw) : 0.f;
      }
    }
    __syncthreads();
    constexpr int U = 8;                                          // pixels per iteration: 8 independent dz loads in flight, and a
    for (int ow0 = warp * U; ow0 < Wo; ow0 += nwarps * U) {       // filter row's 2 U + 3 inputs are read ONCE (broadcast) for 5 U FMAs
      float dv[U][Q];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < Q; ++q)
          dv[u][q] = (ow0 + u < Wo) ? __ldg(dz + ((size_t)row * Wo + ow0 + u) * Cout + lane + 32 * q) : 0.f;
#pragma unroll
      for (int kh = 0; kh < K; ++kh) {
        float xr[2 * U + 4];                                      // 2 U + K - 2 = 19 inputs: five 16-byte broadcast reads
        const float4* xrow = reinterpret_cast<const float4*>(xs + kh * XW + 2 * ow0);
#pragma unroll
        for (int i = 0; i < (2 * U + 4) / 4; ++i) { const float4 t = xrow[i]; xr[4 * i] = t.x; xr[4 * i + 1] = t.y; xr[4 * i + 2] = t.z; xr[4 * i + 3] = t.w; }
#pragma unroll
        for (int kw = 0; kw < K; ++kw)
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int q = 0; q < Q; ++q) acc[kh * K + kw][q] = fmaf(xr[2 * u + kw], dv[u][q], acc[kh * K + kw][q]);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < K * K; ++t)
#pragma unroll
    for (int q = 0; q < Q; ++q) red[(size_t)warp * K * K * Cout + t * Cout + lane + 32 * q] = acc[t][q];
  __syncthreads();
  for (int i = threadIdx.x; i < K * K * Cout; i += blockDim.x) {
    float s = 0.f;
    for (int wv = 0; wv < nwarps; ++wv) s += red[(size_t)wv * K * K * Cout + i];
    partial[(size_t)blockIdx.x * K * K * Cout + i] = s;
  }
}

int uad_conv_c1_wgrad_blocks(int B, int H) {
  int rows = B * (H / 2);
  return rows < 4 * UAD_NUM_SMS ? rows : 4 * UAD_NUM_SMS;
}

int uad_launch_conv_c1_wgrad(const float* x, const float* dz, float* dw, int B, int H, int W, int Cout, int ksize,
                             int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  UAD_REQUIRE(ksize == 5, "conv_c1_wgrad: only k=5 (got %d)", ksize);
  UAD_REQUIRE(Cout == 32 || Cout == 64, "conv_c1_wgrad: Cout=%d must be 32 or 64", Cout);
  const int pad_lo = (ksize - 2) / 2;
  const int blocks = uad_conv_c1_wgrad_blocks(B, H);
  const int nthreads = 256;
  size_t n = (size_t)ksize * ksize * Cout;
  size_t need = (size_t)blocks * n * sizeof(float);
  UAD_REQUIRE(ws && ws_bytes >= need, "conv_c1_wgrad: workspace too small (%zu < %zu)", ws_bytes, need);
  UAD_REQUIRE(W % 4 == 0, "conv_c1_wgrad: W=%d must be a multiple of 4", W);
  size_t smem = ((size_t)ksize * (W + 24) + (size_t)(nthreads / 32) * n) * sizeof(float);
  float* partial = reinterpret_cast<float*>(ws);
  static bool attr_set = false;   // set once, outside any stream capture in practice (first eager warm-up step)
  if (!attr_set) {
    UAD_CUDA(cudaFuncSetAttribute(conv_c1_wgrad_kernel<5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    UAD_CUDA(cudaFuncSetAttribute(conv_c1_wgrad_kernel<5, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  if (Cout == 32) {
    conv_c1_wgrad_kernel<5, 1><<<blocks, nthreads, smem, st>>>(x, dz, partial, B, H, W, Cout, pad_lo);
  } else {
    conv_c1_wgrad_kernel<5, 2><<<blocks, nthreads, smem, st>>>(x, dz, partial, B, H, W, Cout, pad_lo);
  }
  UAD_LAUNCH_CHECK("conv_c1_wgrad");
  return uad_launch_splitk_reduce(partial, blocks, n, dw, accumulate, st);
}

// dx[b,ih,iw] = sum_{t,co} dz[b,oh,ow,co] * w[t][co]  with ih = 2*oh + kh - pad_lo  (conv dgrad for Cin == 1; ceVAE anomaly)
template <int K>
__global__ void __launch_bounds__(256) conv_c1_dgrad_kernel(const float* __restrict__ dz, const float* __restrict__ w,
                                                            float* __restrict__ dx, int B, int H, int W, int Cout, int pad_lo) {
  extern __shared__ __align__(16) float ws[];   // [K*K][Cout]
  for (int i = threadIdx.x; i < K * K * Cout; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int Ho = H / 2, Wo = W / 2;
  const int lane = threadIdx.x & 31;
  const size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const size_t npix = (size_t)B * H * W;
  for (size_t pix = warp_global; pix < npix; pix += nwarps) {
    int iw = pix % W, ih = (pix / W) % H, b = pix / ((size_t)W * H);
    float s = 0.f;
    for (int kh = 0; kh < K; ++kh) {
      int th = ih + pad_lo - kh;
      if (th < 0 || (th & 1)) continue;
      int oh = th >> 1;
      if (oh >= Ho) continue;
      for (int kw = 0; kw < K; ++kw) {
        int tw = iw + pad_lo - kw;
        if (tw < 0 || (tw & 1)) continue;
        int ow = tw >> 1;
        if (ow >= Wo) continue;
        const float* d = dz + (((size_t)b * Ho + oh) * Wo + ow) * Cout;
        const float* wr = ws + (kh * K + kw) * Cout;
        for (int c = lane; c < Cout; c += 32) s = fmaf(d[c], wr[c], s);
      }
    }
    s = uad_warp_sum(s);
    if (lane == 0) dx[pix] = s;
  }
}

int uad_launch_conv_c1_dgrad(const float* dz, const float* w, float* dx, int B, int H, int W, int Cout, int ksize,
                             cudaStream_t st) {
  UAD_REQUIRE(ksize == 5, "conv_c1_dgrad: only k=5 (got %d)", ksize);
  const int pad_lo = (ksize - 2) / 2;
  size_t smem = (size_t)ksize * ksize * Cout * sizeof(float);
  UAD_REQUIRE(smem <= 48 * 1024, "conv_c1_dgrad: Cout=%d too large", Cout);
  conv_c1_dgrad_kernel<5><<<UAD_NUM_SMS * 8, 256, smem, st>>>(dz, w, dx, B, H, W, Cout, pad_lo);
  UAD_LAUNCH_CHECK("conv_c1_dgrad");
  return 0;
}
