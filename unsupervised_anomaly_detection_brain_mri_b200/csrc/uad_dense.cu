// Small dense / 1x1-conv GEMMs (bottleneck layers: < 0.1 % of the step's FLOPs) with the same fused epilogue as the convs.
#include <stdlib.h>

#include "uad_common.cuh"

// C[M,N] = sum_k A(m,k) * B(k,n), A(m,k) = A[m*sa_m + k*sa_k] * (Amul ? Amul[...] * mulscale : 1), same for B.
struct SmallGemm {
  const float* A; const float* Amul; long long sa_m, sa_k;
  const float* Bm; const float* Bmul; long long sb_k, sb_n;
  float mulscale;
  int M, N, K;
  // epilogue
  const float* bias; const float* mask; float mask_scale;
  const float* gamma; const float* beta; float bn_c; int act; float alpha;
  float* z_out; float* a_out;   // row-major [M,N]
  int accumulate;               // z_out += (plain accumulation mode: no bias/mask/affine expected)
  float* partial;               // split-K: raw partial sums [splits][M][N] (epilogue runs in small_gemm_epilogue_kernel)
  int kchunk;                   // K range per blockIdx.z
};

__global__ void __launch_bounds__(256) small_gemm_kernel(const __grid_constant__ SmallGemm g) {
  constexpr int T = 64, BK = 16;
  __shared__ float As[BK][T + 1];
  __shared__ float Bs[BK][T + 1];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * T, n0 = blockIdx.y * T;
  const int tx = tid % 16, ty = tid / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int kbeg = blockIdx.z * g.kchunk;
  const int kend = min(g.K, kbeg + g.kchunk);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    for (int i = tid; i < T * BK; i += 256) {
      int kk, mm;
      if (g.sa_k == 1) { kk = i % BK; mm = i / BK; } else { mm = i % T; kk = i / T; }
      int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < g.M && k < kend) {
        long long off = m * g.sa_m + k * g.sa_k;
        v = g.A[off];
        if (g.Amul) v *= g.Amul[off] * g.mulscale;
      }
      As[kk][mm] = v;
    }
    for (int i = tid; i < T * BK; i += 256) {
      int kk, nn;
      if (g.sb_n == 1) { nn = i % T; kk = i / T; } else { kk = i % BK; nn = i / BK; }
      int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < g.N && k < kend) {
        long long off = k * g.sb_k + n * g.sb_n;
        v = g.Bm[off];
        if (g.Bmul) v *= g.Bmul[off] * g.mulscale;
      }
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty + 16 * i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx + 16 * j;
      if (n >= g.N) continue;
      size_t o = (size_t)m * g.N + n;
      float z = acc[i][j];
      if (g.partial) { g.partial[(size_t)blockIdx.z * g.M * g.N + o] = z; continue; }
      if (g.accumulate) { g.z_out[o] += z; continue; }
      if (g.bias) z += g.bias[n];
      if (g.mask) z *= g.mask[o] * g.mask_scale;
      if (g.z_out) g.z_out[o] = z;
      if (g.a_out) {
        float u = z;
        if (g.gamma) u = g.gamma[n] * g.bn_c * z + g.beta[n];
        g.a_out[o] = uad_act(u, g.act, g.alpha);
      }
    }
  }
}

// split-K second stage: deterministic sum of the partials + the fused epilogue.  G = 8 (from 32 splits on): block = 32 outputs x
// 8 split groups (a thread per output walking up to 128 partials serially is latency bound: 16 us), the group sums are added in
// a fixed order; G = 1: one thread per output.
template <int G>
__global__ void __launch_bounds__(256) small_gemm_epilogue_kernel(const __grid_constant__ SmallGemm g, int splits) {
  __shared__ float sh[G][256 / G];
  constexpr int OUTS = 256 / G;
  const int lo = threadIdx.x % OUTS, grp = threadIdx.x / OUTS;
  const size_t o = (size_t)blockIdx.x * OUTS + lo;
  const size_t MN = (size_t)g.M * g.N;
  float part = 0.f;
  if (o < MN)
    for (int s = grp; s < splits; s += G) part += g.partial[(size_t)s * MN + o];
  float z = part;
  if (G > 1) {
    sh[grp][lo] = part;
    __syncthreads();
    if (grp != 0) return;
    z = 0.f;
#pragma unroll
    for (int k = 0; k < G; ++k) z += sh[k][lo];
  }
  if (o >= MN) return;
  const int n = (int)(o % g.N);
  if (g.accumulate) { g.z_out[o] += z; return; }
  if (g.bias) z += g.bias[n];
  if (g.mask) z *= g.mask[o] * g.mask_scale;
  if (g.z_out) g.z_out[o] = z;
  if (g.a_out) {
    float u = z;
    if (g.gamma) u = g.gamma[n] * g.bn_c * z + g.beta[n];
    g.a_out[o] = uad_act(u, g.act, g.alpha);
  }
}

static int plan_splits(int M, int N, int K, int* kchunk) {
  const int tiles = uad_cdiv(M, 64) * uad_cdiv(N, 64);
  int splits = (2 * UAD_NUM_SMS) / tiles;
  const int max_splits = uad_cdiv(K, 32);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  *kchunk = uad_cdiv(uad_cdiv(K, splits), 16) * 16;
  return uad_cdiv(K, *kchunk);
}

static int launch_small(SmallGemm g, cudaStream_t st, void* ws, size_t ws_bytes) {
  int kchunk;
  const int splits = plan_splits(g.M, g.N, g.K, &kchunk);
  if (splits > 1 && ws && ws_bytes >= (size_t)splits * g.M * g.N * sizeof(float)) {
    g.partial = reinterpret_cast<float*>(ws);
    g.kchunk = kchunk;
    dim3 grid(uad_cdiv(g.M, 64), uad_cdiv(g.N, 64), splits);
    small_gemm_kernel<<<grid, 256, 0, st>>>(g);
    UAD_LAUNCH_CHECK("small_gemm");
    if (splits >= 32) small_gemm_epilogue_kernel<8><<<uad_cdiv((size_t)g.M * g.N, 32), 256, 0, st>>>(g, splits);
    else small_gemm_epilogue_kernel<1><<<uad_cdiv((size_t)g.M * g.N, 256), 256, 0, st>>>(g, splits);
    UAD_LAUNCH_CHECK("small_gemm_epilogue");
    return 0;
  }
  g.partial = nullptr;
  g.kchunk = g.K;
  dim3 grid(uad_cdiv(g.M, 64), uad_cdiv(g.N, 64));
  small_gemm_kernel<<<grid, 256, 0, st>>>(g);
  UAD_LAUNCH_CHECK("small_gemm");
  return 0;
}

// workspace regions of uad_dense_bwd: its three parts (dx, dw, dbias) run concurrently, each with its own split-K partials
static size_t ws_align(size_t floats) { return (floats * sizeof(float) + 255) & ~(size_t)255; }
static size_t bwd_dx_bytes(int M, int K, int N) { int kc; return ws_align((size_t)plan_splits(M, K, N, &kc) * M * K); }
static size_t bwd_dw_bytes(int M, int K, int N) { int kc; return ws_align((size_t)plan_splits(K, N, M, &kc) * K * N); }
static size_t bwd_db_bytes(int N) { return ws_align((size_t)64 * N); }

extern "C" size_t uad_dense_workspace_bytes(int M, int K, int N) {
  // forward: y[M,N] over K; backward: dx[M,K] over N, dw[K,N] over M and the column-sum partials side by side
  int kc;
  const size_t fwd = ws_align((size_t)plan_splits(M, N, K, &kc) * M * N);
  const size_t bwd = bwd_dx_bytes(M, K, N) + bwd_dw_bytes(M, K, N) + bwd_db_bytes(N);
  return (fwd > bwd ? fwd : bwd) + 256;
}

// Fork / join of the three independent parts of a dense backward (46 launches of ~5 us each make up the bottleneck chain of a
// step: latency, not work).  Two internal side streams wait for an event recorded on the caller's stream and the caller's stream
// waits for theirs, so the call is still ordered like one operation on `st` - also under CUDA-graph capture, where the side
// streams join the capture as parallel branches.  UAD_DENSE_FORK=0 issues the parts one after the other.
struct DenseFork {
  cudaStream_t side[2];
  cudaEvent_t fork, join[2];
  bool ok;
};
static DenseFork* dense_fork() {
  static DenseFork f;
  static int state = 0;   // 0 = untried, 1 = ready, -1 = disabled
  if (state == 0) {
    const char* e = getenv("UAD_DENSE_FORK");
    state = -1;
    if (!(e && atoi(e) == 0)) {
      bool ok = cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming) == cudaSuccess;
      for (int i = 0; i < 2 && ok; ++i)
        ok = cudaStreamCreateWithFlags(&f.side[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&f.join[i], cudaEventDisableTiming) == cudaSuccess;
      if (ok) state = 1; else cudaGetLastError();
    }
  }
  return state == 1 ? &f : nullptr;
}

// dbias[n] (+)= sum_m dz[m,n] * (mask ? mask*scale : 1): grid (N/32, row-splits) -> partial[split][N] -> final (deterministic)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dz, const float* __restrict__ mask, float scale,
                                                     float* __restrict__ partial, int M, int N, int rows_per_split) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const int m0 = blockIdx.y * rows_per_split;
  const int m1 = min(M, m0 + rows_per_split);
  float s = 0.f;
  if (n < N)
    for (int m = m0 + warp; m < m1; m += 8) {
      float v = dz[(size_t)m * N + n];
      if (mask) v *= mask[(size_t)m * N + n] * scale;
      s += v;
    }
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][lane];
    partial[(size_t)blockIdx.y * N + n] = t;
  }
}

__global__ void colsum_final_kernel(const float* __restrict__ partial, int splits, int N, float* __restrict__ out, int accumulate) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float t = 0.f;
  for (int s = 0; s < splits; ++s) t += partial[(size_t)s * N + n];
  out[n] = accumulate ? out[n] + t : t;
}

extern "C" int uad_dense_fwd(const float* x, const float* w, const float* bias, const float* mask, float mask_scale,
                             const float* gamma, const float* beta, float* z_out, float* a_out, int M, int K, int N,
                             int act, float alpha, float bn_c, void* ws, size_t ws_bytes, void* stream) {
  UAD_REQUIRE(M > 0 && K > 0 && N > 0, "uad_dense_fwd: bad dims");
  UAD_REQUIRE((gamma == nullptr) == (beta == nullptr), "uad_dense_fwd: gamma/beta must both be set or both NULL");
  SmallGemm g = {};
  g.A = x; g.sa_m = K; g.sa_k = 1;
  g.Bm = w; g.sb_k = N; g.sb_n = 1;
  g.mulscale = 1.f;
  g.M = M; g.N = N; g.K = K;
  g.bias = bias; g.mask = mask; g.mask_scale = mask_scale;
  g.gamma = gamma; g.beta = beta; g.bn_c = bn_c; g.act = act; g.alpha = alpha;
  g.z_out = z_out; g.a_out = a_out;
  return launch_small(g, (cudaStream_t)stream, ws, ws_bytes);
}

extern "C" int uad_dense_bwd(const float* x, const float* w, const float* dz, const float* mask, float mask_scale,
                             float* dx, float* dw, float* dbias, int M, int K, int N, int accumulate, void* ws, size_t ws_bytes,
                             void* stream) {
  UAD_REQUIRE(M > 0 && K > 0 && N > 0, "uad_dense_bwd: bad dims");
  cudaStream_t st = (cudaStream_t)stream;
  // workspace regions (each part falls back to an unsplit GEMM if its region is missing)
  const size_t dx_b = bwd_dx_bytes(M, K, N), dw_b = bwd_dw_bytes(M, K, N), db_b = bwd_db_bytes(N);
  const bool regions = ws && ws_bytes >= dx_b + dw_b + db_b;
  char* wsc = (char*)ws;
  void* ws_dx = ws; size_t wsb_dx = regions ? dx_b : ws_bytes;
  void* ws_dw = regions ? wsc + dx_b : ws; size_t wsb_dw = regions ? dw_b : ws_bytes;
  void* ws_db = regions ? wsc + dx_b + dw_b : ws; size_t wsb_db = regions ? db_b : ws_bytes;
  const int nparts = (dx ? 1 : 0) + (dw ? 1 : 0) + (dbias ? 1 : 0);
  DenseFork* fk = (regions && nparts > 1) ? dense_fork() : nullptr;
  cudaStream_t st_dw = st, st_db = st;
  if (fk) {
    UAD_CUDA(cudaEventRecord(fk->fork, st));
    if (dx && dw) { UAD_CUDA(cudaStreamWaitEvent(fk->side[0], fk->fork, 0)); st_dw = fk->side[0]; }
    if ((dx || dw) && dbias) { UAD_CUDA(cudaStreamWaitEvent(fk->side[1], fk->fork, 0)); st_db = fk->side[1]; }
  }
  if (dx) {   // dx[M,K] = (dz*mask)[M,N] . W^T[N,K]
    SmallGemm g = {};
    g.A = dz; g.Amul = mask; g.sa_m = N; g.sa_k = 1;
    g.Bm = w; g.sb_k = 1; g.sb_n = N;           // B(k'=n, n'=k) = W[k*N + n]
    g.mulscale = mask_scale;
    g.M = M; g.N = K; g.K = N;
    g.z_out = dx;
    if (int e = launch_small(g, st, ws_dx, wsb_dx)) return e;
  }
  if (dw) {   // dw[K,N] (+)= x^T[K,M] . (dz*mask)[M,N]
    SmallGemm g = {};
    g.A = x; g.sa_m = 1; g.sa_k = K;            // A(m'=k, k'=m) = x[m*K + k]
    g.Bm = dz; g.Bmul = mask; g.sb_k = N; g.sb_n = 1;
    g.mulscale = mask_scale;
    g.M = K; g.N = N; g.K = M;
    g.z_out = dw; g.accumulate = accumulate;
    if (int e = launch_small(g, st_dw, ws_dw, wsb_dw)) return e;
  }
  if (dbias) {
    int splits = uad_cdiv(M, 64);
    if (splits > 64) splits = 64;
    const int rps = uad_cdiv(M, splits);
    splits = uad_cdiv(M, rps);
    UAD_REQUIRE(ws_db && wsb_db >= (size_t)splits * N * sizeof(float), "uad_dense_bwd: workspace too small");
    colsum_kernel<<<dim3(uad_cdiv(N, 32), splits), 256, 0, st_db>>>(dz, mask, mask_scale, (float*)ws_db, M, N, rps);
    UAD_LAUNCH_CHECK("colsum");
    colsum_final_kernel<<<uad_cdiv(N, 128), 128, 0, st_db>>>((const float*)ws_db, splits, N, dbias, accumulate);
    UAD_LAUNCH_CHECK("colsum_final");
  }
  if (fk) {
    if (st_dw != st) { UAD_CUDA(cudaEventRecord(fk->join[0], st_dw)); UAD_CUDA(cudaStreamWaitEvent(st, fk->join[0], 0)); }
    if (st_db != st) { UAD_CUDA(cudaEventRecord(fk->join[1], st_db)); UAD_CUDA(cudaStreamWaitEvent(st, fk->join[1], 0)); }
  }
  return 0;
}
