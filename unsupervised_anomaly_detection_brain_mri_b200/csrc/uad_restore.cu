// Iterative MAP restoration (reference trainers/VAE_You.py:53-54,125-147): the per-iteration elementwise kernels.
//   grads = d/dx [ sum_hw |xhat - x| + kl + lambda * TV(x - xhat) ]          (tf.gradients of a per-sample vector = sum)
// The network part (d/dx through encoder/decoder) is the ordinary dgrad chain seeded with
//   g = dL/dxhat = sign(xhat - x) - lambda * T,   T = dTV(d)/dd at d = x - xhat,
// and x's direct dependence contributes -g, so one iteration is   x <- x - lr * (gx_network - g).
// Both kernels are single HBM passes (8-12 B/pixel); single-channel images (all reference datasets).
#include "uad_common.cuh"

__device__ __forceinline__ float uad_sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// tf.image.total_variation(d) = sum |d[i+1,j]-d[i,j]| + sum |d[i,j+1]-d[i,j]|  (per image, no wrap-around).
// One block row-tiles the image: 32 x 8 pixels per block with a 1-pixel halo kept in shared memory.
__global__ void __launch_bounds__(256) tv_restore_seed_kernel(const float* __restrict__ x, const float* __restrict__ xhat,
                                                              float lambda, float* __restrict__ g, float* __restrict__ tv_partial,
                                                              int H, int W) {
  __shared__ float d[10][34];
  __shared__ float red[8];
  const int b = blockIdx.z;
  const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 8;
  const size_t base = (size_t)b * H * W;
  for (int t = threadIdx.x; t < 10 * 34; t += 256) {
    const int li = t / 34, lj = t % 34;
    const int i = i0 + li - 1, j = j0 + lj - 1;
    float v = 0.f;
    if (i >= 0 && i < H && j >= 0 && j < W) v = x[base + (size_t)i * W + j] - xhat[base + (size_t)i * W + j];
    d[li][lj] = v;
  }
  __syncthreads();
  const int lj = (threadIdx.x & 31) + 1, li = (threadIdx.x >> 5) + 1;
  const int i = i0 + li - 1, j = j0 + lj - 1;
  float tv = 0.f;
  if (i < H && j < W) {
    const float c = d[li][lj];
    float T = 0.f;
    if (i > 0) T += uad_sgn(c - d[li - 1][lj]);
    if (i < H - 1) { const float e = d[li + 1][lj] - c; T -= uad_sgn(e); tv += fabsf(e); }
    if (j > 0) T += uad_sgn(c - d[li][lj - 1]);
    if (j < W - 1) { const float e = d[li][lj + 1] - c; T -= uad_sgn(e); tv += fabsf(e); }
    g[base + (size_t)i * W + j] = uad_sgn(-c) - lambda * T;            // sign(xhat - x) = sign(-d)
  }
  if (tv_partial) {
    tv = uad_warp_sum(tv);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tv;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w];
      tv_partial[((size_t)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = t;
    }
  }
}

__global__ void tv_final_kernel(const float* __restrict__ partial, int per_sample, float* __restrict__ tv, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s = 0.f;
  for (int k = 0; k < per_sample; ++k) s += partial[(size_t)b * per_sample + k];
  tv[b] = s;
}

extern "C" size_t uad_tv_restore_workspace_bytes(int B, int H, int W) {
  return (size_t)B * uad_cdiv(H, 8) * uad_cdiv(W, 32) * sizeof(float) + 256;
}

extern "C" int uad_tv_restore_seed(const float* x, const float* xhat, float tv_lambda, float* g, float* tv, int B, int H, int W,
                                   void* ws, size_t ws_bytes, void* stream) {
  UAD_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535, "uad_tv_restore_seed: bad dims B=%d H=%d W=%d", B, H, W);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(uad_cdiv(W, 32), uad_cdiv(H, 8), B);
  float* partial = nullptr;
  if (tv) {
    UAD_REQUIRE(ws && ws_bytes >= (size_t)B * grid.x * grid.y * sizeof(float), "uad_tv_restore_seed: workspace too small");
    partial = (float*)ws;
  }
  tv_restore_seed_kernel<<<grid, 256, 0, st>>>(x, xhat, tv_lambda, g, partial, H, W);
  UAD_LAUNCH_CHECK("tv_restore_seed");
  if (tv) {
    tv_final_kernel<<<uad_cdiv(B, 128), 128, 0, st>>>(partial, (int)(grid.x * grid.y), tv, B);
    UAD_LAUNCH_CHECK("tv_final");
  }
  return 0;
}

// x <- x - lr * (gx - g): gx = gradient through the network, -g = direct dependence of the L1 / TV terms on x.
// grads_out (nullable) receives the full gradient (what the reference fetches as losses['grads']).
__global__ void restore_update_kernel(float* __restrict__ x, const float* __restrict__ gx, const float* __restrict__ g, float lr,
                                      float* __restrict__ grads_out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gr = gx[i] - g[i];
  if (grads_out) grads_out[i] = gr;
  x[i] -= lr * gr;
}

extern "C" int uad_restore_update(float* x, const float* gx, const float* g, float lr, float* grads_out, size_t n, void* stream) {
  restore_update_kernel<<<uad_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, gx, g, lr, grads_out, n);
  UAD_LAUNCH_CHECK("restore_update");
  return 0;
}
