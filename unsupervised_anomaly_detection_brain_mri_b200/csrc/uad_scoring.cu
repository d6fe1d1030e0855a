// Residual-map anomaly scoring kernels (HBM-bound; bit-exact against the numpy reference semantics).
//   utils/Evaluation.py:282-291  residual, brain mask, hyper-intensity prior
//   utils/Evaluation.py:453-457  diffs > t        (float64 compare)
//   trainers/Metrics.py:67-72    Dice counts      (integer sums)
#include "uad_common.cuh"

__global__ void residual_score_kernel(const float* __restrict__ x, const float* __restrict__ xhat,
                                      const uint8_t* __restrict__ mask, double prior, int keep_positive, int apply_prior,
                                      float* __restrict__ diff, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n4 = n / 4;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + q);
    const float4 hv = __ldg(reinterpret_cast<const float4*>(xhat) + q);
    uchar4 mv = make_uchar4(1, 1, 1, 1);
    if (mask) mv = __ldg(reinterpret_cast<const uchar4*>(mask) + q);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, hs[4] = {hv.x, hv.y, hv.z, hv.w};
    const unsigned char ms[4] = {mv.x, mv.y, mv.z, mv.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float r = xs[j] - hs[j];                       // fp32 subtract, as numpy float32 - float32
      float d = keep_positive ? fmaxf(r, 0.f) : fabsf(r);
      d = ms[j] ? d : 0.f * d;                             // np.multiply(bool, f32): 0*d (keeps -0/NaN semantics)
      if (apply_prior && (double)xs[j] < prior) d = 0.f;
      o[j] = d;
    }
    reinterpret_cast<float4*>(diff)[q] = make_float4(o[0], o[1], o[2], o[3]);
  }
  // tail
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float r = x[i] - xhat[i];
    float d = keep_positive ? fmaxf(r, 0.f) : fabsf(r);
    if (mask) d = mask[i] ? d : 0.f * d;
    if (apply_prior && (double)x[i] < prior) d = 0.f;
    diff[i] = d;
  }
}

extern "C" int uad_residual_score(const float* x, const float* xhat, const uint8_t* mask, double prior_quantile,
                                  int keep_positive, int apply_prior, float* diff, size_t n, void* stream) {
  if (n == 0) return 0;
  UAD_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)xhat % 16 == 0) && ((uintptr_t)diff % 16 == 0) &&
              (!mask || (uintptr_t)mask % 4 == 0), "uad_residual_score: unaligned buffers");
  long long blocks = (long long)((n / 4 + 255) / 256);
  if (blocks > 8 * UAD_NUM_SMS) blocks = 8 * UAD_NUM_SMS;
  if (blocks < 1) blocks = 1;
  residual_score_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, xhat, mask, prior_quantile, keep_positive,
                                                                       apply_prior, diff, n);
  UAD_LAUNCH_CHECK("residual_score");
  return 0;
}

#define UAD_MAX_THR 32
struct ThrList { int n; double t[UAD_MAX_THR]; };

__global__ void __launch_bounds__(256) threshold_counts_kernel(const float* __restrict__ diff, const uint8_t* __restrict__ label,
                                                               size_t n, const __grid_constant__ ThrList thr,
                                                               unsigned long long* __restrict__ counts,
                                                               uint8_t* __restrict__ mask_out) {
  __shared__ unsigned int sh[2 * UAD_MAX_THR + 1];
  for (int i = threadIdx.x; i < 2 * UAD_MAX_THR + 1; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  unsigned int cp[UAD_MAX_THR], cpg[UAD_MAX_THR], cg = 0u;
#pragma unroll
  for (int k = 0; k < UAD_MAX_THR; ++k) { cp[k] = 0u; cpg[k] = 0u; }
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double d = (double)diff[i];
    const unsigned int g = label ? (label[i] != 0) : 0u;
    cg += g;
#pragma unroll
    for (int k = 0; k < UAD_MAX_THR; ++k) {
      if (k < thr.n) {
        const unsigned int pbit = d > thr.t[k] ? 1u : 0u;
        cp[k] += pbit;
        cpg[k] += pbit & g;
        if (k == 0 && mask_out) mask_out[i] = (uint8_t)pbit;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < UAD_MAX_THR; ++k) {
    if (k < thr.n) {
      unsigned int a = cp[k], b = cpg[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
      if ((threadIdx.x & 31) == 0) { atomicAdd(&sh[2 * k], b); atomicAdd(&sh[2 * k + 1], a); }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cg += __shfl_xor_sync(0xffffffffu, cg, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&sh[2 * UAD_MAX_THR], cg);
  __syncthreads();
  // integer atomics: order-independent, hence deterministic
  for (int k = threadIdx.x; k < thr.n; k += blockDim.x) {
    atomicAdd(&counts[3 * k + 0], (unsigned long long)sh[2 * k]);
    atomicAdd(&counts[3 * k + 1], (unsigned long long)sh[2 * k + 1]);
    atomicAdd(&counts[3 * k + 2], (unsigned long long)sh[2 * UAD_MAX_THR]);
  }
}

extern "C" int uad_threshold_counts(const float* diff, const uint8_t* label, size_t n, const double* thresholds_host,
                                    int n_thr, int64_t* counts_dev, uint8_t* mask_out, void* stream) {
  UAD_REQUIRE(n_thr >= 1 && n_thr <= UAD_MAX_THR, "uad_threshold_counts: n_thr=%d must be in [1,%d]", n_thr, UAD_MAX_THR);
  UAD_REQUIRE(n < ((size_t)1 << 40), "uad_threshold_counts: n too large");
  cudaStream_t st = (cudaStream_t)stream;
  ThrList thr;
  thr.n = n_thr;
  for (int k = 0; k < UAD_MAX_THR; ++k) thr.t[k] = k < n_thr ? thresholds_host[k] : 0.0;
  UAD_CUDA(cudaMemsetAsync(counts_dev, 0, (size_t)3 * n_thr * sizeof(int64_t), st));
  if (n == 0) return 0;
  // per-thread 32-bit counters: keep each thread below 2^31 elements (always true for grid >= 1 and n < 2^40 / ...)
  long long blocks = (long long)((n + 256 * 16 - 1) / (256 * 16));
  if (blocks > 8 * UAD_NUM_SMS) blocks = 8 * UAD_NUM_SMS;
  if (blocks < 1) blocks = 1;
  UAD_REQUIRE(n / ((size_t)blocks * 256) < ((size_t)1 << 31), "uad_threshold_counts: n too large for 32-bit lane counters");
  threshold_counts_kernel<<<(int)blocks, 256, 0, st>>>(diff, label, n, thr, (unsigned long long*)counts_dev, mask_out);
  UAD_LAUNCH_CHECK("threshold_counts");
  return 0;
}

// ------------------------------------------------------------------------------------------------ post-processing stencils
// Brain-mask erosion (utils/Evaluation.py:84-89): scipy.ndimage.binary_erosion(mask, generate_binary_structure(2,1),
// iterations=12) per slice, border_value = 0.  One CTA erodes a 32x32 output tile in shared memory: the tile plus a halo of
// `iterations` pixels is loaded once, the cross-shaped erosion is applied `iterations` times ping-pong, the centre is stored.
#define UAD_ER_TILE 32
#define UAD_ER_MAXIT 24
__global__ void __launch_bounds__(256) erode_cross_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W,
                                                          int iters) {
  extern __shared__ uint8_t er_sm[];
  const int D = UAD_ER_TILE + 2 * iters;
  uint8_t* a = er_sm;
  uint8_t* b = er_sm + D * D;
  const size_t base = (size_t)blockIdx.z * H * W;
  const int x0 = blockIdx.x * UAD_ER_TILE - iters, y0 = blockIdx.y * UAD_ER_TILE - iters;
  for (int i = threadIdx.x; i < D * D; i += 256) {
    const int gy = y0 + i / D, gx = x0 + i % D;
    a[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? (in[base + (size_t)gy * W + gx] != 0) : 0;   // outside the image: 0
  }
  __syncthreads();
  for (int t = 0; t < iters; ++t) {
    for (int i = threadIdx.x; i < D * D; i += 256) {
      const int y = i / D, x = i % D;
      uint8_t v = 0;
      if (y > 0 && y < D - 1 && x > 0 && x < D - 1) v = a[i] & a[i - 1] & a[i + 1] & a[i - D] & a[i + D];
      b[i] = v;                 // cells within t+1 of the smem border are not meaningful; the centre stays exact
    }
    __syncthreads();
    uint8_t* tmp = a; a = b; b = tmp;
  }
  for (int i = threadIdx.x; i < UAD_ER_TILE * UAD_ER_TILE; i += 256) {
    const int ty = i / UAD_ER_TILE, tx = i % UAD_ER_TILE;
    const int gy = blockIdx.y * UAD_ER_TILE + ty, gx = blockIdx.x * UAD_ER_TILE + tx;
    if (gy < H && gx < W) out[base + (size_t)gy * W + gx] = a[(ty + iters) * D + tx + iters];
  }
}

extern "C" int uad_binary_erosion_cross(const uint8_t* mask, uint8_t* out, int N, int H, int W, int iterations, void* stream) {
  UAD_REQUIRE(iterations >= 1 && iterations <= UAD_ER_MAXIT, "uad_binary_erosion_cross: iterations must be in [1,%d]", UAD_ER_MAXIT);
  UAD_REQUIRE(N > 0 && H > 0 && W > 0 && N <= 65535, "uad_binary_erosion_cross: bad dims");
  const int D = UAD_ER_TILE + 2 * iterations;
  erode_cross_kernel<<<dim3(uad_cdiv(W, UAD_ER_TILE), uad_cdiv(H, UAD_ER_TILE), N), 256, 2 * D * D, (cudaStream_t)stream>>>(
      mask, out, H, W, iterations);
  UAD_LAUNCH_CHECK("erode_cross");
  return 0;
}

// 5x5x5 median (utils/Evaluation.py:108-110, applied at :311-312): scipy.ndimage.median_filter(volume, (5,5,5)), boundary
// mode 'reflect' (d c b a | a b c d | d c b a), rank 62 of the 125 sorted neighbours.  Each thread selects its voxel's median
// by a most-significant-bit-first radix search over order-preserving integer keys of the 125 shared-memory neighbours - exact,
// no sorting network, early exit when the neighbourhood is constant (the masked background).
__device__ __forceinline__ uint32_t med_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float med_unkey(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ int med_reflect(int i, int n) {
  while (i < 0 || i >= n) i = (i < 0) ? (-i - 1) : (2 * n - 1 - i);
  return i;
}
#define UAD_MED_TX 32
#define UAD_MED_TY 8
__global__ void __launch_bounds__(UAD_MED_TX * UAD_MED_TY) median3d5_kernel(const float* __restrict__ vol, float* __restrict__ out,
                                                                           int Z, int H, int W) {
  constexpr int SX = UAD_MED_TX + 4, SY = UAD_MED_TY + 4;
  __shared__ uint32_t sm[5][SY][SX];
  const int z = blockIdx.z, y0 = blockIdx.y * UAD_MED_TY, x0 = blockIdx.x * UAD_MED_TX;
  const int tid = threadIdx.y * UAD_MED_TX + threadIdx.x;
  for (int i = tid; i < 5 * SY * SX; i += UAD_MED_TX * UAD_MED_TY) {
    const int dz = i / (SY * SX), r = i % (SY * SX), sy = r / SX, sx = r % SX;
    const int gz = med_reflect(z + dz - 2, Z), gy = med_reflect(y0 + sy - 2, H), gx = med_reflect(x0 + sx - 2, W);
    sm[dz][sy][sx] = med_key(vol[((size_t)gz * H + gy) * W + gx]);
  }
  __syncthreads();
  const int y = y0 + threadIdx.y, x = x0 + threadIdx.x;
  if (y >= H || x >= W) return;
  uint32_t lo = 0xffffffffu, hi = 0u;
  for (int dz = 0; dz < 5; ++dz)
    for (int dy = 0; dy < 5; ++dy)
#pragma unroll
      for (int dx = 0; dx < 5; ++dx) {
        const uint32_t k = sm[dz][threadIdx.y + dy][threadIdx.x + dx];
        lo = min(lo, k);
        hi = max(hi, k);
      }
  uint32_t prefix = lo;
  if (lo != hi) {
    const int top = 31 - __clz(lo ^ hi);                 // highest bit in which the neighbourhood differs
    const uint32_t keep = (top == 31) ? 0u : (0xffffffffu << (top + 1));
    prefix = lo & keep;
    uint32_t known = keep;                               // bits of the answer fixed so far
    int kth = 62;                                        // rank (0-based) within the candidates matching the prefix
    for (int bit = top; bit >= 0; --bit) {
      const uint32_t m = 1u << bit;
      int cnt0 = 0;
      for (int dz = 0; dz < 5; ++dz)
        for (int dy = 0; dy < 5; ++dy)
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) {
            const uint32_t k = sm[dz][threadIdx.y + dy][threadIdx.x + dx];
            cnt0 += ((k & known) == prefix && !(k & m)) ? 1 : 0;
          }
      if (kth >= cnt0) { kth -= cnt0; prefix |= m; }
      known |= m;
    }
  }
  out[((size_t)z * H + y) * W + x] = med_unkey(prefix);
}

extern "C" int uad_median_filter3d_5(const float* vol, float* out, int Z, int H, int W, void* stream) {
  UAD_REQUIRE(Z > 0 && H > 0 && W > 0 && Z <= 65535, "uad_median_filter3d_5: bad dims");
  UAD_REQUIRE(vol != out, "uad_median_filter3d_5: in-place operation is not supported");
  median3d5_kernel<<<dim3(uad_cdiv(W, UAD_MED_TX), uad_cdiv(H, UAD_MED_TY), Z), dim3(UAD_MED_TX, UAD_MED_TY), 0,
                     (cudaStream_t)stream>>>(vol, out, Z, H, W);
  UAD_LAUNCH_CHECK("median3d5");
  return 0;
}
