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
