// Shared helpers for libuad_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/uad_b200.h"

// thread-local error text behind uad_last_error()
int uad_set_error(const char* fmt, ...);

#define UAD_REQUIRE(cond, ...) \
  do { if (!(cond)) return uad_set_error(__VA_ARGS__); } while (0)

extern long long g_uad_launches;
#define UAD_LAUNCH_CHECK(what) \
  do { ++g_uad_launches; cudaError_t e__ = cudaGetLastError(); \
       if (e__ != cudaSuccess) return uad_set_error("%s: launch failed: %s", what, cudaGetErrorString(e__)); } while (0)

#define UAD_CUDA(call) \
  do { cudaError_t e__ = (call); \
       if (e__ != cudaSuccess) return uad_set_error("%s failed: %s", #call, cudaGetErrorString(e__)); } while (0)

static inline int uad_ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
static inline bool uad_is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
static inline int uad_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

#define UAD_NUM_SMS 148

// y = act((x - mean) * rstd * gamma[hw] + beta[hw])  (uad_elementwise.cu; shared with the training forward in uad_fanogan.cu)
int uad_layernorm_hw_apply(const float* x, const float* mean, const float* rstd, const float* gamma_hw, const float* beta_hw,
                           float* y, int B, int HW, int C, int act, float alpha, cudaStream_t st);

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ float uad_act(float u, int act, float alpha) {
  switch (act) {
    case UAD_ACT_LEAKY:   return u > 0.f ? u : alpha * u;
    case UAD_ACT_RELU:    return u > 0.f ? u : 0.f;
    case UAD_ACT_SIGMOID: return 1.f / (1.f + expf(-u));
    case UAD_ACT_TANH:    return tanhf(u);
    default:              return u;
  }
}

// d act(u) / du given u (pre-activation)
__device__ __forceinline__ float uad_act_grad(float u, int act, float alpha) {
  switch (act) {
    case UAD_ACT_LEAKY:   return u > 0.f ? 1.f : alpha;
    case UAD_ACT_RELU:    return u > 0.f ? 1.f : 0.f;
    case UAD_ACT_SIGMOID: { float s = 1.f / (1.f + expf(-u)); return s * (1.f - s); }
    case UAD_ACT_TANH:    { float t = tanhf(u); return 1.f - t * t; }
    default:              return 1.f;
  }
}

// pre-activation u recovered from the OUTPUT a = act(u) of a piecewise-linear activation (UAD_ACT_FROM_OUTPUT mode of the
// backward kernels).  LeakyReLU: exact up to one rounding, sign(u) == sign(a).  ReLU: u is unknown where a == 0, but the
// activation gradient there is 0 and every use of u is multiplied by it.
__device__ __forceinline__ float uad_preact_from_output(float a, int act, float inv_alpha) {
  return (act == UAD_ACT_LEAKY && !(a > 0.f)) ? a * inv_alpha : a;
}

__device__ __forceinline__ float uad_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double uad_warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ long long uad_warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
