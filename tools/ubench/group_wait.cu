// Micro-benchmark (developer aid): cost of ONE wait on an already completed mbarrier phase for a 128-thread group (the
// converter warpgroup of the tcgen05 kernels), per polling strategy.  Cycles per wait, dependent chain of REP waits.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint32_t try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { while (!try_wait(bar, parity)) {} }
constexpr int REP = 256;
// strategy 0: every thread polls; 1: lane 0 of each warp polls, __syncwarp; 2: thread 0 polls, bar.sync 1, nthreads;
// 3: every thread polls, 8 distinct barriers round-robin (same as 0 but tests address reuse)
__global__ void k(long long* out, int strategy, int nthreads) {
  __shared__ __align__(8) unsigned long long bars[8];
  __shared__ long long t_end[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) mbar_arrive(smem_u32(&bars[i]));
  __syncthreads();
  if ((int)threadIdx.x >= nthreads) return;
  const int lane = threadIdx.x & 31;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < REP; ++i) {
    const uint32_t bar = smem_u32(&bars[(i + acc) & 7]);
    if (strategy == 0 || strategy == 3) { mbar_wait(bar, 0); }
    else if (strategy == 1) { if (lane == 0) mbar_wait(bar, 0); __syncwarp(); }
    else { if (threadIdx.x == 0) mbar_wait(bar, 0); asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }
    acc += (uint32_t)(clock64() & 0);        // keep the chain dependent without changing the address
  }
  long long t1 = clock64();
  if (lane == 0) t_end[threadIdx.x >> 5] = t1 - t0;
  __syncwarp();
  if (threadIdx.x == 0) out[0] = (t1 - t0);
}
int main() {
  long long* d;
  cudaMalloc(&d, 8 * sizeof(long long));
  const char* names[] = {"all threads poll", "lane 0 polls + __syncwarp", "thread 0 polls + bar.sync"};
  for (int nthreads : {32, 128, 256})
    for (int strat = 0; strat < 3; ++strat) {
      cudaMemset(d, 0, 8 * sizeof(long long));
      k<<<148, 256>>>(d, strat, nthreads);
      cudaError_t e = cudaDeviceSynchronize();
      long long h;
      cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%3d threads, %-28s: %6.1f cycles per wait (%s)\n", nthreads, names[strat], (double)h / REP, cudaGetErrorString(e));
    }
  return 0;
}
