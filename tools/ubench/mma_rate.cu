// Micro-benchmark (developer aid): sustained cost of tcgen05.mma instructions (cycles per instruction) for the shapes the
// 3xTF32 kernels issue.  One CTA per SM (148), one issuing thread, REP back-to-back MMAs + one commit, clock64 around it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_rate tools/ubench/mma_rate.cu -lcuda && build/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ uint32_t try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

constexpr int REP = 512;
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// mode: 0 tf32 TS, 1 tf32 SS, 2 bf16 TS, 3 bf16 SS;  nd = number of independent accumulators the stream rotates over
template <int mode, int nd>
__global__ void __launch_bounds__(128, 1) k(long long* out, int N) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the operand area (values do not matter for timing; avoid NaN slow paths just in case)
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const bool f16 = mode >= 2;
    // idesc: D fp32 (1<<4); A/B format at [7,10)/[10,13): tf32 = 2, bf16 = 1; N>>3 at [17,23); M>>4 at [24,29)
    const uint32_t fmt = f16 ? 1u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t adesc = make_sw128_desc(base);
    const uint64_t bdesc = make_sw128_desc(base + 32 * 1024);
    const uint32_t a_tmem = tmem + 256;
    long long t0 = clock64();
    for (int i0 = 0; i0 < REP; i0 += 8) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int i = i0 + j;
          const uint32_t d = tmem + (uint32_t)((j % nd) * N);     // nd == 1: dependent accumulation chain
          const uint32_t kk = (j & 3) * 2;                        // walk 4 K-slices of the 128-byte row
          if (mode == 0) mma_tf32_ts(d, a_tmem + (j & 3) * 8, bdesc + kk, idesc, i >= nd);
          else if (mode == 1) mma_tf32_ss(d, adesc + kk, bdesc + kk, idesc, i >= nd);
          else if (mode == 2) mma_f16_ts(d, a_tmem + (j & 3) * 8, bdesc + kk, idesc, i >= nd);
          else mma_f16_ss(d, adesc + kk, bdesc + kk, idesc, i >= nd);
        }
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (elect_one()) tc_commit(smem_u32(&bar));
    __syncwarp();
    while (!try_wait(smem_u32(&bar), 0)) {}
    long long t2 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int mode, int nd>
static void launch1(long long* d, int N) {
  cudaFuncSetAttribute(k<mode, nd>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<mode, nd><<<148, 128, 100 * 1024>>>(d, N);
}
template <int mode>
static void launch0(int nd, long long* d, int N) {
  if (nd == 1) launch1<mode, 1>(d, N); else if (nd == 2) launch1<mode, 2>(d, N); else launch1<mode, 4>(d, N);
}
static void launch(int mode, int nd, long long* d, int N) {
  if (mode == 0) launch0<0>(nd, d, N); else if (mode == 1) launch0<1>(nd, d, N); else if (mode == 2) launch0<2>(nd, d, N); else launch0<3>(nd, d, N);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16 * sizeof(long long));
  const char* names[] = {"tf32 A=TMEM", "tf32 A=smem", "bf16 A=TMEM", "bf16 A=smem"};
  for (int mode = 0; mode < 4; ++mode)
    for (int N : {32, 64, 128})
      for (int nd : {1, 2, 4}) {
        if (nd * N > 256) continue;
        cudaMemset(d, 0, 16 * sizeof(long long));
        launch(mode, nd, d, N);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-12s M=128 N=%3d K=32B  %d accumulator(s): issue %6.1f cyc/mma, complete %6.1f cyc/mma  (%s)\n", names[mode], N, nd,
               (double)h[0] / REP, (double)h[1] / REP, cudaGetErrorString(e));
      }
  return 0;
}
