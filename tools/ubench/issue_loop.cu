// Micro-benchmark (developer aid): cycles per k-block of the MMA-issuing warp's loop of the 3xTF32 kernels
// (8 paired tcgen05.mma of N=64, two tcgen05.commit, two mbarrier waits) in isolation and next to a warpgroup that keeps
// writing TMEM A slots (tcgen05.st) - which part of the loop serialises with what.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/issue_loop tools/ubench/issue_loop.cu -lcuda
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]),
      "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

constexpr int KB = 256;     // k-blocks per run
// variant bits: 1 = two tcgen05.commit per k-block, 2 = two mbarrier waits (completed barriers) per k-block,
//               4 = a warpgroup streams tcgen05.st into two other A slots meanwhile, 8 = one commit per k-block instead of two
//               16 = the two waits are replaced by early probes (issued before the MMAs, consumed after)
//               32 = no tcgen05.fence::after_thread_sync per k-block, 64 = two k-blocks per loop iteration (one fence, one elect
//               block of 16 MMAs, waits/commits for both), 128 = waits polled by lane 0 only (no __syncwarp after)
__global__ void __launch_bounds__(256, 1) k(long long* out, int variant, int N) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) unsigned long long bars[8];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stop = 0;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  // bars[0], bars[1]: "completed" barriers the issuer waits on (phase 0 completed once); bars[2..3]: commit targets
  if (threadIdx.x == 0) { mbar_arrive(smem_u32(&bars[0])); mbar_arrive(smem_u32(&bars[1])); }
  __syncthreads();
  if (warp == 1) {
    const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idescN = idesc_base | ((uint32_t)(N >> 3) << 17);
    const uint32_t idesc2N = idesc_base | ((uint32_t)((2 * N) >> 3) << 17);
    const uint64_t bdesc = make_sw128_desc(base);
    const uint32_t a_hi = tmem + 256, a_lo = a_hi + 32;
    uint32_t ok0 = 0, ok1 = 0;
    long long t0 = clock64();
    if (variant & 64) {
      for (int i = 0; i < KB; i += 2) {
        if (variant & 2) { mbar_wait(smem_u32(&bars[0]), 0); mbar_wait(smem_u32(&bars[1]), 0); mbar_wait(smem_u32(&bars[0]), 0); mbar_wait(smem_u32(&bars[1]), 0); }
        if (!(variant & 32)) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t d_pair = tmem + h * 2 * N;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              mma_tf32_ts(d_pair, a_hi + j * 8, bdesc + 2 * j, idesc2N, (i >= 2) | (j != 0));
              mma_tf32_ts(d_pair + N, a_lo + j * 8, bdesc + 2 * j, idescN, 1u);
            }
            if (variant & 1) { tc_commit(smem_u32(&bars[2])); tc_commit(smem_u32(&bars[3])); }
          }
        }
        __syncwarp();
      }
    } else
    for (int i = 0; i < KB; ++i) {
      if (variant & 128) { if (lane == 0) { mbar_wait(smem_u32(&bars[0]), 0); mbar_wait(smem_u32(&bars[1]), 0); } }
      if (variant & 2) { mbar_wait(smem_u32(&bars[0]), 0); mbar_wait(smem_u32(&bars[1]), 0); }
      if (variant & 16) { if (!ok0) mbar_wait(smem_u32(&bars[0]), 0); if (!ok1) mbar_wait(smem_u32(&bars[1]), 0);
                          ok0 = try_wait(smem_u32(&bars[0]), 0); ok1 = try_wait(smem_u32(&bars[1]), 0); }
      if (!(variant & 32)) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_pair = tmem + (i & 1) * 2 * N;
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mma_tf32_ts(d_pair, a_hi + j * 8, bdesc + 2 * j, idesc2N, (i >= 2) | (j != 0));
          mma_tf32_ts(d_pair + N, a_lo + j * 8, bdesc + 2 * j, idescN, 1u);
        }
        if (variant & 1) { tc_commit(smem_u32(&bars[2])); tc_commit(smem_u32(&bars[3])); }
        if (variant & 8) { tc_commit(smem_u32(&bars[2])); }
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (elect_one()) tc_commit(smem_u32(&bars[4]));
    __syncwarp();
    mbar_wait(smem_u32(&bars[4]), 0);
    long long t2 = clock64();
    if (lane == 0) { stop = 1; if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; } }
  } else if (warp >= 4 && (variant & 4)) {
    // converter stand-in: keep storing {hi, lo} rows into two other A slots until the issuer is done
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0x3f800000u + j;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    int t = 0;
    while (!stop) {
      tmem_st32(lane_base + 320 + t * 64, v);
      tmem_st32(lane_base + 320 + t * 64 + 32, v);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      t ^= 1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  struct { int v; const char* name; } vs[] = {
      {0, "8 MMAs only"}, {32, "8 MMAs, no fence"}, {3, "kernel's loop (2 commits + 2 waits)"}, {35, "kernel's loop, no fence"},
      {129, "2 commits + lane-0 waits"}, {161, "2 commits + lane-0 waits, no fence"},
      {64, "2 k-blocks/iter: 16 MMAs"}, {67, "2 k-blocks/iter: kernel's loop"}, {99, "2 k-blocks/iter: kernel's loop, no fence"},
      {71, "2 k-blocks/iter: kernel's loop, TMEM stores running"}};
  for (int N : {32, 64})
    for (auto& v : vs) {
      cudaMemset(d, 0, 16 * sizeof(long long));
      k<<<148, 256, 80 * 1024>>>(d, v.v, N);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("N=%3d %-50s issue %7.1f  complete %7.1f cycles per k-block  (%s)\n", N, v.name, (double)h[0] / KB, (double)h[1] / KB,
             cudaGetErrorString(e));
    }
  return 0;
}
