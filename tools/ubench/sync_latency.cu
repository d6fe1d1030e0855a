// Micro-benchmark (developer aid): issue-to-issue cost, in SM cycles, of the synchronisation primitives the tcgen05 pipelines
// use per k-block.  One CTA, measured by lane 0 of a converged warp with clock64 around REP back-to-back operations.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/sync_latency tools/ubench/sync_latency.cu && build/sync_latency
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
__device__ __forceinline__ uint32_t test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

constexpr int REP = 64;
__global__ void k(long long* out) {
  __shared__ __align__(8) unsigned long long bars[REP + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < REP + 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp != 0) return;
  uint32_t acc = 0;
  long long t0, t1;
  // complete phase 0 of every barrier
  if (lane == 0) for (int i = 0; i < REP + 8; ++i) mbar_arrive(smem_u32(&bars[i]));
  __syncwarp();
  // (0) 32-lane try_wait on completed barriers (distinct barriers, back to back)
  t0 = clock64();
  for (int i = 0; i < REP; ++i) acc += try_wait(smem_u32(&bars[i]), 0);
  t1 = clock64();
  if (lane == 0) out[0] = (t1 - t0) / REP;
  __syncwarp();
  // (1) single-lane try_wait on completed barriers
  t0 = clock64();
  if (lane == 0) for (int i = 0; i < REP; ++i) acc += try_wait(smem_u32(&bars[i]), 0);
  t1 = clock64();
  if (lane == 0) out[1] = (t1 - t0) / REP;
  __syncwarp();
  // (2) 32-lane test_wait
  t0 = clock64();
  for (int i = 0; i < REP; ++i) acc += test_wait(smem_u32(&bars[i]), 0);
  t1 = clock64();
  if (lane == 0) out[2] = (t1 - t0) / REP;
  __syncwarp();
  // (3) single-lane test_wait
  t0 = clock64();
  if (lane == 0) for (int i = 0; i < REP; ++i) acc += test_wait(smem_u32(&bars[i]), 0);
  t1 = clock64();
  if (lane == 0) out[3] = (t1 - t0) / REP;
  __syncwarp();
  // (4) dependent chain: try_wait result feeds the next address (true latency)
  t0 = clock64();
  {
    uint32_t idx = 0;
    for (int i = 0; i < REP; ++i) idx = (idx + try_wait(smem_u32(&bars[idx]), 0)) & (REP - 1);
    acc += idx;
  }
  t1 = clock64();
  if (lane == 0) out[4] = (t1 - t0) / REP;
  __syncwarp();
  // (5) same with test_wait
  t0 = clock64();
  {
    uint32_t idx = 0;
    for (int i = 0; i < REP; ++i) idx = (idx + test_wait(smem_u32(&bars[idx]), 0)) & (REP - 1);
    acc += idx;
  }
  t1 = clock64();
  if (lane == 0) out[5] = (t1 - t0) / REP;
  __syncwarp();
  // (6) single-lane mbarrier.arrive (phase 1 of each barrier)
  t0 = clock64();
  if (lane == 0) for (int i = 0; i < REP; ++i) mbar_arrive(smem_u32(&bars[i]));
  t1 = clock64();
  if (lane == 0) out[6] = (t1 - t0) / REP;
  __syncwarp();
  // (7) arrive -> visible to try_wait of the same thread (round trip), single lane
  t0 = clock64();
  if (lane == 0)
    for (int i = 0; i < REP; ++i) {                 // barriers are now in phase 2 (parity 0 again): arrive completes it
      mbar_arrive(smem_u32(&bars[i]));
      while (!try_wait(smem_u32(&bars[i]), 0)) {}
    }
  t1 = clock64();
  if (lane == 0) out[7] = (t1 - t0) / REP;
  __syncwarp();
  // (8) tcgen05.commit with nothing outstanding -> try_wait (round trip), single lane (barriers in phase 3, parity 1)
  t0 = clock64();
  if (lane == 0)
    for (int i = 0; i < REP; ++i) {
      tc_commit(smem_u32(&bars[i]));
      while (!try_wait(smem_u32(&bars[i]), 1)) {}
    }
  t1 = clock64();
  if (lane == 0) out[8] = (t1 - t0) / REP;
  __syncwarp();
  // (9) tcgen05.commit issue cost alone (phase 4, parity 0)
  t0 = clock64();
  if (lane == 0) for (int i = 0; i < REP; ++i) tc_commit(smem_u32(&bars[i]));
  t1 = clock64();
  if (lane == 0) out[9] = (t1 - t0) / REP;
  if (lane == 0) for (int i = 0; i < REP; ++i) while (!try_wait(smem_u32(&bars[i]), 0)) {}
  __syncwarp();
  // (10) clock64 pair overhead
  t0 = clock64();
  t1 = clock64();
  if (lane == 0) out[10] = t1 - t0;
  // (11) cross-warp signal latency is measured in the main kernels' traces
  if (lane == 0) out[15] = acc;
}

int main() {
  long long* d;
  cudaMalloc(&d, 16 * sizeof(long long));
  cudaMemset(d, 0, 16 * sizeof(long long));
  k<<<1, 64>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"try_wait x32 lanes (completed, independent)", "try_wait 1 lane", "test_wait x32 lanes", "test_wait 1 lane",
                         "try_wait dependent chain (latency)", "test_wait dependent chain (latency)", "mbarrier.arrive 1 lane",
                         "arrive -> try_wait round trip", "tcgen05.commit (idle pipe) -> try_wait round trip", "tcgen05.commit issue",
                         "clock64 pair"};
  printf("status: %s\n", cudaGetErrorString(e));
  for (int i = 0; i < 11; ++i) printf("%-55s %lld cycles\n", names[i], h[i]);
  return 0;
}
