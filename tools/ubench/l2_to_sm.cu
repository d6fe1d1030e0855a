// Micro-benchmark (developer aid, round-2 entry; not yet run): sustained L2 -> shared-memory throughput of bulk copies with every
// SM loading at once - the wall behind the gather kernels once their internal pipelines are fixed (a k-block
// moves 24 / 32 / 48 KB at N = 32 / 64 / 128 against 206 / 384 / 768 cycles of tensor work).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/l2_to_sm tools/ubench/l2_to_sm.cu && build/l2_to_sm
// One CTA per SM, a ring of 4 x 32 KB stages filled by cp.async.bulk (one elected thread), no consumer work.  Sources:
//   distinct : every CTA streams its own slice of a 96 MB buffer that was just written (L2-resident, 126 MB L2)
//   shared   : every CTA streams the SAME 1.6 MB region (a layer's weight images): does the L2 serve duplicates cheaper?
//   mixed    : half of each stage from the shared region, half distinct (what a k-block of the N = 64 layers looks like)
// Reported: bytes per SM clock per SM and chip-wide, and TB/s at the measured clock.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int kStage = 32 * 1024, kStages = 4;

// mode 0 distinct, 1 shared, 2 mixed.  slice_bytes: this CTA's private region; shared_bytes: the common region
__global__ void __launch_bounds__(32, 1) k(const uint8_t* __restrict__ buf, size_t slice_bytes, size_t shared_bytes, int mode, int iters,
                                         long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) unsigned long long bars[kStages];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (threadIdx.x != 0) return;
  const uint8_t* mine = buf + shared_bytes + (size_t)blockIdx.x * slice_bytes;
  const uint8_t* common = buf;
  size_t off_m = 0, off_c = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters + kStages; ++i) {
    const int s = i % kStages;
    const uint32_t bar = smem_u32(&bars[s]);
    if (i >= kStages) {                                  // the copy issued kStages iterations ago has landed -> reuse its stage
      const uint32_t parity = ((i / kStages) - 1) & 1;
      while (!try_wait(bar, parity)) {}
    }
    if (i < iters) {
      mbar_expect_tx(bar, kStage);
      const uint32_t dst = base + s * kStage;
      if (mode == 0) {
        bulk_load(dst, mine + off_m, kStage, bar);
        off_m += kStage; if (off_m + kStage > slice_bytes) off_m = 0;
      } else if (mode == 1) {
        bulk_load(dst, common + off_c, kStage, bar);
        off_c += kStage; if (off_c + kStage > shared_bytes) off_c = 0;
      } else {
        bulk_load(dst, common + off_c, kStage / 2, bar);
        bulk_load(dst + kStage / 2, mine + off_m, kStage / 2, bar);
        off_c += kStage / 2; if (off_c + kStage / 2 > shared_bytes) off_c = 0;
        off_m += kStage / 2; if (off_m + kStage / 2 > slice_bytes) off_m = 0;
      }
    }
  }
  cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const size_t shared_bytes = 1600 * 1024;                       // 25 taps x 2 channel blocks x 32 KB
  const size_t slice_bytes = (size_t)(96u << 20) / sms / kStage * kStage;
  const size_t total = shared_bytes + slice_bytes * sms;
  uint8_t* buf; long long* cyc;
  cudaMalloc(&buf, total); cudaMalloc(&cyc, sms * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kStage * kStages + 128);
  const int iters = 4096;                                        // 128 MB per SM per run
  const char* names[3] = {"distinct", "shared  ", "mixed   "};
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {                          // rep 0 warms the L2 (cudaMemset leaves the lines resident)
      cudaMemset(buf, rep + 1, total);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      k<<<sms, 32, kStage * kStages + 128>>>(buf, slice_bytes, shared_bytes, mode, iters, cyc);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      long long h[512]; cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes_sm = (double)iters * kStage;
      if (rep == 1)
        printf("%s  %6.1f B/clk/SM  %7.0f B/clk chip  %6.2f TB/s (events)  [%d SMs, %.0f MHz nominal, slowest SM %lld cycles]\n", names[mode],
               bytes_sm / mx, bytes_sm * sms / mx, bytes_sm * sms / (ms * 1e-3) / 1e12, sms, khz / 1e3, mx);
    }
  }
  return 0;
}
