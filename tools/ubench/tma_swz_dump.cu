// Developer aid: dumps the shared-memory byte order a TMA tile load produces under each 128-byte swizzle mode
// (tile = 16 rows x 32 floats, element value = row * 32 + col) so the MN-major descriptor layout can be matched to it.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void dump(const __grid_constant__ CUtensorMap m, float* out, int nfloats, uint32_t off) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  float* gen = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nfloats * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(base + off), "l"(&m), "r"(smem_u32(&bar)), "r"(0), "r"(0) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = gen[off / 4 + i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* sym = nullptr; cudaDriverEntryPointQueryResult q; cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeTiledFn encode = (EncodeTiledFn)sym;
  const int R = 16, n = R * 32;
  std::vector<float> h(n); for (int i = 0; i < n; ++i) h[i] = (float)i;
  float *dg, *dout; cudaMalloc(&dg, n * 4); cudaMalloc(&dout, n * 4); cudaMemcpy(dg, h.data(), n * 4, cudaMemcpyHostToDevice);
  const char* names[] = {"128B", "128B_ATOM_32B", "128B_ATOM_32B_FLIP_8B", "128B_ATOM_64B"};
  const CUtensorMapSwizzle modes[] = {CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B_FLIP_8B, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_64B};
  for (int mi = 0; mi < 4; ++mi)
    for (uint32_t off : {0u, 256u}) {
      CUtensorMap m; cuuint64_t dims[2] = {32, R}, strides[1] = {128}; cuuint32_t box[2] = {32, R}, es[2] = {1, 1};
      CUresult cr = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dg, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, modes[mi], CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) { printf("%s: encode failed %d\n", names[mi], (int)cr); continue; }
      cudaFuncSetAttribute(dump, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
      dump<<<1, 128, 32 * 1024>>>(m, dout, n, off);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: launch failed\n", names[mi]); cudaDeviceReset(); return 1; }
      std::vector<float> o(n); cudaMemcpy(o.data(), dout, n * 4, cudaMemcpyDeviceToHost);
      printf("mode %s, tile at 1024-aligned base + %u: for each smem row (128 B), the source 16-byte chunk index found at chunk position 0..7 (source row in brackets if different)\n", names[mi], off);
      for (int r = 0; r < 8; ++r) {
        printf("  smem row %d:", r);
        for (int c = 0; c < 8; ++c) { int v = (int)o[r * 32 + c * 4]; int sr = v / 32, sc = (v % 32) / 4; if (sr == r) printf(" %d", sc); else printf(" %d[%d]", sc, sr); }
        printf("\n");
      }
    }
  return 0;
}
