// Hardware probe (developer aid, round 2): the operand forms the Form-W redesign (wgrad_ss) relies on, END TO END with TMA:
//   * NHWC tiles TMA-loaded with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B are valid MN-major tf32 operands (descriptor layout type 1)
//   * A: M = 128 = four 32-channel groups that are the SAME halo tile shifted by one pixel each (LBO = 128 bytes), the descriptor
//     starting at an arbitrary pixel row of a pitch-18 halo (start address bits [7,9) != 0), K = 8 pixels = two 4-row groups (SBO = 512)
//   * B: N = nb * Co = groups (row shift j, 32-channel block cb) in arithmetic progression (LBO = one tile row of one channel block)
//   D[(g, ci)][(j, cb, co)] = sum_{r, px} X[r][px + g + a0][ci] * O[r + j][cb][px][co]      (the two-sided shift of DESIGN 4.2)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/mn_probe tools/ubench/mn_probe.cu -lcuda && build/mn_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

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
__device__ __forceinline__ void mma_tf32_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

constexpr int RX = 4, RO = 6, OW = 16;             // X block: 4 rows x 18 pixels; O block: 6 rows x 16 pixels

struct P {
  int ncb, nb, a0, PXW;      // channel blocks of O, row shifts, first pixel shift of A, X halo pitch (pixels)
  uint32_t idesc;
};

__global__ void __launch_bounds__(128, 1) mn_probe(const __grid_constant__ CUtensorMap mx, const __grid_constant__ CUtensorMap mo, P p,
                                                   float* __restrict__ d_out, int* status) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) unsigned long long bar[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t xs = base, os = base + 16384;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int N = p.nb * p.ncb * 32;
  if (threadIdx.x == 32) {
    const uint32_t xb = RX * p.PXW * 128, ob = RO * p.ncb * OW * 128;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(xb + ob) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(xs), "l"(&mx), "r"(smem_u32(&bar[0])), "r"(0), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(os), "l"(&mo), "r"(smem_u32(&bar[0])), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
    uint32_t spins = 0;
    while (!try_wait(smem_u32(&bar[0]), 0)) { if (++spins > (1u << 24)) { *status = 2; __trap(); } }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // descriptors: MN-major SWIZZLE_128B_BASE32B (layout type 1), LBO = group stride, SBO = 512 (two 4-pixel K groups per instruction)
    const uint64_t da = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
    const uint64_t db = ((uint64_t)((OW * 128) >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
    int first = 1;
    for (int r = 0; r < RX; ++r)
      for (int q = 0; q < 2; ++q) {
        const uint32_t a_addr = xs + (uint32_t)((r * p.PXW + 8 * q + p.a0) * 128);
        const uint32_t b_addr = os + (uint32_t)((r * p.ncb * OW + 8 * q) * 128);
        mma_tf32_ss(tmem, da | (uint64_t)((a_addr & 0x3FFFF) >> 4), db | (uint64_t)((b_addr & 0x3FFFF) >> 4), p.idesc, first ? 0u : 1u);
        first = 0;
      }
    tc_commit(smem_u32(&bar[1]));
  }
  uint32_t spins = 0;
  while (!try_wait(smem_u32(&bar[1]), 0)) { if (++spins > (1u << 24)) { if (threadIdx.x == 0) *status = 1; __trap(); } }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) d_out[threadIdx.x * 256 + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }

int main() {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) { printf("no encode fn\n"); return 1; }
  EncodeTiledFn encode = (EncodeTiledFn)sym;
  srand(7);
  auto rnd = [] { return trunc_tf32((float)rand() / RAND_MAX * 2.f - 1.f); };
  for (int PXW : {24, 18})
  for (int nb = 1; nb <= 3; nb += 2)
  for (int ncb = 1; ncb <= 2; ++ncb)
    for (int a0 = 0; a0 < 2; ++a0) {
      const int swz = 0;
      const int Co = 32 * ncb, N = nb * Co;
      // global tensors: X [RX + 2][PXW + 6][32] (a larger plane; the box starts at (1, 2)), O [RO + 2][OW + 4][Co]
      const int XH = RX + 2, XW = 24 + 6, OH = RO + 2, OWG = OW + 4;
      std::vector<float> X(XH * XW * 32), O((size_t)OH * OWG * Co);
      for (auto& v : X) v = rnd();
      for (auto& v : O) v = rnd();
      float *dX, *dO, *dD; int* dS;
      cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dO, O.size() * 4); cudaMalloc(&dD, 128 * 256 * 4); cudaMalloc(&dS, 4);
      cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dO, O.data(), O.size() * 4, cudaMemcpyHostToDevice);
      cudaMemset(dD, 0xff, 128 * 256 * 4); cudaMemset(dS, 0, 4);
      const CUtensorMapSwizzle mode = swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
      CUtensorMap mx, mo;
      {
        cuuint64_t dims[3] = {32, (cuuint64_t)XW, (cuuint64_t)XH}, strides[2] = {128, (cuuint64_t)XW * 128};
        cuuint32_t box[3] = {32, (cuuint32_t)PXW, RX}, es[3] = {1, 1, 1};
        CUresult cr = encode(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dX + (1 * XW + 2) * 32, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             mode, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { printf("encode X failed %d\n", (int)cr); return 1; }
      }
      {
        cuuint64_t dims[4] = {32, (cuuint64_t)OWG, (cuuint64_t)ncb, (cuuint64_t)OH}, strides[3] = {(cuuint64_t)Co * 4, 128, (cuuint64_t)OWG * Co * 4};
        cuuint32_t box[4] = {32, OW, (cuuint32_t)ncb, RO}, es[4] = {1, 1, 1, 1};
        CUresult cr = encode(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dO + ((size_t)1 * OWG + 2) * Co, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             mode, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { printf("encode O failed %d\n", (int)cr); return 1; }
      }
      P p{ncb, nb, a0, PXW, (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24)};
      cudaFuncSetAttribute(mn_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      mn_probe<<<1, 128, 64 * 1024>>>(mx, mo, p, dD, dS);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("swizzle %s ncb %d a0 %d: CUDA error %s\n", swz ? "128B" : "128B_ATOM_32B", ncb, a0, cudaGetErrorString(e)); cudaDeviceReset(); continue; }
      std::vector<float> D(128 * 256);
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      double mx_ref = 0, worst = 0, blk[4][8] = {};
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          const int g = m >> 5, ci = m & 31, j = n / Co, cb = (n % Co) / 32, co = n % 32;
          double s = 0;
          for (int r = 0; r < RX; ++r)
            for (int px = 0; px < 16; ++px)
              s += (double)X[((1 + r) * XW + 2 + px + g + a0) * 32 + ci] * O[((size_t)(1 + r + j) * OWG + 2 + px) * Co + cb * 32 + co];
          mx_ref = fmax(mx_ref, fabs(s));
          const double d = fabs((double)D[m * 256 + n] - s);
          if (!(d <= worst)) worst = d;
          if (d > blk[g][n / 32]) blk[g][n / 32] = d;
        }
      printf("pitch %d nb %d ncb %d (N = %3d) a0 %d: rel-err %.3e  %s   per (M group g, N group): ", PXW, nb, ncb, N, a0, worst / mx_ref, worst / mx_ref < 1e-5 ? "PASS" : "FAIL");
      for (int g = 0; g < 4; ++g) { for (int n = 0; n < N / 32; ++n) printf("%c", blk[g][n] / mx_ref < 1e-5 ? '.' : 'X'); printf(" "); }
      printf("\n");
      cudaFree(dX); cudaFree(dO); cudaFree(dD); cudaFree(dS);
    }
  return 0;
}
