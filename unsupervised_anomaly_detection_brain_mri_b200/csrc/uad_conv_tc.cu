// tcgen05 / TMEM / TMA implicit-GEMM kernel for the conv family (Form F and Form T of uad_conv.cuh), sm_100a only.
//
//   D[128 pixels, N] += A[128 pixels, 32 ch] . B[32 ch, N]     per k-block (one filter tap x one 32-channel block)
//
// fp32 parity on tensor cores: 3xTF32.  Every fp32 operand x is split exactly into hi = x & 0xffffe000 (tf32-exact)
// and lo = x - hi; the product is accumulated as lo*hi + hi*lo + hi*hi in the fp32 TMEM accumulator (dropped lo*lo
// term ~2^-22 relative).  Weights are split once per call by a prep kernel into pre-swizzled smem images; the
// activation tile is split on the fly by a converter warpgroup that moves it from shared memory INTO TENSOR MEMORY, so
// the three MMA passes read A from TMEM (tcgen05.mma "ts" form) and only B from shared memory - in "ss" form the
// A re-reads would make the kernel shared-memory-bandwidth bound at N <= 128.
//
// Pipeline (one 128-pixel tile per CTA, up to 2 CTAs per SM so one CTA's epilogue overlaps the other's main loop):
//   warp 0      TMA producer: 5-D tiled tensor map on the NHWC input (im2col by coordinates; OOB zero fill == SAME padding),
//               + one bulk copy of the pre-swizzled {hi,lo} weight image                          -> full[s]
//   warps 4-7   converters: smem row -> registers -> hi/lo -> tcgen05.st into TMEM A slot t        -> afull[t]
//   warp 1      MMA issuer (one thread): 12 x tcgen05.mma.kind::tf32 per k-block, tcgen05.commit   -> empty[s], aempty[t]
//   warps 4-7   epilogue: tcgen05.ld accumulator -> +bias, frozen-BN affine, activation -> smem transpose ->
//               coalesced 512-byte row stores of z and/or a
#include <cuda.h>
#include <stdlib.h>

#include "uad_conv.cuh"

namespace {

constexpr int kTileM = 128;
constexpr int kKBlk = 32;                    // fp32 channels per k-block = one 128-byte swizzle row
constexpr int kABytes = kTileM * kKBlk * 4;  // 16 KB
// TMEM columns: G "main" accumulators (hi*hi products, k-blocks dealt round-robin) + 1 "correction" accumulator
// (lo*hi + hi*lo) of N columns each, then two {hi,lo} A slots of 64 columns.  The tensor core adds into its fp32
// accumulator with truncation (measured: error grows linearly with the number of accumulations), so the large hi*hi
// stream is spread over G accumulators and the small correction terms never disturb it; the epilogue sums them in
// registers with round-to-nearest.
// developer trace (UAD_TC_DEBUG bit 16): clock64 stamps of one steady-state CTA of the N = 32 kernel
__device__ long long g_tc_trace[64];
constexpr uint32_t kSpinLimit = 1u << 22;    // bounded mbarrier spins: trap instead of hanging the GPU

struct TcParams {
  int TW, TH, TB, lgTW, lgTH;
  int tiles_w, tiles_h;
  int B, C, Cblks, N;
  int OH, OW, osh;
  int stride2;
  int stages;
  int G;               // number of main accumulators (1 or 2)
  int tmem_cols;       // 256 or 512
  int nacc;            // accumulators of N columns: 2G (paired layout, N <= 64) or G + 1 (N = 128)
  int nslots;          // TMEM A slots (2..4)
  int acc_bufs;        // accumulator sets in TMEM (2 = epilogue overlaps the next item's MMAs)
  int n_items;         // work items = tiles * nclasses
  int nclasses;
  int debug;           // developer timing switches (UAD_TC_DEBUG): 1 = converters skip their work, 2 = MMA issuer skips the MMAs,
                       // 4 / 8 = N=32 kernel: no global stores / no epilogue pass, 16 = clock64 trace, 32 / 64 = no A / no B load
  float* z_out;
  float* a_out;
  const float* bias;
  const float* gamma;
  const float* beta;
  float bn_c, alpha;
  int act;
  const float* wimg;   // [k*k][Cblks][2][N][32] pre-swizzled
  TapSet taps[4];
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && spin > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// warp-converged single-lane election (elect.sync): keeps the surrounding values provably warp-uniform so the compiler
// builds MMA descriptors in uniform registers instead of R2UR-ing them per instruction inside a divergent branch
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
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
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 bytes, 8-row groups 1024 bytes apart)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);         // start address  [0,14)
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset = 1024 B [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (sm_100) [46,48)
  d |= (uint64_t)2 << 61;                              // layout type SWIZZLE_128B [61,64)
  return d;
}

// ------------------------------------------------------------------------------------------------ the kernel
// PERSISTENT: one CTA per SM walks work items (tile, output-parity class) = blockIdx.x + k * gridDim.x; the smem ring,
// the TMEM A-slot ring and the mbarrier phases run on across items, so the TMA producer prefetches the next item's
// operands while the epilogue warps drain the previous accumulators (double-buffered in TMEM when they fit).
//
// TMEM columns: `acc_bufs` accumulator sets of `nacc` x N columns, then `nslots` {hi,lo} A slots of 64 columns.
//   N <= 64 : accumulator pairs [main_g | corr_g], g < G.  Per K=8 slice TWO instructions:
//             (a_hi) x [B_hi ; B_lo]  as ONE 2N-wide MMA into [main_g | corr_g]   (hi*hi and hi*lo at once)
//             (a_lo) x  B_hi          as an N-wide MMA into corr_g
//   N = 128 : [main_0 .. main_{G-1} | corr], three N-wide MMAs per slice.
// Why several accumulators: the tensor core adds into its fp32 accumulator with truncation (measured: error grows
// linearly with the number of accumulations), so the large hi*hi stream is dealt round-robin over G accumulators and
// the small correction terms never disturb it; the epilogue sums all of them in registers with round-to-nearest.
//
// 12 warps: 0 TMA producer | 1 MMA issuer | 2 TMEM alloc | 3 epilogue constants | 4-7 converters | 8-11 epilogue.
__global__ void __launch_bounds__(384, 1)
gather_gemm_tc(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int N = p.N;
  const uint32_t b_bytes = 2u * N * 128u;
  const uint32_t stage_bytes = kABytes + b_bytes;
  const int S = p.stages;
  const int NS = p.nslots;
  // bookkeeping lives after the pipeline stages
  const uint32_t misc = smem_base + S * stage_bytes;
  const uint32_t bar_full = misc;                       // S x 8   (S <= 8)
  const uint32_t bar_empty = misc + 64;                 // S x 8
  const uint32_t bar_afull = misc + 128;                // NS x 8
  const uint32_t bar_aempty = misc + 160;               // NS x 8
  const uint32_t bar_accfull = misc + 192;              // 2 x 8
  const uint32_t bar_accempty = misc + 208;             // 2 x 8
  const uint32_t tmem_slot = misc + 224;
  float* epi = reinterpret_cast<float*>(smem_gen + S * stage_bytes + 256);               // bias[N], scale[N], shift[N]
  float* stg_base = reinterpret_cast<float*>(smem_gen + S * stage_bytes + 256 + 3 * N * 4);   // 4 x 32 x (N+4) staging

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t aoff = p.acc_bufs * p.nacc * N;        // first A slot column
  const int nclasses = p.nclasses;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < NS; ++i) { mbar_init(bar_afull + 8 * i, 128); mbar_init(bar_aempty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_accfull + 8 * i, 1); mbar_init(bar_accempty + 8 * i, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 3) {
    for (int n = lane; n < N; n += 32) {
      epi[n] = p.bias ? p.bias[n] : 0.f;
      epi[N + n] = p.gamma ? p.gamma[n] * p.bn_c : 1.f;
      epi[2 * N + n] = p.beta ? p.beta[n] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // all ring indices / phase bits are carried incrementally: no runtime div/mod in the per-k-block loops
  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      int s = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int cls = item % nclasses, tile = item / nclasses;
        const TapSet& ts = p.taps[cls];
        const int twi = tile % p.tiles_w, thi = (tile / p.tiles_w) % p.tiles_h, tbi = tile / (p.tiles_w * p.tiles_h);
        const int s0 = twi * p.TW, r0 = thi * p.TH, b0 = tbi * p.TB;
        const int nkb = ts.n * p.Cblks;
        int tap = 0, cb = 0;
        for (int i = 0; i < nkb; ++i) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          const uint32_t full = bar_full + 8 * s;
          // developer timing switches: 32 = skip the activation (A) load, 64 = skip the weight (B) load
          mbar_expect_tx(full, ((p.debug & 32) ? 0u : (uint32_t)kABytes) + ((p.debug & 64) ? 0u : b_bytes));
          const int dh = ts.dh[tap], dw = ts.dw[tap], wt = ts.wt[tap];
          const uint32_t a_dst = smem_base + s * stage_bytes;
          if (p.debug & 32) {
          } else if (p.stride2)
            tma_load_5d(a_dst, &tmap, full, (dw & 1) * p.C + cb * kKBlk, s0 + (dw >> 1), dh & 1, r0 + (dh >> 1), b0);
          else
            tma_load_5d(a_dst, &tmap, full, cb * kKBlk, s0 + dw, 0, r0 + dh, b0);
          const float* wsrc = p.wimg + ((size_t)(wt * p.Cblks + cb)) * 2 * N * kKBlk;
          if (!(p.debug & 64)) bulk_load(a_dst + kABytes, wsrc, b_bytes, full);
          if (++cb == p.Cblks) { cb = 0; ++tap; }
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    {
      // the whole warp walks the loop (converged); one elected lane issues.
      // instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t idescN = idesc_base | ((uint32_t)(N >> 3) << 17);
      const uint32_t idesc2N = idesc_base | ((uint32_t)((2 * N) >> 3) << 17);
      const uint64_t bdesc0 = make_sw128_desc(smem_base + kABytes);          // B image of stage 0 (hi rows then lo rows)
      const uint32_t stage_units = stage_bytes >> 4, lo_units = (uint32_t)(N * 128) >> 4;
      const bool paired = (N <= 64);
      int s = 0, t = 0, buf = 0;
      uint32_t ph = 0, pht = 0, phb = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int nkb = p.taps[item % nclasses].n * p.Cblks;
        mbar_wait(bar_accempty + 8 * buf, phb ^ 1);             // epilogue has drained this accumulator set
        tc_fence_after();
        const uint32_t acc0 = tmem_base + buf * p.nacc * N;
        int g = 0;
        for (int i = 0; i < nkb; ++i) {
          mbar_wait(bar_full + 8 * s, ph);                      // weight image landed (async proxy -> visible)
          mbar_wait(bar_afull + 8 * t, pht);                    // converters filled TMEM A slot t
          tc_fence_after();
          const uint64_t dhi0 = bdesc0 + (uint64_t)(s * stage_units);
          const uint32_t a_hi = tmem_base + aoff + t * 64;
          const uint32_t a_lo = a_hi + 32;
          const uint32_t first = (i >= p.G) ? 1u : 0u;          // accumulator g already holds a partial sum of this item?
          if (elect_one()) {
            if (p.debug & 2) {
            } else if (paired) {
              const uint32_t d_pair = acc0 + g * 2 * N;
#pragma unroll
              for (int j = 0; j < 4; ++j) {                     // K = 8 tf32 per instruction -> 32 bytes (2 x 16 B) along the row
                mma_tf32_ts(d_pair, a_hi + j * 8, dhi0 + 2 * j, idesc2N, first | (j != 0));
                mma_tf32_ts(d_pair + N, a_lo + j * 8, dhi0 + 2 * j, idescN, 1u);
              }
            } else {
              const uint32_t d_main = acc0 + g * N;
              const uint32_t d_corr = acc0 + p.G * N;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                mma_tf32_ts(d_corr, a_lo + j * 8, dhi0 + 2 * j, idescN, (i | j) != 0);
                mma_tf32_ts(d_corr, a_hi + j * 8, dhi0 + lo_units + 2 * j, idescN, 1u);
                mma_tf32_ts(d_main, a_hi + j * 8, dhi0 + 2 * j, idescN, first | (j != 0));
              }
            }
            tc_commit(bar_empty + 8 * s);                       // smem stage reusable once these MMAs retire
            tc_commit(bar_aempty + 8 * t);                      // TMEM A slot reusable
            if (i == nkb - 1) tc_commit(bar_accfull + 8 * buf); // accumulators of this item complete
          }
          __syncwarp();
          if (++s == S) { s = 0; ph ^= 1; }
          if (++t == NS) { t = 0; pht ^= 1; }
          if (++g == p.G) g = 0;
        }
        if (++buf == p.acc_bufs) { buf = 0; phb ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================================================================== converters
    const int row = threadIdx.x - 128;                          // tile row == TMEM lane
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    int s = 0, t = 0;
    uint32_t ph = 0, pht = 0;
    const uint32_t swz = (uint32_t)(row & 7);
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int nkb = p.taps[item % nclasses].n * p.Cblks;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(bar_full + 8 * s, ph);
        const uint8_t* arow = smem_gen + s * stage_bytes + row * 128;
        uint32_t hi[32], lo[32];
        if (p.debug & 1) {
          mbar_wait(bar_aempty + 8 * t, pht ^ 1);
          mbar_arrive(bar_afull + 8 * t);
          if (++s == S) { s = 0; ph ^= 1; }
          if (++t == NS) { t = 0; pht ^= 1; }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {                           // 16-byte chunk j of this row sits at (j ^ (row & 7))
          const float4 v = *reinterpret_cast<const float4*>(arow + ((j ^ swz) << 4));
          const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t h = __float_as_uint(f[e]) & 0xffffe000u;
            hi[4 * j + e] = h;
            lo[4 * j + e] = __float_as_uint(f[e] - __uint_as_float(h));
          }
        }
        mbar_wait(bar_aempty + 8 * t, pht ^ 1);
        tc_fence_after();
        const uint32_t a_slot = lane_base + aoff + t * 64;
        tmem_st32(a_slot, hi);
        tmem_st32(a_slot + 32, lo);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bar_afull + 8 * t);
        if (++s == S) { s = 0; ph ^= 1; }
        if (++t == NS) { t = 0; pht ^= 1; }
      }
    }
  } else if (warp >= 8) {
    // ===================================================================== epilogue warps
    const int row = threadIdx.x - 256;                          // tile row == TMEM lane
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int ldw = N + 4;
    float* stg = stg_base + (size_t)q * 32 * ldw;               // this warp's 32 x (N+4) staging rows
    const int lanes_per_row = N / 4;                            // float4 lanes covering one output row
    const int rows_per_it = 32 / lanes_per_row;
    const int npass = (p.z_out ? 1 : 0) + (p.a_out ? 1 : 0);
    int buf = 0;
    uint32_t phb = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int cls = item % nclasses, tile = item / nclasses;
      const TapSet& ts = p.taps[cls];
      const int twi = tile % p.tiles_w, thi = (tile / p.tiles_w) % p.tiles_h, tbi = tile / (p.tiles_w * p.tiles_h);
      const int s0 = twi * p.TW, r0 = thi * p.TH, b0 = tbi * p.TB;
      // output pixel of THIS thread's row (shuffled to the storing lanes below)
      const int tw = row & (p.TW - 1);
      const int th = (row >> p.lgTW) & (p.TH - 1);
      const int tb = row >> (p.lgTW + p.lgTH);
      const int b = b0 + tb;
      const long long my_off = (b < p.B)
          ? (((long long)b * p.OH + ((r0 + th) * p.osh + ts.oh0)) * p.OW + ((s0 + tw) * p.osh + ts.ow0)) * (long long)N
          : -1;
      mbar_wait(bar_accfull + 8 * buf, phb);
      tc_fence_after();
      const uint32_t acc0 = lane_base + buf * p.nacc * N;
      int done = 0;
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        float* out = pass == 0 ? p.z_out : p.a_out;
        if (!out) continue;
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t v[32], u[32];
          tmem_ld32(acc0 + c0, v);
          for (int k = 1; k < p.nacc; ++k) {
            tmem_ld32(acc0 + k * N + c0, u);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
          }
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int n = c0 + j + e;
              const float z = __uint_as_float(v[j + e]) + epi[n];
              o[e] = pass == 0 ? z : uad_act(epi[N + n] * z + epi[2 * N + n], p.act, p.alpha);
            }
            *reinterpret_cast<float4*>(stg + lane * ldw + c0 + j) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        if (++done == npass) {                                  // last TMEM read of this item: hand the accumulators back
          tc_fence_before();
          mbar_arrive(bar_accempty + 8 * buf);
        }
        __syncwarp();
        {
          // batches of 8 store instructions: all row offsets and smem reads first, then the global stores (the
          // dependent shfl -> LDS -> STG chain per row group is latency-bound otherwise)
          const int c = (lane % lanes_per_row) * 4;
          for (int rr = 0; rr < 32; rr += 8 * rows_per_it) {
            long long offs[8];
            float4 vals[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int r = (rr + it * rows_per_it + lane / lanes_per_row) & 31;
              offs[it] = __shfl_sync(0xffffffffu, my_off, r);
              vals[it] = *reinterpret_cast<const float4*>(stg + r * ldw + c);
            }
#pragma unroll
            for (int it = 0; it < 8; ++it)
              if (rr + it * rows_per_it < 32 && offs[it] >= 0) *reinterpret_cast<float4*>(out + offs[it] + c) = vals[it];
          }
        }
        __syncwarp();
      }
      if (++buf == p.acc_bufs) { buf = 0; phb ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ the kernel, v2
// Same tiles, rings, numerics and epilogue math as gather_gemm_tc, different ROLE STRUCTURE, built on this round's
// measurements: the tensor pipe sat 25-45 % busy because ONE issuing warp serialises its barrier waits,
// a blocking MMA issue and its commits, and ONE converter warpgroup serialises waits, smem reads and TMEM stores, while
// the four dedicated epilogue warps idle (no VAE-256 layer fits two accumulator sets in TMEM, so the epilogue never
// overlapped the MMAs anyway).  Here
//   * TWO converter warpgroups (warps 4-7 / 8-11) fill alternate k-blocks' TMEM A slots, and share the epilogue
//     (alternate 32-column chunks) once the item's accumulators are complete;
//   * TWO MMA-issuing warps (1 / 2) issue alternate k-blocks into their OWN accumulator pair [main_g | corr_g]
//     (N <= 64; the unpaired N = 128 layout has one shared correction accumulator, so there a single warp issues);
//     while one warp polls its barriers and commits, the other's MMAs keep the pipe busy.
// 12 warps: 0 TMA producer | 1 issuer A | 2 TMEM alloc, then issuer B | 3 epilogue constants | 4-7, 8-11 converter groups.
// ------------------------------------------------------------------------------------------------ weight images
// raw weights -> per (tap, 32-channel block): {hi, lo} images of [N rows][32 k] fp32 in the SWIZZLE_128B byte order the
// UMMA descriptor expects (16-byte chunk index XOR (row & 7)).  transposed=false: raw[t][c][n]; true: raw[t][n][c].
__global__ void weight_image_kernel(const float* __restrict__ w, float* __restrict__ img, int taps, int C, int N, int transposed) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)taps * C * N;
  if (i >= total) return;
  const int k = i % kKBlk;                 // channel within block
  const int n = (i / kKBlk) % N;
  const int cb = (i / ((size_t)kKBlk * N)) % (C / kKBlk);
  const int t = i / ((size_t)C * N);
  const int c = cb * kKBlk + k;
  const float v = transposed ? w[((size_t)t * N + n) * C + c] : w[((size_t)t * C + c) * N + n];
  const uint32_t h = __float_as_uint(v) & 0xffffe000u;
  const float lo = v - __uint_as_float(h);
  const size_t base = ((size_t)(t * (C / kKBlk) + cb)) * 2 * N * kKBlk;
  const int pos = n * kKBlk + ((((k >> 2) ^ (n & 7)) << 2) | (k & 3));
  img[base + pos] = __uint_as_float(h);
  img[base + (size_t)N * kKBlk + pos] = lo;
}

// ================================================================================================ Form W (wgrad)
// dW[(t, c), co] = sum_pix G[gather(pix, t), c] * O[pix, co]      (stride-2 gather; G fine tensor, O coarse tensor)
//
// GEMM roles: M = 128 rows = 4 "quads" (tap t, 32-channel block cb) x 32 channels, N = Co, K = pixels.
//   * one CTA owns a channel block cb, a tap range (<= 4*QT taps -> QT accumulator tiles in TMEM) and a range of
//     32-pixel blocks (split-K over pixels; partials reduced deterministically afterwards)
//   * per pixel block (4 x 8 coarse pixels of one image) TMA loads the four stride-2 parity planes of the 6 x 10 halo
//     of G ONCE (all taps gather from it) and the O tile
//   * converters (warps 4-7): split the O tile into tf32 hi / lo in shared memory (B operand, MN-major SW128), and per
//     accumulator tile gather-transpose 32 pixels x 32 channels per quad from the halo into a TMEM A slot as hi / lo
//   * MMA issuer: 12 x tcgen05.mma.kind::tf32 (A from TMEM, B MN-major from smem) per (pixel block, tile)
constexpr int kWPH = 4, kWPW = 8;                    // pixel block
constexpr int kHaloPitch = 16;                       // halo rows are loaded 16 pixels wide (10 needed): with a pitch that is a
                                                     // multiple of 8 the 128B-swizzle phase of a gathered pixel depends only on
                                                     // its column -> 8 precomputed bases + immediate offsets per tile
constexpr int kHaloRows = (kWPH + 2) * kHaloPitch;   // 96 rows of 128 B per parity plane
constexpr int kPlaneBytes = kHaloRows * 128;         // 12 KB (multiple of 1024)
constexpr int kHaloBytes = 4 * kPlaneBytes;

struct TcWgradParams {
  int B, lgMH, lgMW;         // coarse (M-grid) dims
  int Cg, Co, ncb;           // gathered channels, other channels, Cg/32
  int ngroups, taps_per_group, QT;
  int nblocks;               // number of 32-pixel blocks = B * (MH/4) * (MW/8)
  int blocks_per_chunk;
  int Mp;                    // 25 * Cg
  int stages;
  int debug;                 // developer switch (UAD_WGRAD_DEBUG): 1 = A forced to 1.0
  float* partial;            // [nchunks][Mp][Co]
  signed char dh[UAD_MAX_TAPS], dw[UAD_MAX_TAPS];
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// (An MN-major SWIZZLE_128B tf32 B descriptor - K rows of 128 bytes straight from TMA - reads back zeros on B200:
// 32-bit MN-major operands need the 32-byte-atom swizzle.  The O tile is therefore re-laid out K-major by the converters.)
//
// 12 warps: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM alloc, 3 idle, 4-7 = converter group 0, 8-11 = converter group 1.
// Accumulator tiles are converted alternately by the two groups (tile u -> group u & 1), each group owning two TMEM A
// slots, so one group's smem->TMEM latency chain overlaps the other's and the MMAs of the tile in between.
__global__ void __launch_bounds__(384, 1)
wgrad_tc(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_o,
         const __grid_constant__ TcWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int Co = p.Co;
  const uint32_t o_bytes = 32u * Co * 4u;                 // raw O tile [32 px][Co]; then K-major B_hi, B_lo [Co][32 px]
  const uint32_t stage_bytes = kHaloBytes + 3 * o_bytes;
  const int S = p.stages;
  const uint32_t misc = smem_base + S * stage_bytes;
  const uint32_t bar_full = misc, bar_empty = misc + 64, bar_oready = misc + 128, bar_afull = misc + 192,
                 bar_aempty = misc + 224, bar_acc = misc + 256, tmem_slot = misc + 264;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int QT = p.QT;
  // work assignment
  const int grp = blockIdx.y % p.ngroups;
  const int cb = blockIdx.y / p.ngroups;
  const int t0 = grp * p.taps_per_group;
  const int ntaps = min(p.taps_per_group, 25 - t0);
  const int blk_begin = blockIdx.x * p.blocks_per_chunk;
  const int blk_end = min(p.nblocks, blk_begin + p.blocks_per_chunk);
  const int nkb = blk_end - blk_begin;
  const int bw = (1 << p.lgMW) / kWPW, bh = (1 << p.lgMH) / kWPH;      // pixel blocks per image row / column

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); mbar_init(bar_oready + 8 * i, 128); }
    for (int i = 0; i < 4; ++i) { mbar_init(bar_afull + 8 * i, 128); mbar_init(bar_aempty + 8 * i, 1); }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t aoff = QT * Co;                       // four A slots after the QT accumulator tiles

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_g) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_o) : "memory");
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const uint32_t full = bar_full + 8 * s;
        mbar_expect_tx(full, (uint32_t)kHaloBytes + o_bytes);
        const int blk = blk_begin + i;
        const int bx = blk % bw, by = (blk / bw) % bh, b = blk / (bw * bh);
        const int r0 = by * kWPH, s0 = bx * kWPW;
        const uint32_t st_base = smem_base + s * stage_bytes;
        for (int pl = 0; pl < 4; ++pl)                 // plane (ph, pw) = (pl >> 1, pl & 1)
          tma_load_5d(st_base + pl * kPlaneBytes, &tmap_g, full, (pl & 1) * p.Cg + cb * 32, s0 - 1, pl >> 1, r0 - 1, b);
        for (int a = 0; a < Co / 32; ++a)
          tma_load_4d(st_base + kHaloBytes + a * 4096, &tmap_o, full, a * 32, s0, r0, b);
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    {
      // whole warp converged, one elected lane issues.  D=f32, A=B=tf32, both K-major, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Co >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t bdesc0 = make_sw128_desc(smem_base + kHaloBytes + o_bytes);
      const uint32_t stage_units = stage_bytes >> 4, lo_units = o_bytes >> 4;
      int s = 0;
      uint32_t ph = 0, u = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(bar_full + 8 * s, ph);
        mbar_wait(bar_oready + 8 * s, ph);                    // O tile transposed + split (generic writes fenced to async proxy)
        const uint64_t dhi0 = bdesc0 + (uint64_t)(s * stage_units);
        for (int mt = 0; mt < QT; ++mt, ++u) {
          const uint32_t slot = ((u & 1) << 1) | ((u >> 1) & 1);   // group (u & 1), its slot ((u >> 1) & 1)
          mbar_wait(bar_afull + 8 * slot, (u >> 2) & 1);
          tc_fence_after();
          const uint32_t a_hi = tmem_base + aoff + slot * 64, a_lo = a_hi + 32;
          const uint32_t d = tmem_base + mt * Co;
          if (elect_one()) {
          if (!(p.debug & 4))
#pragma unroll
          for (int j = 0; j < 4; ++j) {                        // 8 pixels per instruction = 32 bytes along the swizzle row
            mma_tf32_ts(d, a_lo + j * 8, dhi0 + 2 * j, idesc, (i | j) != 0);
            mma_tf32_ts(d, a_hi + j * 8, dhi0 + lo_units + 2 * j, idesc, 1u);
            mma_tf32_ts(d, a_hi + j * 8, dhi0 + 2 * j, idesc, 1u);
          }
          tc_commit(bar_aempty + 8 * slot);
          }
          __syncwarp();
        }
        if (elect_one()) tc_commit(bar_empty + 8 * s);
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1; }
      }
      if (elect_one()) tc_commit(bar_acc);
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int cg = (warp - 4) >> 2;                            // converter group 0 / 1
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t chunk_swz = (uint32_t)(lane >> 2), word = (uint32_t)(lane & 3) << 2;
    int s = 0;
    uint32_t ph = 0, u = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(bar_full + 8 * s, ph);
      uint8_t* st = smem_gen + s * stage_bytes;
      // ---- O tile (group 0): transpose [32 px][Co] (TMA, pixel rows) -> K-major B_hi / B_lo [Co rows][32 px] (the layout
      //      the fwd kernel's weight images use), splitting into tf32 hi / lo on the way.  lane = channel, 4 pixels per
      //      store: reads are one 128-byte row per warp, STS.128 quarter-warps hit 8 distinct swizzle chunks.
      if (cg == 0) {
        const uint8_t* raw = st + kHaloBytes;
        uint8_t* bhi = st + kHaloBytes + o_bytes;
        uint8_t* blo = bhi + o_bytes;
        for (int a = 0; a < Co / 32; ++a) {
          const int n = a * 32 + lane;
#pragma unroll
          for (int g2 = 0; g2 < 2; ++g2) {
            const int pg = 2 * q + g2;                    // pixel group: pixels 4*pg .. 4*pg+3
            float h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int px = 4 * pg + e;
              const float v = *reinterpret_cast<const float*>(raw + a * 4096 + px * 128 + (((chunk_swz ^ (px & 7)) << 4) | word));
              h[e] = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
              l[e] = v - h[e];
            }
            const uint32_t off = n * 128 + ((pg ^ (n & 7)) << 4);
            *reinterpret_cast<float4*>(bhi + off) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(blo + off) = make_float4(l[0], l[1], l[2], l[3]);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(bar_oready + 8 * s);
      }
      // ---- A tiles: quad q of accumulator tile mt is tap (t0 + 4*mt + q); lane = channel within the block
      for (int mt = 0; mt < QT; ++mt, ++u) {
        if ((int)(u & 1) != cg) continue;
        const uint32_t slot = ((u & 1) << 1) | ((u >> 1) & 1);
        const int tap = 4 * mt + q;
        uint32_t hi[32], lo[32];
        if (p.debug & 2) {
          mbar_wait(bar_aempty + 8 * slot, ((u >> 2) & 1) ^ 1);
          mbar_arrive(bar_afull + 8 * slot);
          continue;
        }
        if (tap < ntaps) {
          const int dh = p.dh[t0 + tap], dw = p.dw[t0 + tap];
          const uint8_t* plane = st + (((dh & 1) << 1) | (dw & 1)) * kPlaneBytes;
          const int rbase = ((dh >> 1) + 1) * kHaloPitch + ((dw >> 1) + 1);
          const uint8_t* colp[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {                         // swizzle phase depends only on the column (pitch % 8 == 0)
            const int r = rbase + c;
            colp[c] = plane + r * 128 + (((chunk_swz ^ (uint32_t)(r & 7)) << 4) | word);
          }
#pragma unroll
          for (int px = 0; px < 32; ++px) {
            const float v = *reinterpret_cast<const float*>(colp[px & 7] + (px >> 3) * (kHaloPitch * 128));
            const uint32_t h = __float_as_uint(v) & 0xffffe000u;
            hi[px] = h;
            lo[px] = __float_as_uint(v - __uint_as_float(h));
            if (p.debug == 1) { hi[px] = __float_as_uint(1.0f); lo[px] = 0u; }
          }
        } else {
#pragma unroll
          for (int px = 0; px < 32; ++px) { hi[px] = 0u; lo[px] = 0u; }
        }
        mbar_wait(bar_aempty + 8 * slot, ((u >> 2) & 1) ^ 1);
        tc_fence_after();
        const uint32_t a_slot = lane_base + aoff + slot * 64;
        tmem_st32(a_slot, hi);
        tmem_st32(a_slot + 32, lo);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bar_afull + 8 * slot);
      }
      if (++s == S) { s = 0; ph ^= 1; }
    }
    // ---- epilogue: accumulator tiles -> partial[chunk][(tap, channel)][co]; tiles split between the two groups
    if (nkb > 0) {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
    }
    for (int mt = cg; mt < QT; mt += 2) {
      const int tap = 4 * mt + q;
      for (int c0 = 0; c0 < Co; c0 += 32) {
        uint32_t v[32];
        if (nkb > 0) {
          tmem_ld32(lane_base + mt * Co + c0, v);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        if (tap < ntaps) {
          const size_t row = (size_t)(t0 + tap) * p.Cg + cb * 32 + lane;
          float* dst = p.partial + ((size_t)blockIdx.x * p.Mp + row) * Co + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                              __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

}  // namespace

int uad_tc_gather_supported(int Cin, int N, int lgMH, int lgMW) {
  if (Cin % kKBlk != 0 || Cin < kKBlk) return 0;
  if (!(N == 32 || N == 64 || N == 128)) return 0;
  if (lgMH < 3 || lgMW < 3) return 0;          // M-grid at least 8x8 (two images per tile)
  return 1;
}

size_t uad_tc_gather_ws_bytes(int ksize, int Cin, int N) {
  if (Cin % kKBlk != 0) return 0;
  return (size_t)ksize * ksize * Cin * N * 2 * sizeof(float) + 1024;
}

int uad_launch_gather_tc(const GatherParams& g, int nclasses, int ksize, bool weights_transposed, const float* w_raw,
                         int math_mode, void* ws, size_t ws_bytes, cudaStream_t st) {
  UAD_REQUIRE(math_mode == UAD_MATH_TC_3XTF32, "gather_gemm_tc: only the 3xTF32 mode is implemented (got %d)", math_mode);
  const int C = g.Cin, N = g.N;
  const size_t need = uad_tc_gather_ws_bytes(ksize, C, N);
  UAD_REQUIRE(ws && ws_bytes >= need, "gather_gemm_tc: workspace too small (%zu < %zu)", ws_bytes, need);
  UAD_REQUIRE(((uintptr_t)ws % 128) == 0 && ((uintptr_t)g.in % 16) == 0, "gather_gemm_tc: unaligned buffers");
  EncodeTiledFn encode = get_encode_fn();
  UAD_REQUIRE(encode != nullptr, "gather_gemm_tc: cuTensorMapEncodeTiled entry point unavailable");

  float* img = reinterpret_cast<float*>(ws);
  {
    const size_t total = (size_t)ksize * ksize * C * N;
    weight_image_kernel<<<uad_cdiv(total, 256), 256, 0, st>>>(w_raw, img, ksize * ksize, C, N, weights_transposed ? 1 : 0);
    UAD_LAUNCH_CHECK("weight_image");
  }

  TcParams p;
  memset(&p, 0, sizeof(p));
  const int MW = 1 << g.lgMW, MH = 1 << g.lgMH;
  p.TW = MW < kTileM ? MW : kTileM;
  p.TH = (kTileM / p.TW) < MH ? (kTileM / p.TW) : MH;
  p.TB = kTileM / (p.TW * p.TH);
  p.lgTW = uad_ilog2(p.TW);
  p.lgTH = uad_ilog2(p.TH);
  p.tiles_w = MW / p.TW;
  p.tiles_h = MH / p.TH;
  const int tiles_b = uad_cdiv(g.B, p.TB);
  p.B = g.B; p.C = C; p.Cblks = C / kKBlk; p.N = N;
  p.OH = g.OH; p.OW = g.OW; p.osh = g.osh;
  p.stride2 = (g.sh == 2);
  {
    int max_taps = 0;
    for (int c = 0; c < nclasses; ++c) max_taps = g.taps[c].n > max_taps ? g.taps[c].n : max_taps;
    p.G = (max_taps * C > 640) ? 2 : 1;      // deep reductions: halve the accumulation chain length
    p.nacc = (N <= 64) ? 2 * p.G : p.G + 1;
    p.acc_bufs = (2 * p.nacc * N + 128 <= 512) ? 2 : 1;
    const int acc_cols = p.acc_bufs * p.nacc * N;
    p.tmem_cols = 512;                         // one persistent CTA per SM owns all of TMEM
    p.nslots = (512 - acc_cols) / 64;
    if (p.nslots > 4) p.nslots = 4;
    UAD_REQUIRE(p.nslots >= 2, "gather_gemm_tc: TMEM budget exceeded");
  }
  p.nclasses = nclasses;
  p.n_items = p.tiles_w * p.tiles_h * tiles_b * nclasses;
  p.z_out = g.z_out; p.a_out = g.a_out; p.bias = g.bias; p.gamma = g.gamma; p.beta = g.beta;
  p.bn_c = g.bn_c; p.alpha = g.alpha; p.act = g.act;
  p.wimg = img;
  { const char* dbg = getenv("UAD_TC_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
  for (int c = 0; c < 4; ++c) p.taps[c] = g.taps[c];
  UAD_REQUIRE(p.z_out || p.a_out, "gather_gemm_tc: no output requested");

  // 5-D tensor map over the NHWC input: (channel [x parity], W, parity/1, H, B); box = (32 ch, TW, 1, TH, TB)
  CUtensorMap tmap;
  cuuint64_t dims[5], strides[4];
  const cuuint64_t e = sizeof(float);
  if (p.stride2) {
    dims[0] = 2ull * C; dims[1] = g.IW / 2; dims[2] = 2; dims[3] = g.IH / 2; dims[4] = g.B;
    strides[0] = 2ull * C * e; strides[1] = (cuuint64_t)g.IW * C * e; strides[2] = 2ull * g.IW * C * e;
    strides[3] = (cuuint64_t)g.IH * g.IW * C * e;
  } else {
    dims[0] = C; dims[1] = g.IW; dims[2] = 1; dims[3] = g.IH; dims[4] = g.B;
    strides[0] = (cuuint64_t)C * e; strides[1] = (cuuint64_t)g.IW * C * e; strides[2] = (cuuint64_t)g.IW * C * e;
    strides[3] = (cuuint64_t)g.IH * g.IW * C * e;
  }
  cuuint32_t box[5] = {(cuuint32_t)kKBlk, (cuuint32_t)p.TW, 1u, (cuuint32_t)p.TH, (cuuint32_t)p.TB};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(g.in), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UAD_REQUIRE(cr == CUDA_SUCCESS, "gather_gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr);

  const size_t stage_bytes = kABytes + 2u * N * 128u;
  const size_t tail = 256 + 3 * N * sizeof(float) + 4 * 32 * (size_t)(N + 4) * sizeof(float) + 64;
  p.stages = (int)((220 * 1024 - 1024 - tail) / stage_bytes);
  if (p.stages > 8) p.stages = 8;
  UAD_REQUIRE(p.stages >= 2, "gather_gemm_tc: shared-memory budget exceeded");
  const size_t smem = 1024 + p.stages * stage_bytes + tail;
  static bool attr_set = false;
  if (!attr_set) {
    UAD_CUDA(cudaFuncSetAttribute(gather_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    attr_set = true;
  }
  const int grid = p.n_items < UAD_NUM_SMS ? p.n_items : UAD_NUM_SMS;
  gather_gemm_tc<<<grid, 384, smem, st>>>(tmap, p);
  UAD_LAUNCH_CHECK("gather_gemm_tc");
  return 0;
}

// ------------------------------------------------------------------------------------------------ Form W launcher
int uad_tc_wgrad_supported(int Cg, int Co, int lgMH, int lgMW) {
  if (Cg % 32 != 0 || Cg < 32) return 0;
  if (!(Co == 32 || Co == 64 || Co == 128)) return 0;
  if (lgMH < 2 || lgMW < 3) return 0;
  return 1;
}

static void wgrad_tc_plan(int Cg, int Co, int P, int* ngroups, int* tpg, int* QT, int* nchunks, int* bpc) {
  const int qt_max = (512 - 256) / Co;                  // accumulator tiles that fit beside the four A slots
  int qt = qt_max > 7 ? 7 : qt_max;
  *ngroups = uad_cdiv(25, 4 * qt);
  *tpg = uad_cdiv(25, *ngroups);
  *QT = uad_cdiv(*tpg, 4);
  const int nblocks = P / 32;
  int target = (4 * UAD_NUM_SMS) / ((*ngroups) * (Cg / 32));
  if (target < 1) target = 1;
  if (target > nblocks) target = nblocks;
  *bpc = uad_cdiv(nblocks, target);
  *nchunks = uad_cdiv(nblocks, *bpc);
}

size_t uad_tc_wgrad_ws_bytes(int Cg, int Co, int P) {
  int ng, tpg, qt, nch, bpc;
  wgrad_tc_plan(Cg, Co, P, &ng, &tpg, &qt, &nch, &bpc);
  return (size_t)nch * 25 * Cg * Co * sizeof(float) + 1024;
}

int uad_launch_wgrad_tc(const WgradParams& w, float* out, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int Cg = w.Cg, Co = w.Co;
  UAD_REQUIRE(w.sh == 2 && w.taps.n == 25, "wgrad_tc: only the 5x5 stride-2 gather is implemented");
  EncodeTiledFn encode = get_encode_fn();
  UAD_REQUIRE(encode != nullptr, "wgrad_tc: cuTensorMapEncodeTiled entry point unavailable");
  TcWgradParams p;
  memset(&p, 0, sizeof(p));
  int nchunks;
  wgrad_tc_plan(Cg, Co, w.P, &p.ngroups, &p.taps_per_group, &p.QT, &nchunks, &p.blocks_per_chunk);
  const size_t need = (size_t)nchunks * w.Mp * Co * sizeof(float);
  UAD_REQUIRE(ws && ws_bytes >= need, "wgrad_tc: workspace too small (%zu < %zu)", ws_bytes, need);
  p.B = w.B; p.lgMH = w.lgMH; p.lgMW = w.lgMW; p.Cg = Cg; p.Co = Co; p.ncb = Cg / 32;
  p.nblocks = w.P / 32; p.Mp = w.Mp;
  p.partial = reinterpret_cast<float*>(ws);
  for (int t = 0; t < 25; ++t) { p.dh[t] = w.taps.dh[t]; p.dw[t] = w.taps.dw[t]; }
  const size_t stage_bytes = kHaloBytes + 3u * 32u * Co * 4u;
  p.stages = (Co == 128) ? 2 : 3;
  { const char* dbg = getenv("UAD_WGRAD_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }

  const cuuint64_t e = sizeof(float);
  const int MH = 1 << w.lgMH, MW = 1 << w.lgMW;
  CUtensorMap tmap_g, tmap_o;
  {   // gathered fine tensor [B, GH, GW, Cg] viewed as (2*Cg, GW/2, 2, GH/2, B); box = (32 ch, 16, 1, 6, 1)
    cuuint64_t dims[5] = {2ull * Cg, (cuuint64_t)w.GW / 2, 2, (cuuint64_t)w.GH / 2, (cuuint64_t)w.B};
    cuuint64_t strides[4] = {2ull * Cg * e, (cuuint64_t)w.GW * Cg * e, 2ull * w.GW * Cg * e, (cuuint64_t)w.GH * w.GW * Cg * e};
    cuuint32_t box[5] = {32, kHaloPitch, 1, kWPH + 2, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult cr = encode(&tmap_g, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(w.g), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UAD_REQUIRE(cr == CUDA_SUCCESS, "wgrad_tc: cuTensorMapEncodeTiled(g) failed (%d)", (int)cr);
  }
  {   // coarse tensor [B, MH, MW, Co]; box = (32 ch, 8, 4, 1)
    cuuint64_t dims[4] = {(cuuint64_t)Co, (cuuint64_t)MW, (cuuint64_t)MH, (cuuint64_t)w.B};
    cuuint64_t strides[3] = {(cuuint64_t)Co * e, (cuuint64_t)MW * Co * e, (cuuint64_t)MH * MW * Co * e};
    cuuint32_t box[4] = {32, kWPW, kWPH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = encode(&tmap_o, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(w.o), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UAD_REQUIRE(cr == CUDA_SUCCESS, "wgrad_tc: cuTensorMapEncodeTiled(o) failed (%d)", (int)cr);
  }
  const size_t smem = 1024 + p.stages * stage_bytes + 512;
  static bool attr_set = false;
  if (!attr_set) {
    UAD_CUDA(cudaFuncSetAttribute(wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  dim3 grid(nchunks, p.ngroups * p.ncb);
  wgrad_tc<<<grid, 384, smem, st>>>(tmap_g, tmap_o, p);
  UAD_LAUNCH_CHECK("wgrad_tc");
  return uad_launch_splitk_reduce(p.partial, nchunks, (size_t)w.Mp * Co, out, accumulate, st);
}

// developer aid: copy the clock64 trace of the last traced gather_gemm_tc_np launch (UAD_TC_DEBUG bit 16)
extern "C" int uad_debug_trace(long long* out64) {
  UAD_CUDA(cudaMemcpyFromSymbol(out64, g_tc_trace, sizeof(long long) * 64));
  return 0;
}
