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
#include "uad_halo.h"
#include "uad_staging.h"
#include "uad_wgrad_tiles.h"

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
  int n_issuers;       // v2 kernel: MMA-issuing warps (2 for the paired N <= 64 layout, 1 for N = 128)
  int split_n;         // v2 kernel, N = 128: the two issuers split the output columns (64 each) of every k-block
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

// ------------------------------------------------------------------------------------------------ N = 32 variant
// One 128-pixel tile per CTA, 256 threads (the converter warps double as epilogue warps), two CTAs per SM so that the
// converter -> MMA latency chains of two tiles interleave (measured faster than the persistent kernel below for
// N = 32; the persistent kernel wins for N >= 64).  For the transposed form the CTA walks all four output-parity
// classes of its tile back to back: prologue / TMEM allocation are paid once per tile and the TMA producer keeps
// prefetching the next class's operands while the current class is being written out.  Same numerics as below.
// kHalo (round-2 candidate `gather_gemm_tc_np_halo`, opt-in UAD_TC_HALO=1, stride-1 form only, never run on hardware): the
// activation operand is not TMA-loaded per k-block.  The (TH + 2) x (TW + 2) pixel halo of the 8 x 16 pixel tile is loaded ONCE
// per 32-channel block and every tap's converter pass reads its shifted window from it, so a k-block moves only its 2 N x 128 B
// weight image (8 KB instead of 24 KB at N = 32: DESIGN.md 4.1 (6), the L2 -> shared-memory ceiling), and the converters no
// longer wait on the per-k-block TMA barrier.
template <bool kHalo>
__device__ __forceinline__ void gather_gemm_tc_np_body(const CUtensorMap& tmap, const TcParams& p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int N = p.N;
  const uint32_t b_bytes = 2u * N * 128u;
  const uint32_t a_bytes = kHalo ? 0u : (uint32_t)kABytes;   // halo form: the stages hold weight images only
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const int S = p.stages;
  const int NS = p.nslots;
  const uint32_t misc = smem_base + S * stage_bytes;
  const uint32_t bar_full = misc;                       // S x 8
  const uint32_t bar_empty = misc + 64;                 // S x 8
  const uint32_t bar_afull = misc + 128;                // NS x 8
  const uint32_t bar_aempty = misc + 160;               // NS x 8
  const uint32_t bar_acc = misc + 192;
  const uint32_t tmem_slot = misc + 200;
  const uint32_t bar_halo = misc + 208;                 // halo form: all channel blocks of the tile's halo landed
  float* epi = reinterpret_cast<float*>(smem_gen + S * stage_bytes + 256);                 // bias[N], scale[N], shift[N]
  float* stg_base = reinterpret_cast<float*>(smem_gen + S * stage_bytes + 256 + 3 * N * 4);     // 4 x 32 x (N+4) staging

  // halo form: per 32-channel block a (TH + 2) x (TW + 2) pixel tile of 128-byte rows behind the staging rows, 1024-byte aligned
  const int halo_w = p.TW + 2;
  const uint32_t halo_rows = (uint32_t)((p.TH + 2) * halo_w);
  const uint32_t halo_bytes = (halo_rows * 128u + 1023u) & ~1023u;
  const uint32_t halo_off = (S * stage_bytes + 256u + 3u * N * 4u + 4u * 32u * (N + 4) * 4u + 1023u) & ~1023u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t aoff = p.nacc * N;                     // first A slot column
  const int ncls = p.nclasses;

  // tile origin on the M-grid
  const int tile = blockIdx.x;
  const int twi = tile % p.tiles_w;
  const int thi = (tile / p.tiles_w) % p.tiles_h;
  const int tbi = tile / (p.tiles_w * p.tiles_h);
  const int s0 = twi * p.TW, r0 = thi * p.TH, b0 = tbi * p.TB;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < NS; ++i) { mbar_init(bar_afull + 8 * i, 128); mbar_init(bar_aempty + 8 * i, 1); }
    mbar_init(bar_acc, 1);
    if (kHalo) mbar_init(bar_halo, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 3) {
    for (int n = lane; n < N; n += 32) {
      epi[n] = p.bias ? p.bias[n] : 0.f;
      epi[N + n] = p.gamma ? p.gamma[n] * p.bn_c : 1.f;
      epi[2 * N + n] = p.beta ? p.beta[n] : 0.f;
    }
  }
  const bool tracer = (p.debug & 16) && blockIdx.x == gridDim.x / 2 && threadIdx.x == 128;
  int tri = 0;
  if (tracer) g_tc_trace[tri++] = clock64();            // [0] entry
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tracer) g_tc_trace[tri++] = clock64();            // [1] prologue done
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // all ring indices / phase bits are carried incrementally across k-blocks AND classes
  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      int s = 0;
      uint32_t ph = 0;
      if (kHalo) {                                              // the tile's halo, once: box (32 ch, TW + 2, 1, TH + 2, 1)
        mbar_expect_tx(bar_halo, (uint32_t)p.Cblks * halo_rows * 128u);
        for (int cb = 0; cb < p.Cblks; ++cb)
          tma_load_5d(smem_base + halo_off + cb * halo_bytes, &tmap, bar_halo, cb * kKBlk, s0 - 1, 0, r0 - 1, b0);
      }
      for (int cls = 0; cls < ncls; ++cls) {
        const TapSet& ts = p.taps[cls];
        const int nkb = ts.n * p.Cblks;
        int tap = 0, cb = 0;
        for (int i = 0; i < nkb; ++i) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          const uint32_t full = bar_full + 8 * s;
          // developer timing switches: 32 = skip the activation (A) load, 64 = skip the weight (B) load
          mbar_expect_tx(full, ((kHalo || (p.debug & 32)) ? 0u : (uint32_t)kABytes) + ((p.debug & 64) ? 0u : b_bytes));
          const int dh = ts.dh[tap], dw = ts.dw[tap], wt = ts.wt[tap];
          const uint32_t a_dst = smem_base + s * stage_bytes;
          if (kHalo || (p.debug & 32)) {
          } else if (p.stride2)
            tma_load_5d(a_dst, &tmap, full, (dw & 1) * p.C + cb * kKBlk, s0 + (dw >> 1), dh & 1, r0 + (dh >> 1), b0);
          else
            tma_load_5d(a_dst, &tmap, full, cb * kKBlk, s0 + dw, 0, r0 + dh, b0);
          const float* wsrc = p.wimg + ((size_t)(wt * p.Cblks + cb)) * 2 * N * kKBlk;
          if (!(p.debug & 64)) bulk_load(a_dst + a_bytes, wsrc, b_bytes, full);
          if (++cb == p.Cblks) { cb = 0; ++tap; }
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (whole warp converged, one elected lane issues)
    const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTileM >> 4) << 24);
    const uint32_t idescN = idesc_base | ((uint32_t)(N >> 3) << 17);
    const uint32_t idesc2N = idesc_base | ((uint32_t)((2 * N) >> 3) << 17);
    const uint64_t bdesc0 = make_sw128_desc(smem_base + a_bytes);          // B image of stage 0 (hi rows then lo rows)
    const uint32_t stage_units = stage_bytes >> 4;
    int s = 0, t = 0;
    uint32_t ph = 0, pht = 0;
    for (int cls = 0; cls < ncls; ++cls) {
      const int nkb = p.taps[cls].n * p.Cblks;
      int g = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(bar_full + 8 * s, ph);                        // weight image landed (async proxy -> visible)
        mbar_wait(bar_afull + 8 * t, pht);                      // converters filled TMEM A slot t
        tc_fence_after();
        const uint64_t dhi0 = bdesc0 + (uint64_t)(s * stage_units);
        const uint32_t a_hi = tmem_base + aoff + t * 64;
        const uint32_t a_lo = a_hi + 32;
        const uint32_t first = (i >= p.G) ? 1u : 0u;            // accumulator pair g already holds a partial sum of this class?
        if (elect_one()) {
          if (!(p.debug & 2)) {
            const uint32_t d_pair = tmem_base + g * 2 * N;      // [main_g | corr_g]
#pragma unroll
            for (int j = 0; j < 4; ++j) {                       // K = 8 tf32 per instruction -> 32 bytes (2 x 16 B) along the row
              mma_tf32_ts(d_pair, a_hi + j * 8, dhi0 + 2 * j, idesc2N, first | (j != 0));
              mma_tf32_ts(d_pair + N, a_lo + j * 8, dhi0 + 2 * j, idescN, 1u);
            }
          }
          tc_commit(bar_empty + 8 * s);                         // smem stage reusable once these MMAs retire
          tc_commit(bar_aempty + 8 * t);                        // TMEM A slot reusable
          if (i == nkb - 1) tc_commit(bar_acc);                 // accumulators of this class complete
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1; }
        if (++t == NS) { t = 0; pht ^= 1; }
        if (++g == p.G) g = 0;
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== converters, then this class's epilogue
    const int row = threadIdx.x - 128;                          // tile row == TMEM lane
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t swz = (uint32_t)(row & 7);
    const int tw = row & (p.TW - 1);
    const int th = (row >> p.lgTW) & (p.TH - 1);
    const int tb = row >> (p.lgTW + p.lgTH);
    const int b = b0 + tb;
    const int ldw = N + 4;
    float* stg = stg_base + (size_t)q * 32 * ldw;               // this warp's 32 x (N+4) staging rows
    const int lanes_per_row = N / 4;                            // float4 lanes covering one output row
    const int rows_per_it = 32 / lanes_per_row;
    int s = 0, t = 0;
    uint32_t ph = 0, pht = 0;
    if (kHalo) mbar_wait(bar_halo, 0);                          // every window of every tap is read from the resident halo
    for (int cls = 0; cls < ncls; ++cls) {
      const TapSet& ts = p.taps[cls];
      const int nkb = ts.n * p.Cblks;
      int tap = 0, cb = 0;
      for (int i = 0; i < nkb; ++i) {
        if (!kHalo) mbar_wait(bar_full + 8 * s, ph);
        if (p.debug & 1) {
          mbar_wait(bar_aempty + 8 * t, pht ^ 1);
          mbar_arrive(bar_afull + 8 * t);
        } else {
          // halo form: pixel (th + dh, tw + dw) of the tile's halo, i.e. halo row (th + dh + 1) * (TW + 2) + tw + dw + 1
          const int hrow = kHalo ? uad_halo_row(th, tw, ts.dh[tap], ts.dw[tap], p.TW) : 0;
          const uint8_t* arow = kHalo ? smem_gen + halo_off + cb * halo_bytes + hrow * 128 : smem_gen + s * stage_bytes + row * 128;
          const uint32_t swz_k = kHalo ? (uint32_t)(hrow & 7) : swz;
          uint32_t hi[32], lo[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {                         // 16-byte chunk j of a row sits at (j ^ (row index & 7))
            const float4 v = *reinterpret_cast<const float4*>(arow + ((j ^ swz_k) << 4));
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
        }
        if (kHalo) { if (++cb == p.Cblks) { cb = 0; ++tap; } }
        if (++s == S) { s = 0; ph ^= 1; }
        if (++t == NS) { t = 0; pht ^= 1; }
      }

      // ---- epilogue of this class: accumulators -> z, a -> smem transpose -> coalesced rows.  The next class's MMAs
      //      cannot start before these warps convert its first k-block, i.e. after the TMEM reads below completed.
      if (tracer) g_tc_trace[tri++] = clock64();          // conversions of this class issued
      mbar_wait(bar_acc, (uint32_t)(cls & 1));
      tc_fence_after();
      if (tracer) g_tc_trace[tri++] = clock64();          // accumulators complete
      if (p.debug & 8) continue;
      const long long my_off = (b < p.B)
          ? (((long long)b * p.OH + ((r0 + th) * p.osh + ts.oh0)) * p.OW + ((s0 + tw) * p.osh + ts.ow0)) * (long long)N
          : -1;
      uint32_t v[32], u[32];                                    // N == 32 on this path: one 32-column chunk
      tmem_ld32(lane_base, v);
      for (int k = 1; k < p.nacc; ++k) {
        tmem_ld32(lane_base + k * N, u);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
      }
      tmem_wait_ld();
      tc_fence_before();
      if (tracer) g_tc_trace[tri++] = clock64();          // TMEM read done
      // N == 32 here: 8 float4 lanes cover one 128-byte output row, 4 rows per store instruction, 8 instructions per
      // warp.  Row offsets are shuffled once; per pass all 8 smem reads are issued before the 8 global stores (a clock64
      // trace showed the dependent shfl -> LDS -> STG chain per iteration cost ~6000 cycles per class).
      long long offs[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) offs[it] = __shfl_sync(0xffffffffu, my_off, it * 4 + (lane >> 3));
      const int cq = (lane & 7) * 4;
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {                    // z then a from the SAME registers, one staging buffer
        float* out = pass == 0 ? p.z_out : p.a_out;
        if (!out) continue;
        if (pass == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(stg + lane * ldw + j) =
                make_float4(__uint_as_float(v[j]) + epi[j], __uint_as_float(v[j + 1]) + epi[j + 1],
                            __uint_as_float(v[j + 2]) + epi[j + 2], __uint_as_float(v[j + 3]) + epi[j + 3]);
        } else if (p.act == UAD_ACT_LEAKY) {                    // the hot case: branch-free inner loop
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float uu = epi[N + j + e] * (__uint_as_float(v[j + e]) + epi[j + e]) + epi[2 * N + j + e];
              o[e] = uu > 0.f ? uu : p.alpha * uu;
            }
            *reinterpret_cast<float4*>(stg + lane * ldw + j) = make_float4(o[0], o[1], o[2], o[3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              o[e] = uad_act(epi[N + j + e] * (__uint_as_float(v[j + e]) + epi[j + e]) + epi[2 * N + j + e], p.act, p.alpha);
            *reinterpret_cast<float4*>(stg + lane * ldw + j) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        __syncwarp();
        if (!(p.debug & 4)) {
          float4 vals[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) vals[it] = *reinterpret_cast<const float4*>(stg + (it * 4 + (lane >> 3)) * ldw + cq);
#pragma unroll
          for (int it = 0; it < 8; ++it)
            if (offs[it] >= 0) *reinterpret_cast<float4*>(out + offs[it] + cq) = vals[it];
        }
        __syncwarp();
      }
      if (tracer) g_tc_trace[tri++] = clock64();          // stores of this class issued
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tracer) { g_tc_trace[tri++] = clock64(); g_tc_trace[63] = tri; }
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

__global__ void __launch_bounds__(256, 2)
gather_gemm_tc_np(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcParams p) {
  gather_gemm_tc_np_body<false>(tmap, p);
}

__global__ void __launch_bounds__(256, 2)
gather_gemm_tc_np_halo(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcParams p) {
  gather_gemm_tc_np_body<true>(tmap, p);
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
// kSwzStg (round-2 candidate, opt-in UAD_TC_V2 bit 16, not yet run on hardware): the epilogue's per-warp staging rows are
// 32 floats with the 16-byte column group XOR-ed by (row & 7) instead of 36 padded floats - 32 KB instead of 36 KB for the
// eight warps, which is what lets FOUR 48 KB stages fit at N = 128 (an EVEN ring, so the column-split dual issue is legal).
// Both access patterns stay conflict-free: a quarter-warp writes rows r..r+7 at one logical group (8 distinct physical
// groups) and reads one row at 8 logical groups.
template <bool kSwzStg>
__device__ __forceinline__ void gather_gemm_tc2_body(const CUtensorMap& tmap, const TcParams& p) {
  constexpr int kLdStg = kSwzStg ? 32 : 36;                 // == uad_stg_ld<kSwzStg>()
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int N = p.N;
  const uint32_t b_bytes = 2u * N * 128u;
  const uint32_t stage_bytes = kABytes + b_bytes;
  const int S = p.stages;
  const int NS = p.nslots;                              // even: slot parity == k-block parity == converter group
  const uint32_t misc = smem_base + S * stage_bytes;
  const uint32_t bar_full = misc;                       // S x 8   (S <= 8)
  const uint32_t bar_empty = misc + 64;                 // S x 8
  const uint32_t bar_afull = misc + 128;                // NS x 8  (NS <= 6)
  const uint32_t bar_aempty = misc + 176;               // NS x 8
  const uint32_t bar_accfull = misc + 224;
  const uint32_t bar_accempty = misc + 232;
  const uint32_t tmem_slot = misc + 240;
  float* epi = reinterpret_cast<float*>(smem_gen + S * stage_bytes + 256);               // bias[N], scale[N], shift[N]
  float* stg_base = reinterpret_cast<float*>(smem_gen + S * stage_bytes + 256 + 3 * N * 4);   // 8 warps x 32 x 36 staging

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t aoff = p.nacc * N;                     // first A slot column
  const int nclasses = p.nclasses;
  const int n_iss = p.n_issuers;

  if (threadIdx.x == 0) {
    const int n_rel = p.split_n ? 2 : 1;                // column-split mode: BOTH issuers consume every stage / slot
    for (int i = 0; i < S; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, n_rel); }
    for (int i = 0; i < NS; ++i) { mbar_init(bar_afull + 8 * i, 128); mbar_init(bar_aempty + 8 * i, n_rel); }
    mbar_init(bar_accfull, n_iss);
    mbar_init(bar_accempty, 256);
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

  if (warp == 0) {
    // ===================================================================== TMA producer (as gather_gemm_tc)
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
  } else if (warp == 1 || warp == 2) {
    // ===================================================================== MMA issuers (whole warp converged, one elected lane issues)
    const int me = warp - 1;
    if (me < n_iss) {
      const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t idescN = idesc_base | ((uint32_t)(N >> 3) << 17);
      const uint32_t idesc2N = idesc_base | ((uint32_t)((2 * N) >> 3) << 17);
      const uint64_t bdesc0 = make_sw128_desc(smem_base + kABytes);
      const uint32_t stage_units = stage_bytes >> 4, lo_units = (uint32_t)(N * 128) >> 4;
      const bool paired = (N <= 64);
      int s = 0, t = 0;
      uint32_t ph = 0, pht = 0, kc = 0, il = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++il) {
        const int nkb = p.taps[item % nclasses].n * p.Cblks;
        bool have_acc = false;
        int g = 0;                                               // single-issuer mode: round robin over the G main accumulators
        for (int i = 0; i < nkb; ++i, ++kc) {
          const bool mine = (n_iss == 1) || p.split_n || ((int)(kc & 1u) == me);
          if (mine) {
            if (!have_acc) {                                     // the previous item's accumulators have been drained
              mbar_wait(bar_accempty, (il & 1u) ^ 1u);
              have_acc = true;
            }
            mbar_wait(bar_full + 8 * s, ph);                     // weight image landed (async proxy -> visible to this thread's MMAs)
            mbar_wait(bar_afull + 8 * t, pht);                   // converters filled TMEM A slot t
            tc_fence_after();
            const uint64_t dhi0 = bdesc0 + (uint64_t)(s * stage_units);
            const uint32_t a_hi = tmem_base + aoff + t * 64;
            const uint32_t a_lo = a_hi + 32;
            const bool last_mine = (n_iss == 1 || p.split_n) ? (i == nkb - 1) : (i >= nkb - 2);
            if (elect_one()) {
              if (p.debug & 2) {
              } else if (p.split_n) {
                // N = 128, column-split dual issue: this warp owns output columns [64*me, 64*me + 64) of EVERY k-block:
                // B rows 64*me.. of the hi / lo images, accumulators [main_0 | main_1 | corr] of 64 columns at 192*me
                const uint32_t first = (i >= p.G) ? 1u : 0u;
                const uint32_t accb = tmem_base + me * 192;
                const uint32_t d_main = accb + g * 64;
                const uint32_t d_corr = accb + 128;
                const uint64_t dh = dhi0 + (uint64_t)(me * ((64 * 128) >> 4));
                const uint32_t idesc64 = idesc_base | ((uint32_t)(64 >> 3) << 17);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  mma_tf32_ts(d_corr, a_lo + j * 8, dh + 2 * j, idesc64, (i | j) != 0);
                  mma_tf32_ts(d_corr, a_hi + j * 8, dh + lo_units + 2 * j, idesc64, 1u);
                  mma_tf32_ts(d_main, a_hi + j * 8, dh + 2 * j, idesc64, first | (j != 0));
                }
              } else if (paired) {
                const int gi = (n_iss == 2) ? (i & 1) : g;       // dual issue: the pair is owned by the k-block parity
                const uint32_t first = (i >= p.G) ? 1u : 0u;
                const uint32_t d_pair = tmem_base + gi * 2 * N;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  mma_tf32_ts(d_pair, a_hi + j * 8, dhi0 + 2 * j, idesc2N, first | (j != 0));
                  mma_tf32_ts(d_pair + N, a_lo + j * 8, dhi0 + 2 * j, idescN, 1u);
                }
              } else {
                const uint32_t first = (i >= p.G) ? 1u : 0u;
                const uint32_t d_main = tmem_base + g * N;
                const uint32_t d_corr = tmem_base + p.G * N;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  mma_tf32_ts(d_corr, a_lo + j * 8, dhi0 + 2 * j, idescN, (i | j) != 0);
                  mma_tf32_ts(d_corr, a_hi + j * 8, dhi0 + lo_units + 2 * j, idescN, 1u);
                  mma_tf32_ts(d_main, a_hi + j * 8, dhi0 + 2 * j, idescN, first | (j != 0));
                }
              }
              tc_commit(bar_empty + 8 * s);
              tc_commit(bar_aempty + 8 * t);
              if (last_mine) tc_commit(bar_accfull);
            }
            __syncwarp();
          }
          if (++s == S) { s = 0; ph ^= 1; }
          if (++t == NS) { t = 0; pht ^= 1; }
          if (++g == p.G) g = 0;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== converter groups, then the shared epilogue
    const int grp = (warp - 4) >> 2;
    const int row = (threadIdx.x - 128) & 127;                  // tile row == TMEM lane
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t swz = (uint32_t)(row & 7);
    float* stg = stg_base + (size_t)(warp - 4) * 32 * kLdStg;   // this warp's 32 staging rows (padded or swizzled)
    const int nchunks = N >> 5;
    int s = 0, t = 0;
    uint32_t ph = 0, pht = 0, kc = 0, il = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++il) {
      const int cls = item % nclasses, tile = item / nclasses;
      const TapSet& ts = p.taps[cls];
      const int nkb = ts.n * p.Cblks;
      for (int i = 0; i < nkb; ++i, ++kc) {
        if ((int)(kc & 1u) == grp) {
          mbar_wait(bar_full + 8 * s, ph);
          if (p.debug & 1) {
            mbar_wait(bar_aempty + 8 * t, pht ^ 1);
            mbar_arrive(bar_afull + 8 * t);
          } else {
            const uint8_t* arow = smem_gen + s * stage_bytes + row * 128;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {                       // 16-byte chunk j of this row sits at (j ^ (row & 7))
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
          }
        }
        if (++s == S) { s = 0; ph ^= 1; }
        if (++t == NS) { t = 0; pht ^= 1; }
      }
      // ---- epilogue of this item: both groups, alternate 32-column chunks
      const int twi = tile % p.tiles_w, thi = (tile / p.tiles_w) % p.tiles_h, tbi = tile / (p.tiles_w * p.tiles_h);
      const int s0 = twi * p.TW, r0 = thi * p.TH, b0 = tbi * p.TB;
      const int tw = row & (p.TW - 1);
      const int th = (row >> p.lgTW) & (p.TH - 1);
      const int tb = row >> (p.lgTW + p.lgTH);
      const int b = b0 + tb;
      const long long my_off = (b < p.B)
          ? (((long long)b * p.OH + ((r0 + th) * p.osh + ts.oh0)) * p.OW + ((s0 + tw) * p.osh + ts.ow0)) * (long long)N
          : -1;
      mbar_wait(bar_accfull, il & 1u);
      tc_fence_after();
      long long offs[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) offs[it] = __shfl_sync(0xffffffffu, my_off, it * 4 + (lane >> 3));
      const int cq = (lane & 7) * 4;
      const int c_first = (nchunks == 1) ? (grp == 0 ? 0 : nchunks) : grp;
      const int c_step = (nchunks == 1) ? 1 : 2;
      bool released = false;
      for (int c = c_first; c < nchunks; c += c_step) {
        const int c0 = c * 32;
        uint32_t v[32], u[32];
        // accumulator k of output columns c0..c0+31: k * N + c0, or (column-split) 192 * half + 64 * k + (c0 % 64)
        const uint32_t acc_c0 = p.split_n ? (uint32_t)((c0 >> 6) * 192 + (c0 & 63)) : (uint32_t)c0;
        const uint32_t acc_stride = p.split_n ? 64u : (uint32_t)N;
        tmem_ld32(lane_base + acc_c0, v);
        for (int k = 1; k < p.nacc; ++k) {
          tmem_ld32(lane_base + k * acc_stride + acc_c0, u);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
        }
        tmem_wait_ld();
        if (c + c_step >= nchunks) {                            // last TMEM read of this thread for the item
          tc_fence_before();
          mbar_arrive(bar_accempty);
          released = true;
        }
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {                  // z then a from the SAME registers
          float* out = pass == 0 ? p.z_out : p.a_out;
          if (!out) continue;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int n = c0 + j + e;
              const float z = __uint_as_float(v[j + e]) + epi[n];
              o[e] = pass == 0 ? z : uad_act(epi[N + n] * z + epi[2 * N + n], p.act, p.alpha);
            }
            *reinterpret_cast<float4*>(stg + uad_stg_write_index<kSwzStg>(lane, j)) = make_float4(o[0], o[1], o[2], o[3]);
          }
          __syncwarp();
          float4 vals[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) vals[it] = *reinterpret_cast<const float4*>(stg + uad_stg_read_index<kSwzStg>(lane, it));
#pragma unroll
          for (int it = 0; it < 8; ++it)
            if (offs[it] >= 0) *reinterpret_cast<float4*>(out + offs[it] + c0 + cq) = vals[it];
          __syncwarp();
        }
      }
      if (!released) {                                          // this group owns no chunk (N = 32): just hand back
        tc_fence_before();
        mbar_arrive(bar_accempty);
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

__global__ void __launch_bounds__(384, 1)
gather_gemm_tc2(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcParams p) {
  gather_gemm_tc2_body<false>(tmap, p);
}

__global__ void __launch_bounds__(384, 1)
gather_gemm_tc2_swz(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcParams p) {
  gather_gemm_tc2_body<true>(tmap, p);
}

// ------------------------------------------------------------------------------------------------ the kernel, v3 (CANDIDATE)
// Round-2 candidate for the N = 32 layers (1.5 ms of the VAE-256 step), written at the end of round 1 AFTER the GPU budget
// was spent: it compiles, its barrier protocol passes the random-schedule model (tests/test_pipeline_protocol.py), but it
// has NOT run on hardware yet - opt-in only (UAD_TC_V3=1), never selected by default.
// = gather_gemm_tc2 (two converter groups, two issuers on alternate k-blocks, even rings) plus what the N = 32 shapes need:
// items there are short (4-9 k-blocks per output-parity class), so the serial epilogue of v2 would dominate; here a
// DEDICATED epilogue warpgroup (warps 12-15) drains accumulator set b while the issuers already fill set b ^ 1
// (two sets of 4 x 32 columns + four A slots = 512 TMEM columns).
// 16 warps: 0 TMA producer | 1 issuer A | 2 TMEM alloc, issuer B | 3 constants | 4-7, 8-11 converter groups | 12-15 epilogue.
__global__ void __launch_bounds__(512, 1)
gather_gemm_tc3(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int N = p.N;
  const uint32_t b_bytes = 2u * N * 128u;
  const uint32_t stage_bytes = kABytes + b_bytes;
  const int S = p.stages;
  const int NS = p.nslots;                              // even: slot parity == k-block parity == converter group
  const uint32_t misc = smem_base + S * stage_bytes;
  const uint32_t bar_full = misc;                       // S x 8   (S <= 8)
  const uint32_t bar_empty = misc + 64;                 // S x 8
  const uint32_t bar_afull = misc + 128;                // NS x 8  (NS <= 6)
  const uint32_t bar_aempty = misc + 176;               // NS x 8
  const uint32_t bar_accfull = misc + 224;              // 2 x 8 (one per accumulator set)
  const uint32_t bar_accempty = misc + 240;             // 2 x 8
  const uint32_t tmem_slot = misc + 256;
  float* epi = reinterpret_cast<float*>(smem_gen + S * stage_bytes + 320);               // bias[N], scale[N], shift[N]
  float* stg_base = reinterpret_cast<float*>(smem_gen + S * stage_bytes + 320 + 3 * N * 4);   // 4 warps x 32 x 36 staging

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t acc_cols = p.nacc * N;                 // columns of one accumulator set
  const uint32_t aoff = 2 * acc_cols;                   // first A slot column (after the two accumulator sets)
  const int nclasses = p.nclasses;
  const int n_iss = p.n_issuers;

  if (threadIdx.x == 0) {
    const int n_rel = p.split_n ? 2 : 1;                // column-split mode: BOTH issuers consume every stage / slot
    for (int i = 0; i < S; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, n_rel); }
    for (int i = 0; i < NS; ++i) { mbar_init(bar_afull + 8 * i, 128); mbar_init(bar_aempty + 8 * i, n_rel); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_accfull + 8 * i, n_iss); mbar_init(bar_accempty + 8 * i, 128); }
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

  if (warp == 0) {
    // ===================================================================== TMA producer (as gather_gemm_tc)
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
  } else if (warp == 1 || warp == 2) {
    // ===================================================================== MMA issuers (whole warp converged, one elected lane issues)
    const int me = warp - 1;
    if (me < n_iss) {
      const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t idescN = idesc_base | ((uint32_t)(N >> 3) << 17);
      const uint32_t idesc2N = idesc_base | ((uint32_t)((2 * N) >> 3) << 17);
      const uint64_t bdesc0 = make_sw128_desc(smem_base + kABytes);
      const uint32_t stage_units = stage_bytes >> 4, lo_units = (uint32_t)(N * 128) >> 4;
      const bool paired = (N <= 64);
      int s = 0, t = 0;
      uint32_t ph = 0, pht = 0, kc = 0, il = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++il) {
        const int nkb = p.taps[item % nclasses].n * p.Cblks;
        bool have_acc = false;
        int g = 0;                                               // single-issuer mode: round robin over the G main accumulators
        for (int i = 0; i < nkb; ++i, ++kc) {
          const bool mine = (n_iss == 1) || p.split_n || ((int)(kc & 1u) == me);
          if (mine) {
            const uint32_t buf = il & 1u, use = il >> 1;         // accumulator set and how often it has been used
            if (!have_acc) {                                     // the epilogue group has drained this set's previous item
              mbar_wait(bar_accempty + 8 * buf, (use & 1u) ^ 1u);
              have_acc = true;
            }
            const uint32_t acc0 = tmem_base + buf * acc_cols;
            mbar_wait(bar_full + 8 * s, ph);                     // weight image landed (async proxy -> visible to this thread's MMAs)
            mbar_wait(bar_afull + 8 * t, pht);                   // converters filled TMEM A slot t
            tc_fence_after();
            const uint64_t dhi0 = bdesc0 + (uint64_t)(s * stage_units);
            const uint32_t a_hi = tmem_base + aoff + t * 64;
            const uint32_t a_lo = a_hi + 32;
            const bool last_mine = (n_iss == 1 || p.split_n) ? (i == nkb - 1) : (i >= nkb - 2);
            if (elect_one()) {
              if (p.debug & 2) {
              } else if (p.split_n) {
                // N = 128, column-split dual issue: this warp owns output columns [64*me, 64*me + 64) of EVERY k-block:
                // B rows 64*me.. of the hi / lo images, accumulators [main_0 | main_1 | corr] of 64 columns at 192*me
                const uint32_t first = (i >= p.G) ? 1u : 0u;
                const uint32_t accb = acc0 + me * 192;
                const uint32_t d_main = accb + g * 64;
                const uint32_t d_corr = accb + 128;
                const uint64_t dh = dhi0 + (uint64_t)(me * ((64 * 128) >> 4));
                const uint32_t idesc64 = idesc_base | ((uint32_t)(64 >> 3) << 17);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  mma_tf32_ts(d_corr, a_lo + j * 8, dh + 2 * j, idesc64, (i | j) != 0);
                  mma_tf32_ts(d_corr, a_hi + j * 8, dh + lo_units + 2 * j, idesc64, 1u);
                  mma_tf32_ts(d_main, a_hi + j * 8, dh + 2 * j, idesc64, first | (j != 0));
                }
              } else if (paired) {
                const int gi = (n_iss == 2) ? (i & 1) : g;       // dual issue: the pair is owned by the k-block parity
                const uint32_t first = (i >= p.G) ? 1u : 0u;
                const uint32_t d_pair = acc0 + gi * 2 * N;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  mma_tf32_ts(d_pair, a_hi + j * 8, dhi0 + 2 * j, idesc2N, first | (j != 0));
                  mma_tf32_ts(d_pair + N, a_lo + j * 8, dhi0 + 2 * j, idescN, 1u);
                }
              } else {
                const uint32_t first = (i >= p.G) ? 1u : 0u;
                const uint32_t d_main = acc0 + g * N;
                const uint32_t d_corr = acc0 + p.G * N;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  mma_tf32_ts(d_corr, a_lo + j * 8, dhi0 + 2 * j, idescN, (i | j) != 0);
                  mma_tf32_ts(d_corr, a_hi + j * 8, dhi0 + lo_units + 2 * j, idescN, 1u);
                  mma_tf32_ts(d_main, a_hi + j * 8, dhi0 + 2 * j, idescN, first | (j != 0));
                }
              }
              tc_commit(bar_empty + 8 * s);
              tc_commit(bar_aempty + 8 * t);
              if (last_mine) tc_commit(bar_accfull + 8 * buf);
            }
            __syncwarp();
          }
          if (++s == S) { s = 0; ph ^= 1; }
          if (++t == NS) { t = 0; pht ^= 1; }
          if (++g == p.G) g = 0;
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================================================================== converter groups (conversion only)
    const int grp = (warp - 4) >> 2;
    const int row = (threadIdx.x - 128) & 127;                  // tile row == TMEM lane
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t swz = (uint32_t)(row & 7);
    int s = 0, t = 0;
    uint32_t ph = 0, pht = 0, kc = 0, il = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++il) {
      const int nkb = p.taps[item % nclasses].n * p.Cblks;
      for (int i = 0; i < nkb; ++i, ++kc) {
        if ((int)(kc & 1u) == grp) {
          mbar_wait(bar_full + 8 * s, ph);
          if (p.debug & 1) {
            mbar_wait(bar_aempty + 8 * t, pht ^ 1);
            mbar_arrive(bar_afull + 8 * t);
          } else {
            const uint8_t* arow = smem_gen + s * stage_bytes + row * 128;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {                       // 16-byte chunk j of this row sits at (j ^ (row & 7))
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
          }
        }
        if (++s == S) { s = 0; ph ^= 1; }
        if (++t == NS) { t = 0; pht ^= 1; }
      }
    }
  } else if (warp >= 12) {
    // ===================================================================== dedicated epilogue group
    const int row = threadIdx.x - 384;                          // tile row == TMEM lane
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float* stg = stg_base + (size_t)q * 32 * 36;                // this warp's 32 x (32+4) staging rows
    const int nchunks = N >> 5;
    uint32_t il = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++il) {
      const int cls = item % nclasses, tile = item / nclasses;
      const TapSet& ts = p.taps[cls];
      const uint32_t buf = il & 1u, use = il >> 1;
      const uint32_t acc_base = lane_base + buf * acc_cols;
      const int twi = tile % p.tiles_w, thi = (tile / p.tiles_w) % p.tiles_h, tbi = tile / (p.tiles_w * p.tiles_h);
      const int s0 = twi * p.TW, r0 = thi * p.TH, b0 = tbi * p.TB;
      const int tw = row & (p.TW - 1);
      const int th = (row >> p.lgTW) & (p.TH - 1);
      const int tb = row >> (p.lgTW + p.lgTH);
      const int b = b0 + tb;
      const long long my_off = (b < p.B)
          ? (((long long)b * p.OH + ((r0 + th) * p.osh + ts.oh0)) * p.OW + ((s0 + tw) * p.osh + ts.ow0)) * (long long)N
          : -1;
      mbar_wait(bar_accfull + 8 * buf, use & 1u);
      tc_fence_after();
      long long offs[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) offs[it] = __shfl_sync(0xffffffffu, my_off, it * 4 + (lane >> 3));
      const int cq = (lane & 7) * 4;
      for (int c = 0; c < nchunks; ++c) {
        const int c0 = c * 32;
        uint32_t v[32], u[32];
        // accumulator k of output columns c0..c0+31: k * N + c0, or (column-split) 192 * half + 64 * k + (c0 % 64)
        const uint32_t acc_c0 = p.split_n ? (uint32_t)((c0 >> 6) * 192 + (c0 & 63)) : (uint32_t)c0;
        const uint32_t acc_stride = p.split_n ? 64u : (uint32_t)N;
        tmem_ld32(acc_base + acc_c0, v);
        for (int k = 1; k < p.nacc; ++k) {
          tmem_ld32(acc_base + k * acc_stride + acc_c0, u);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
        }
        tmem_wait_ld();
        if (c + 1 == nchunks) {                                 // last TMEM read of this thread for the item: hand the set back
          tc_fence_before();
          mbar_arrive(bar_accempty + 8 * buf);
        }
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {                  // z then a from the SAME registers
          float* out = pass == 0 ? p.z_out : p.a_out;
          if (!out) continue;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int n = c0 + j + e;
              const float z = __uint_as_float(v[j + e]) + epi[n];
              o[e] = pass == 0 ? z : uad_act(epi[N + n] * z + epi[2 * N + n], p.act, p.alpha);
            }
            *reinterpret_cast<float4*>(stg + lane * 36 + j) = make_float4(o[0], o[1], o[2], o[3]);
          }
          __syncwarp();
          float4 vals[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) vals[it] = *reinterpret_cast<const float4*>(stg + (it * 4 + (lane >> 3)) * 36 + cq);
#pragma unroll
          for (int it = 0; it < 8; ++it)
            if (offs[it] >= 0) *reinterpret_cast<float4*>(out + offs[it] + c0 + cq) = vals[it];
          __syncwarp();
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


// ------------------------------------------------------------------------------------------------ the kernel, SS form (CANDIDATE)
// Round-2 candidate written after round 1's GPU budget was spent: compiled, NEVER RUN on hardware, opt-in only (UAD_TC_SS bit
// mask: 1 = N = 128 layers, 2 = N = 64, 4 = N = 32).  Idea: at N = 128 a tf32 MMA costs 64 cycles whether A comes from tensor
// memory or from shared memory (profiles/r1_ubench_mma_rate.txt), so the whole converter apparatus (smem -> registers -> split ->
// tcgen05.st -> barrier round trips, the measured reason the tensor pipe is 25-50 % busy) buys nothing there.  Here the
// activation tensor is split ONCE per call into tf32 {hi, lo} images in the workspace (split_hilo_kernel, HBM-bound: 12 bytes
// per element), TMA loads the hi and the lo tile of a k-block straight into the K-major SWIZZLE_128B form the descriptors read,
// and the issuer's only dependency is the TMA barrier: the classic two-role TMA -> MMA ring.
// 8 warps: 0 TMA producer | 1 MMA issuer | 2 TMEM alloc | 3 epilogue constants | 4-7 epilogue.
__global__ void split_hilo_kernel(const float4* __restrict__ x, float4* __restrict__ hi, float4* __restrict__ lo, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
    if (hi) hi[i] = h;                                   // hi == nullptr: UAD_TC_SS bit 8 (the raw tensor serves as the hi operand)
    lo[i] = l;
  }
}

// D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(256, 1)
gather_gemm_ss(const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo,
               const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int N = p.N;
  const uint32_t b_bytes = 2u * N * 128u;
  const uint32_t stage_bytes = 2u * kABytes + b_bytes;  // [A_hi tile | A_lo tile | B_hi rows, B_lo rows]
  const int S = p.stages;
  const uint32_t misc = smem_base + S * stage_bytes;
  const uint32_t bar_full = misc;                       // S x 8   (S <= 8)
  const uint32_t bar_empty = misc + 64;                 // S x 8
  const uint32_t bar_accfull = misc + 128;              // 2 x 8
  const uint32_t bar_accempty = misc + 144;             // 2 x 8
  const uint32_t tmem_slot = misc + 160;
  float* epi = reinterpret_cast<float*>(smem_gen + S * stage_bytes + 320);               // bias[N], scale[N], shift[N]
  float* stg_base = reinterpret_cast<float*>(smem_gen + S * stage_bytes + 320 + 3 * N * 4);   // 4 warps x 32 x 36 staging

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t acc_cols = p.nacc * N;                 // columns of one accumulator set
  const int nclasses = p.nclasses;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
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

  if (warp == 0) {
    // ===================================================================== TMA producer: hi tile, lo tile, weight image
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_hi) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_lo) : "memory");
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
          mbar_expect_tx(full, 2u * (uint32_t)kABytes + b_bytes);
          const int dh = ts.dh[tap], dw = ts.dw[tap], wt = ts.wt[tap];
          const uint32_t a_dst = smem_base + s * stage_bytes;
          if (p.stride2) {
            tma_load_5d(a_dst, &tmap_hi, full, (dw & 1) * p.C + cb * kKBlk, s0 + (dw >> 1), dh & 1, r0 + (dh >> 1), b0);
            tma_load_5d(a_dst + kABytes, &tmap_lo, full, (dw & 1) * p.C + cb * kKBlk, s0 + (dw >> 1), dh & 1, r0 + (dh >> 1), b0);
          } else {
            tma_load_5d(a_dst, &tmap_hi, full, cb * kKBlk, s0 + dw, 0, r0 + dh, b0);
            tma_load_5d(a_dst + kABytes, &tmap_lo, full, cb * kKBlk, s0 + dw, 0, r0 + dh, b0);
          }
          const float* wsrc = p.wimg + ((size_t)(wt * p.Cblks + cb)) * 2 * N * kKBlk;
          bulk_load(a_dst + 2 * kABytes, wsrc, b_bytes, full);
          if (++cb == p.Cblks) { cb = 0; ++tap; }
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (whole warp converged, one elected lane issues)
    const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTileM >> 4) << 24);
    const uint32_t idescN = idesc_base | ((uint32_t)(N >> 3) << 17);
    const uint32_t idesc2N = idesc_base | ((uint32_t)((2 * N) >> 3) << 17);
    const uint64_t adesc0 = make_sw128_desc(smem_base);                      // A_hi tile of stage 0
    const uint32_t stage_units = stage_bytes >> 4, a_units = (uint32_t)kABytes >> 4, lo_units = (uint32_t)(N * 128) >> 4;
    const bool paired = (N <= 64);
    int s = 0, buf = 0;
    uint32_t ph = 0, phb = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int nkb = p.taps[item % nclasses].n * p.Cblks;
      mbar_wait(bar_accempty + 8 * buf, phb ^ 1);               // epilogue has drained this accumulator set
      tc_fence_after();
      const uint32_t acc0 = tmem_base + buf * acc_cols;
      int g = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(bar_full + 8 * s, ph);                        // all three TMA transfers of the stage landed
        tc_fence_after();
        const uint64_t ahi = adesc0 + (uint64_t)(s * stage_units);
        const uint64_t alo = ahi + a_units;
        const uint64_t bhi = ahi + 2 * a_units;                 // B image: hi rows, then lo rows
        const uint32_t first = (i >= p.G) ? 1u : 0u;            // accumulator g already holds a partial sum of this item?
        if (elect_one()) {
          if (paired) {
            const uint32_t d_pair = acc0 + g * 2 * N;
#pragma unroll
            for (int j = 0; j < 4; ++j) {                       // K = 8 tf32 per instruction -> 32 bytes (2 x 16 B) along the row
              mma_tf32_ss(d_pair, ahi + 2 * j, bhi + 2 * j, idesc2N, first | (j != 0));
              mma_tf32_ss(d_pair + N, alo + 2 * j, bhi + 2 * j, idescN, 1u);
            }
          } else {
            const uint32_t d_main = acc0 + g * N;
            const uint32_t d_corr = acc0 + p.G * N;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              mma_tf32_ss(d_corr, alo + 2 * j, bhi + 2 * j, idescN, (i | j) != 0);
              mma_tf32_ss(d_corr, ahi + 2 * j, bhi + lo_units + 2 * j, idescN, 1u);
              mma_tf32_ss(d_main, ahi + 2 * j, bhi + 2 * j, idescN, first | (j != 0));
            }
          }
          tc_commit(bar_empty + 8 * s);                         // smem stage reusable once these MMAs retire
          if (i == nkb - 1) tc_commit(bar_accfull + 8 * buf);   // accumulators of this item complete
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1; }
        if (++g == p.G) g = 0;
      }
      if (++buf == p.acc_bufs) { buf = 0; phb ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue group (as gather_gemm_tc3's)
    const int row = threadIdx.x - 128;                          // tile row == TMEM lane
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float* stg = stg_base + (size_t)q * 32 * 36;                // this warp's 32 x (32+4) staging rows
    const int nchunks = N >> 5;
    int buf = 0;
    uint32_t phb = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int cls = item % nclasses, tile = item / nclasses;
      const TapSet& ts = p.taps[cls];
      const uint32_t acc_base = lane_base + buf * acc_cols;
      const int twi = tile % p.tiles_w, thi = (tile / p.tiles_w) % p.tiles_h, tbi = tile / (p.tiles_w * p.tiles_h);
      const int s0 = twi * p.TW, r0 = thi * p.TH, b0 = tbi * p.TB;
      const int tw = row & (p.TW - 1);
      const int th = (row >> p.lgTW) & (p.TH - 1);
      const int tb = row >> (p.lgTW + p.lgTH);
      const int b = b0 + tb;
      const long long my_off = (b < p.B)
          ? (((long long)b * p.OH + ((r0 + th) * p.osh + ts.oh0)) * p.OW + ((s0 + tw) * p.osh + ts.ow0)) * (long long)N
          : -1;
      mbar_wait(bar_accfull + 8 * buf, phb);
      tc_fence_after();
      long long offs[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) offs[it] = __shfl_sync(0xffffffffu, my_off, it * 4 + (lane >> 3));
      const int cq = (lane & 7) * 4;
      for (int c = 0; c < nchunks; ++c) {
        const int c0 = c * 32;
        uint32_t v[32], u[32];
        tmem_ld32(acc_base + c0, v);                            // accumulator k of output columns c0..c0+31: k * N + c0
        for (int k = 1; k < p.nacc; ++k) {
          tmem_ld32(acc_base + k * N + c0, u);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
        }
        tmem_wait_ld();
        if (c + 1 == nchunks) {                                 // last TMEM read of this thread for the item: hand the set back
          tc_fence_before();
          mbar_arrive(bar_accempty + 8 * buf);
        }
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {                  // z then a from the SAME registers
          float* out = pass == 0 ? p.z_out : p.a_out;
          if (!out) continue;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int n = c0 + j + e;
              const float z = __uint_as_float(v[j + e]) + epi[n];
              o[e] = pass == 0 ? z : uad_act(epi[N + n] * z + epi[2 * N + n], p.act, p.alpha);
            }
            *reinterpret_cast<float4*>(stg + lane * 36 + j) = make_float4(o[0], o[1], o[2], o[3]);
          }
          __syncwarp();
          float4 vals[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) vals[it] = *reinterpret_cast<const float4*>(stg + (it * 4 + (lane >> 3)) * 36 + cq);
#pragma unroll
          for (int it = 0; it < 8; ++it)
            if (offs[it] >= 0) *reinterpret_cast<float4*>(out + offs[it] + c0 + cq) = vals[it];
          __syncwarp();
        }
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

// ================================================================================================ Form W, v2 (CANDIDATE)
// Plane-resident A (DESIGN.md 4.2, round-2 redesign; index arithmetic in uad_wgrad_tiles.h).  Written after round 1's GPU
// budget was spent: compiled, index arithmetic host-tested, NEVER RUN on hardware - opt-in only (UAD_WGRAD_V2=1), and it
// presumes that tcgen05.mma accepts an A operand at an arbitrary tensor-memory column (tools/ubench/operand_probe.cu, E7).
//   * per 4 x 8 pixel block TMA loads the 6 x 10 halo of the four stride-2 parity planes (pitch 10, 8 KB per plane) and the O tile
//   * BOTH converter groups work on EVERY block (no skipped barrier phases, so any stage count is legal): group g splits halo
//     pixels 32g .. 32g+31 of its lane's plane/channel into tf32 hi / lo and stores them to tensor memory ONCE (lane = 32 * plane
//     + channel, column = halo pixel; hi at [0,64), lo at [64,128) of the block's A buffer, two buffers), and transposes + splits
//     half of the O tile into the K-major B_hi / B_lo images
//   * the issuer then runs the block's MMAs back to back - for each window tile (oh, ow) owned by the CTA and each pixel row r:
//     a_lo.b_hi + a_hi.b_lo + a_hi.b_hi with A read at column 10 (oh + r) + ow - with ONE barrier round trip per block instead of
//     one per accumulator tile
// 12 warps: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM alloc, 3 idle, 4-7 / 8-11 = converter groups 0 / 1 (then the epilogue).
constexpr int kW2PlaneBytes = 8192;                  // 60 halo rows of 128 B, padded to the 1024-byte swizzle period
constexpr int kW2HaloBytes = 4 * kW2PlaneBytes;
constexpr int kW2HaloTx = 4 * UAD_WT_HALO_H * UAD_WT_HALO_W * 128;

struct TcWgrad2Params {
  int B, lgMH, lgMW;
  int Cg, Co, ncb;
  int ngroups, tiles_per_group;      // window tiles [grp * tiles_per_group, ...) of the 9 per CTA
  int nblocks, blocks_per_chunk, Mp, stages;
  float* partial;                    // [nchunks][Mp][Co]
};

__global__ void __launch_bounds__(384, 1)
wgrad_tc2(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_o,
          const __grid_constant__ TcWgrad2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int Co = p.Co;
  const uint32_t o_bytes = 32u * Co * 4u;                 // raw O tile [32 px][Co]; then K-major B_hi, B_lo [Co][32 px]
  const uint32_t stage_bytes = kW2HaloBytes + 3 * o_bytes;
  const int S = p.stages;
  const uint32_t misc = smem_base + S * stage_bytes;
  const uint32_t bar_full = misc, bar_empty = misc + 64, bar_afull = misc + 192, bar_aempty = misc + 224, bar_acc = misc + 256,
                 tmem_slot = misc + 264;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.y % p.ngroups;
  const int cb = blockIdx.y / p.ngroups;
  const int tile0 = grp * p.tiles_per_group;
  const int ntiles = min(p.tiles_per_group, UAD_WT_TILES - tile0);
  const int blk_begin = blockIdx.x * p.blocks_per_chunk;
  const int blk_end = min(p.nblocks, blk_begin + p.blocks_per_chunk);
  const int nkb = blk_end - blk_begin;
  const int bw = (1 << p.lgMW) / kWPW, bh = (1 << p.lgMH) / kWPH;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_afull + 8 * i, 256); mbar_init(bar_aempty + 8 * i, 1); }
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
  const uint32_t aoff = p.tiles_per_group * Co;         // two 128-column A buffers after the accumulator tiles

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_g) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_o) : "memory");
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        const uint32_t full = bar_full + 8 * s;
        mbar_expect_tx(full, (uint32_t)kW2HaloTx + o_bytes);
        const int blk = blk_begin + i;
        const int bx = blk % bw, by = (blk / bw) % bh, b = blk / (bw * bh);
        const int r0 = by * kWPH, s0 = bx * kWPW;
        const uint32_t st_base = smem_base + s * stage_bytes;
        for (int pl = 0; pl < 4; ++pl)                 // plane (ph, pw) = (pl >> 1, pl & 1)
          tma_load_5d(st_base + pl * kW2PlaneBytes, &tmap_g, full, (pl & 1) * p.Cg + cb * 32, s0 - 1, pl >> 1, r0 - 1, b);
        for (int a = 0; a < Co / 32; ++a)
          tma_load_4d(st_base + kW2HaloBytes + a * 4096, &tmap_o, full, a * 32, s0, r0, b);
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // whole warp converged, one elected lane issues.  D=f32, A=B=tf32, both K-major, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Co >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t bdesc0 = make_sw128_desc(smem_base + kW2HaloBytes + o_bytes);
    const uint32_t stage_units = stage_bytes >> 4, lo_units = o_bytes >> 4;
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < nkb; ++i) {
      const uint32_t buf = (uint32_t)i & 1u;
      mbar_wait(bar_full + 8 * s, ph);
      mbar_wait(bar_afull + 8 * buf, ((uint32_t)i >> 1) & 1u);   // A planes in tensor memory + B images in shared memory
      tc_fence_after();
      const uint64_t dhi0 = bdesc0 + (uint64_t)(s * stage_units);
      const uint32_t a_hi0 = tmem_base + aoff + buf * 128, a_lo0 = a_hi0 + 64;
      for (int tl = 0; tl < ntiles; ++tl) {
        const uint32_t d = tmem_base + tl * Co;
        if (elect_one()) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {                          // 8 pixels per instruction = 32 bytes along the swizzle row
            const uint32_t col = (uint32_t)uad_wt_a_column(tile0 + tl, r);
            mma_tf32_ts(d, a_lo0 + col, dhi0 + 2 * r, idesc, (i | r) != 0);
            mma_tf32_ts(d, a_hi0 + col, dhi0 + lo_units + 2 * r, idesc, 1u);
            mma_tf32_ts(d, a_hi0 + col, dhi0 + 2 * r, idesc, 1u);
          }
        }
        __syncwarp();
      }
      if (elect_one()) {
        tc_commit(bar_empty + 8 * s);                            // stage (raw tiles + B images) reusable
        tc_commit(bar_aempty + 8 * buf);                         // A buffer reusable
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
    if (elect_one()) tc_commit(bar_acc);
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;                                    // lane group = parity plane
    const int cg = (warp - 4) >> 2;                            // converter group 0 / 1
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t chunk_swz = (uint32_t)(lane >> 2), word = (uint32_t)(lane & 3) << 2;
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < nkb; ++i) {
      const uint32_t buf = (uint32_t)i & 1u;
      mbar_wait(bar_full + 8 * s, ph);
      uint8_t* st = smem_gen + s * stage_bytes;
      // ---- this group's half of the O tile: pixel group pg = 2 q + cg (pixels 4 pg .. 4 pg + 3) of every channel
      {
        const uint8_t* raw = st + kW2HaloBytes;
        uint8_t* bhi = st + kW2HaloBytes + o_bytes;
        uint8_t* blo = bhi + o_bytes;
        const int pg = 2 * q + cg;
        for (int a = 0; a < Co / 32; ++a) {
          const int n = a * 32 + lane;
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
      // ---- this group's half of the plane copy: halo pixels 32 cg .. 32 cg + 31 (60 real ones) of (plane q, channel lane)
      uint32_t hi[32], lo[32];
      {
        const uint8_t* plane = st + q * kW2PlaneBytes;
        const uint8_t* colp[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) colp[c] = plane + (((chunk_swz ^ (uint32_t)c) << 4) | word);   // swizzle phase = row & 7
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int px = 32 * cg + j;                          // row of the plane tile (cg is warp-uniform)
          float v = 0.f;
          if (px < UAD_WT_HALO_H * UAD_WT_HALO_W) v = *reinterpret_cast<const float*>(colp[j & 7] + px * 128);
          const uint32_t h = __float_as_uint(v) & 0xffffe000u;
          hi[j] = h;
          lo[j] = __float_as_uint(v - __uint_as_float(h));
        }
      }
      mbar_wait(bar_aempty + 8 * buf, (((uint32_t)i >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t a_buf = lane_base + aoff + buf * 128 + cg * 32;
      tmem_st32(a_buf, hi);
      tmem_st32(a_buf + 64, lo);
      tmem_wait_st();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // B images: generic-proxy stores -> tensor core
      tc_fence_before();
      mbar_arrive(bar_afull + 8 * buf);
      if (++s == S) { s = 0; ph ^= 1; }
    }
    // ---- epilogue: accumulator tiles -> partial[chunk][(tap, channel)][co]; tiles split between the two groups
    if (nkb > 0) {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
    }
    for (int tl = cg; tl < ntiles; tl += 2) {
      const int tap = uad_wt_tap(tile0 + tl, q);
      for (int c0 = 0; c0 < Co; c0 += 32) {
        uint32_t v[32];
        if (nkb > 0) {
          tmem_ld32(lane_base + tl * Co + c0, v);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        if (tap >= 0) {
          const size_t row = (size_t)tap * p.Cg + cb * 32 + lane;
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

// UAD_TC_SS (bit mask, default 0): 1 = N = 128 layers, 2 = N = 64, 4 = N = 32 run the candidate kernel gather_gemm_ss;
// 8 = the RAW fp32 tensor is the hi operand (only the lo image is written) - valid iff kind::tf32 truncates the low 13 mantissa
// bits of its operands (experiment E1 of tools/ubench/operand_probe.cu); the numerics are then those of the explicit split
static int tc_ss_mask() {
  static int use_ss = -1;
  if (use_ss < 0) { const char* ev = getenv("UAD_TC_SS"); use_ss = ev ? atoi(ev) : 0; }
  return use_ss;
}
static bool tc_ss_enabled(int N) {
  const int m = tc_ss_mask();
  return (N == 128 && (m & 1)) || (N == 64 && (m & 2)) || (N == 32 && (m & 4));
}

// extra workspace behind the weight images when the candidate SS kernel is switched on: the {hi, lo} images of the input
size_t uad_tc_gather_ss_extra_bytes(int N, size_t in_elems) {
  if (!tc_ss_enabled(N)) return 0;
  return 2 * ((in_elems * sizeof(float) + 1023) & ~(size_t)1023) + 2048;
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

  if (tc_ss_enabled(N)) {
    // ---- UAD_TC_SS (bit mask, developer switch, default 0): the round-2 CANDIDATE kernel gather_gemm_ss - not yet run on hardware
    const size_t in_elems = (size_t)g.B * g.IH * g.IW * C;
    const size_t img_bytes = (need + 1023) & ~(size_t)1023;
    const size_t in_bytes = (in_elems * sizeof(float) + 1023) & ~(size_t)1023;
    UAD_REQUIRE(ws_bytes >= img_bytes + 2 * in_bytes, "gather_gemm_ss: workspace too small (%zu < %zu)", ws_bytes,
                img_bytes + 2 * in_bytes);
    float* xhi = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + img_bytes);
    float* xlo = reinterpret_cast<float*>(reinterpret_cast<char*>(xhi) + in_bytes);
    const bool raw_hi = (tc_ss_mask() & 8) != 0;
    {
      const size_t n4 = in_elems / 4;                      // C % 32 == 0
      size_t blocks = uad_cdiv(n4, 256);
      if (blocks > (size_t)UAD_NUM_SMS * 16) blocks = (size_t)UAD_NUM_SMS * 16;
      split_hilo_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(g.in),
                                                          raw_hi ? nullptr : reinterpret_cast<float4*>(xhi),
                                                          reinterpret_cast<float4*>(xlo), n4);
      UAD_LAUNCH_CHECK("split_hilo");
    }
    CUtensorMap tmap_hi, tmap_lo;
    CUresult c1 = encode(&tmap_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, raw_hi ? const_cast<float*>(g.in) : xhi, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult c2 = encode(&tmap_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, xlo, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UAD_REQUIRE(c1 == CUDA_SUCCESS && c2 == CUDA_SUCCESS, "gather_gemm_ss: cuTensorMapEncodeTiled failed (%d, %d)", (int)c1, (int)c2);
    p.G = 2;                                               // two main accumulators (or pairs): halves the accumulation chains
    p.nacc = (N <= 64) ? 2 * p.G : p.G + 1;
    p.acc_bufs = (2 * p.nacc * N <= 512) ? 2 : 1;          // N <= 64: the epilogue overlaps the next item's MMAs
    const size_t stage_ss = 2u * kABytes + 2u * N * 128u;
    const size_t tail_ss = 320 + 3 * N * sizeof(float) + 4 * 32 * 36 * sizeof(float) + 64;
    p.stages = (int)((226 * 1024 - 1024 - tail_ss) / stage_ss);
    if (p.stages > 8) p.stages = 8;
    UAD_REQUIRE(p.stages >= 2, "gather_gemm_ss: shared-memory budget exceeded");
    const size_t smem_ss = 1024 + p.stages * stage_ss + tail_ss;
    static bool attr_ss = false;
    if (!attr_ss) {
      UAD_CUDA(cudaFuncSetAttribute(gather_gemm_ss, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_ss = true;
    }
    const int grid_ss = p.n_items < UAD_NUM_SMS ? p.n_items : UAD_NUM_SMS;
    gather_gemm_ss<<<grid_ss, 256, smem_ss, st>>>(tmap_hi, tmap_lo, p);
    UAD_LAUNCH_CHECK("gather_gemm_ss");
    return 0;
  }

  const size_t stage_bytes = kABytes + 2u * N * 128u;
  {
    // ---- v2 role structure (two converter groups, two issuers for N <= 64); UAD_TC_V2=0 selects the first-generation kernel
    // UAD_TC_V2 (bit mask, developer switch; default 1): 1 = N = 64 layers (dual issue on alternate k-blocks), 4 = N = 128
    // layers (column-split dual issue, measured 13 % faster; +8: single issuer - no faster than the first generation),
    // 2 = N = 32 single-class layers (measured SLOWER than the two-CTA-per-SM N = 32 kernel); 16 = swizzled epilogue staging
    // (candidate, see below); 0 = never.
    // Why N = 128 is NOT on by default: a role that handles every other k-block must own its stages statically, i.e. the
    // stage ring must be EVEN (as the slot ring is) - with an odd ring successive uses of full[s] alternate between the
    // two groups, each group waits with the parity of the use BEFORE the one it skipped and can pass while the skipped
    // load is still in flight (tests/test_pipeline_protocol.py reproduces it).  N = 128 stages are 48 KB: 3 fit, 4 do not
    // (yet), 2 would starve the pipe - so those layers stay on the first-generation kernel until the staging buffer shrinks.
    static int use_v2 = -1;
    if (use_v2 < 0) { const char* e = getenv("UAD_TC_V2"); use_v2 = e ? atoi(e) : 1; }
    if ((N == 64 && (use_v2 & 1)) || (N == 32 && nclasses == 1 && (use_v2 & 2)) || (N == 128 && (use_v2 & 4))) {
      p.split_n = (N == 128 && !(use_v2 & 8)) ? 1 : 0;
      p.n_issuers = (N <= 64 || p.split_n) ? 2 : 1;
      if (p.n_issuers == 2) p.G = 2;                       // paired: each issuer owns one accumulator pair; split: 2 mains per half
      p.nacc = (N <= 64) ? 2 * p.G : p.G + 1;              // accumulators summed per output column by the epilogue
      p.acc_bufs = 1;
      p.nslots = ((512 - (p.split_n ? 384 : p.nacc * N)) / 64) & ~1;
      if (p.nslots > 6) p.nslots = 6;
      UAD_REQUIRE(p.nslots >= 2, "gather_gemm_tc2: TMEM budget exceeded");
      // bit 16 (round-2 candidate, never run on hardware): swizzled 32 x 32 staging -> at N = 128 four 48 KB stages fit in
      // the 227 KB a block may own (1024 alignment slack + 196608 + 256 + 1536 + 32768 + 64 = 232256 <= 232448), an EVEN
      // ring, so UAD_TC_V2=21 (1 + 4 + 16) runs the column-split dual issue with static stage ownership.
      const bool swz_stg = (use_v2 & 16) != 0;
      const size_t tail2 = 256 + 3 * N * sizeof(float) + 8 * 32 * (swz_stg ? 32 : 36) * sizeof(float) + 64;
      p.stages = (int)(((swz_stg ? 227 : 226) * 1024 - 1024 - tail2) / stage_bytes);
      if (p.stages > 8) p.stages = 8;
      p.stages &= ~1;                                      // EVEN ring: static stage ownership per converter group / issuer
      UAD_REQUIRE(p.stages >= 2, "gather_gemm_tc2: shared-memory budget exceeded");
      const size_t smem2 = 1024 + p.stages * stage_bytes + tail2;
      static bool attr2 = false;
      if (!attr2) {
        UAD_CUDA(cudaFuncSetAttribute(gather_gemm_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        UAD_CUDA(cudaFuncSetAttribute(gather_gemm_tc2_swz, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr2 = true;
      }
      const int grid2 = p.n_items < UAD_NUM_SMS ? p.n_items : UAD_NUM_SMS;
      if (swz_stg) gather_gemm_tc2_swz<<<grid2, 384, smem2, st>>>(tmap, p);
      else gather_gemm_tc2<<<grid2, 384, smem2, st>>>(tmap, p);
      UAD_LAUNCH_CHECK("gather_gemm_tc2");
      return 0;
    }
  }
  if (N == 32) {
    {
      // ---- UAD_TC_V3=1 (developer switch, default 0): the round-2 CANDIDATE kernel gather_gemm_tc3 - not yet run on hardware
      static int use_v3 = -1;
      if (use_v3 < 0) { const char* e = getenv("UAD_TC_V3"); use_v3 = e ? atoi(e) : 0; }
      if (use_v3) {
        p.split_n = 0;
        p.n_issuers = 2;
        p.G = 2;                                           // each issuer owns one accumulator pair
        p.nacc = 4;
        p.acc_bufs = 2;
        p.nslots = ((512 - 2 * p.nacc * N) / 64) & ~1;     // 4
        const size_t tail3 = 320 + 3 * N * sizeof(float) + 4 * 32 * 36 * sizeof(float) + 64;
        p.stages = (int)((226 * 1024 - 1024 - tail3) / stage_bytes);
        if (p.stages > 8) p.stages = 8;
        p.stages &= ~1;                                    // even ring: static stage ownership (see gather_gemm_tc2)
        UAD_REQUIRE(p.nslots >= 2 && p.stages >= 2, "gather_gemm_tc3: TMEM / shared-memory budget exceeded");
        const size_t smem3 = 1024 + p.stages * stage_bytes + tail3;
        static bool attr3 = false;
        if (!attr3) {
          UAD_CUDA(cudaFuncSetAttribute(gather_gemm_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          attr3 = true;
        }
        const int grid3 = p.n_items < UAD_NUM_SMS ? p.n_items : UAD_NUM_SMS;
        gather_gemm_tc3<<<grid3, 512, smem3, st>>>(tmap, p);
        UAD_LAUNCH_CHECK("gather_gemm_tc3");
        return 0;
      }
    }
    // ---- N = 32 variant: one tile (all classes) per CTA, TMEM / smem sized for 2 CTAs per SM where possible
    p.acc_bufs = 1;
    const int acc_cols = p.nacc * N;
    p.tmem_cols = (acc_cols + 128 <= 256) ? 256 : 512;
    p.nslots = (p.tmem_cols - acc_cols) / 64;
    if (p.nslots > 4) p.nslots = 4;
    {
      // ---- UAD_TC_HALO=1 (developer switch, default 0): the round-2 CANDIDATE gather_gemm_tc_np_halo - not yet run on hardware.
      // Stride-1 form only (convT fwd / conv dgrad): 8 x 16 pixel tiles of ONE image, the 10 x 18 halo resident per channel block.
      static int use_halo = -1;
      if (use_halo < 0) { const char* ev = getenv("UAD_TC_HALO"); use_halo = ev ? atoi(ev) : 0; }
      if (use_halo && !p.stride2 && MW >= 16 && MH >= 8) {
        p.TW = 16; p.TH = 8; p.TB = 1; p.lgTW = 4; p.lgTH = 3;
        p.tiles_w = MW / p.TW;
        p.tiles_h = MH / p.TH;
        p.n_items = p.tiles_w * p.tiles_h * g.B * nclasses;
        CUtensorMap tmap_halo;
        cuuint32_t hbox[5] = {(cuuint32_t)kKBlk, (cuuint32_t)(p.TW + 2), 1u, (cuuint32_t)(p.TH + 2), 1u};
        CUresult ch = encode(&tmap_halo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(g.in), dims, strides, hbox, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        UAD_REQUIRE(ch == CUDA_SUCCESS, "gather_gemm_tc_np_halo: cuTensorMapEncodeTiled failed (%d)", (int)ch);
        p.stages = 4;                                      // weight images only: 4 x 2 N x 128 B
        const size_t stage_h = 2u * N * 128u;
        const size_t halo_bytes = (((size_t)(p.TH + 2) * (p.TW + 2) * 128u) + 1023u) & ~(size_t)1023u;
        const size_t halo_off = (p.stages * stage_h + 256 + 3 * N * sizeof(float) + 4 * 32 * (size_t)(N + 4) * sizeof(float) + 1023u) &
                                ~(size_t)1023u;
        const size_t smem_h = 1024 + halo_off + p.Cblks * halo_bytes + 64;
        UAD_REQUIRE(smem_h <= 200 * 1024, "gather_gemm_tc_np_halo: shared-memory budget exceeded (%zu)", smem_h);
        static bool attr_h = false;
        if (!attr_h) {
          UAD_CUDA(cudaFuncSetAttribute(gather_gemm_tc_np_halo, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
          attr_h = true;
        }
        gather_gemm_tc_np_halo<<<p.tiles_w * p.tiles_h * g.B, 256, smem_h, st>>>(tmap_halo, p);
        UAD_LAUNCH_CHECK("gather_gemm_tc_np_halo");
        return 0;
      }
    }
    p.stages = 3;
    const size_t smem_np = 1024 + p.stages * stage_bytes + 256 + 3 * N * sizeof(float) + 4 * 32 * (size_t)(N + 4) * sizeof(float) + 64;
    static bool attr_np = false;
    if (!attr_np) {
      UAD_CUDA(cudaFuncSetAttribute(gather_gemm_tc_np, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_np = true;
    }
    gather_gemm_tc_np<<<p.tiles_w * p.tiles_h * tiles_b, 256, smem_np, st>>>(tmap, p);
    UAD_LAUNCH_CHECK("gather_gemm_tc_np");
    return 0;
  }
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

// plan of the plane-resident candidate (wgrad_tc2): 9 window tiles over CTA groups so that tiles * Co + 2 * 128 A columns <= 512
static void wgrad_tc2_plan(int Cg, int Co, int P, int* ngroups, int* tpg, int* nchunks, int* bpc) {
  const int nt_max = 256 / Co;
  *ngroups = uad_cdiv(UAD_WT_TILES, nt_max);
  *tpg = uad_cdiv(UAD_WT_TILES, *ngroups);
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
  {                                                     // the candidate kernel shares the workspace: size it for either plan
    int ng2, tpg2, nch2, bpc2;
    wgrad_tc2_plan(Cg, Co, P, &ng2, &tpg2, &nch2, &bpc2);
    if (nch2 > nch) nch = nch2;
  }
  return (size_t)nch * 25 * Cg * Co * sizeof(float) + 1024;
}

// UAD_WGRAD_V2=1 (developer switch, default 0): the round-2 CANDIDATE kernel wgrad_tc2 - not yet run on hardware
// (window tiles per CTA: 5 at Co = 32, 3 at Co = 64, 2 at Co = 128; stages 4 / 3 / 2)
static int launch_wgrad_tc2(const WgradParams& w, float* out, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st,
                            EncodeTiledFn encode) {
  const int Cg = w.Cg, Co = w.Co;
  for (int t = 0; t < 25; ++t)
    UAD_REQUIRE(w.taps.dh[t] == t / 5 - 1 && w.taps.dw[t] == t % 5 - 1, "wgrad_tc2: unexpected tap table");
  TcWgrad2Params p;
  memset(&p, 0, sizeof(p));
  int nchunks;
  wgrad_tc2_plan(Cg, Co, w.P, &p.ngroups, &p.tiles_per_group, &nchunks, &p.blocks_per_chunk);
  UAD_REQUIRE(p.tiles_per_group * Co + 256 <= 512, "wgrad_tc2: TMEM budget exceeded");
  const size_t need = (size_t)nchunks * w.Mp * Co * sizeof(float);
  UAD_REQUIRE(ws && ws_bytes >= need, "wgrad_tc2: workspace too small (%zu < %zu)", ws_bytes, need);
  p.B = w.B; p.lgMH = w.lgMH; p.lgMW = w.lgMW; p.Cg = Cg; p.Co = Co; p.ncb = Cg / 32;
  p.nblocks = w.P / 32; p.Mp = w.Mp;
  p.partial = reinterpret_cast<float*>(ws);
  const size_t stage_bytes = kW2HaloBytes + 3u * 32u * Co * 4u;
  p.stages = (int)((220 * 1024 - 1024 - 512) / stage_bytes);
  if (p.stages > 4) p.stages = 4;
  UAD_REQUIRE(p.stages >= 2, "wgrad_tc2: shared-memory budget exceeded");

  const cuuint64_t e = sizeof(float);
  const int MH = 1 << w.lgMH, MW = 1 << w.lgMW;
  CUtensorMap tmap_g, tmap_o;
  {   // gathered fine tensor [B, GH, GW, Cg] viewed as (2*Cg, GW/2, 2, GH/2, B); box = (32 ch, 10, 1, 6, 1): one plane's halo
    cuuint64_t dims[5] = {2ull * Cg, (cuuint64_t)w.GW / 2, 2, (cuuint64_t)w.GH / 2, (cuuint64_t)w.B};
    cuuint64_t strides[4] = {2ull * Cg * e, (cuuint64_t)w.GW * Cg * e, 2ull * w.GW * Cg * e, (cuuint64_t)w.GH * w.GW * Cg * e};
    cuuint32_t box[5] = {32, UAD_WT_HALO_W, 1, UAD_WT_HALO_H, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult cr = encode(&tmap_g, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(w.g), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UAD_REQUIRE(cr == CUDA_SUCCESS, "wgrad_tc2: cuTensorMapEncodeTiled(g) failed (%d)", (int)cr);
  }
  {   // coarse tensor [B, MH, MW, Co]; box = (32 ch, 8, 4, 1)
    cuuint64_t dims[4] = {(cuuint64_t)Co, (cuuint64_t)MW, (cuuint64_t)MH, (cuuint64_t)w.B};
    cuuint64_t strides[3] = {(cuuint64_t)Co * e, (cuuint64_t)MW * Co * e, (cuuint64_t)MH * MW * Co * e};
    cuuint32_t box[4] = {32, kWPW, kWPH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = encode(&tmap_o, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(w.o), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UAD_REQUIRE(cr == CUDA_SUCCESS, "wgrad_tc2: cuTensorMapEncodeTiled(o) failed (%d)", (int)cr);
  }
  const size_t smem = 1024 + p.stages * stage_bytes + 512;
  static bool attr_set = false;
  if (!attr_set) {
    UAD_CUDA(cudaFuncSetAttribute(wgrad_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  dim3 grid(nchunks, p.ngroups * p.ncb);
  wgrad_tc2<<<grid, 384, smem, st>>>(tmap_g, tmap_o, p);
  UAD_LAUNCH_CHECK("wgrad_tc2");
  return uad_launch_splitk_reduce(p.partial, nchunks, (size_t)w.Mp * Co, out, accumulate, st);
}

int uad_launch_wgrad_tc(const WgradParams& w, float* out, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int Cg = w.Cg, Co = w.Co;
  UAD_REQUIRE(w.sh == 2 && w.taps.n == 25, "wgrad_tc: only the 5x5 stride-2 gather is implemented");
  EncodeTiledFn encode = get_encode_fn();
  UAD_REQUIRE(encode != nullptr, "wgrad_tc: cuTensorMapEncodeTiled entry point unavailable");
  {
    static int use_v2 = -1;
    if (use_v2 < 0) { const char* ev = getenv("UAD_WGRAD_V2"); use_v2 = ev ? atoi(ev) : 0; }
    if (use_v2) return launch_wgrad_tc2(w, out, accumulate, ws, ws_bytes, st, encode);
  }
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
