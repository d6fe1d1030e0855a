// Halo-resident SS-form implicit GEMM on tcgen05 for the conv family (Form F and Form T of uad_conv.cuh), sm_100a only.
//
// Built on three hardware answers of tools/ubench/operand_probe.cu (profiles/r2_operand_probe.txt):
//   E1      kind::tf32 TRUNCATES fp32 operand words  -> a TMA-loaded fp32 tile IS the tf32 'hi' operand; only lo = x - trunc(x)
//           has to be produced (one elementwise pass over the tile, same byte offsets)
//   E6/E8   the 128-byte swizzle is a function of the absolute shared-memory address -> a K-major descriptor may start at ANY
//           128-byte row of a TMA-written tile and its 8-row groups may be any number of rows apart (SBO = halo pitch)
// and on two measurements of this kernel's earlier versions (profiles/r2_hs_issue_loop.md):
//   (1) a table-driven issue loop (tap tables in constant memory, ~100 dependent SASS instructions per k-block) is
//       instruction-latency bound - 880 cycles per k-block, tensor pipe 25 % busy.  The tap geometry is therefore COMPILE-TIME:
//       the kernel is a template over the form and every k-block of a work item is unrolled (window offsets, accumulator columns
//       and weight-slot offsets are immediates);
//   (2) even then ONE issuing warp needs ~900 cycles per 16 MMAs (branch resolution ~40 cycles per branch, barrier polls,
//       uniform-register descriptor arithmetic) while the pipe executes them in ~700: TWO issuing warps share every work item,
//       each with its own accumulator columns, its own weight ring and its own weight producer, so one polls / commits while
//       the other's MMAs run.
//
// Tile = 16 x 8 pixels of the M-grid (M = 128 rows; one 8-pixel tile row = one 8-row descriptor group).  Per 32-channel block
// the (16+2) x (8+2) pixel halo of the tile (stride-1 form) or of one stride-2 parity plane (strided form) is TMA-loaded ONCE
// as 180 rows of 128 bytes (SWIZZLE_128B); every filter tap that reads from it is a descriptor START ADDRESS
// (halo + (wh * 10 + ww) * 128 bytes, SBO = 1280).  fp32 parity = 3xTF32 with the paired-B trick:
//     acc[main | corr] (+)= A_raw . [B_hi ; B_lo]^T          one MMA of width 2 NI  (hi*hi -> main, hi*lo -> corr)
//     acc[corr]         +=  A_lo  .  B_hi^T                  one MMA of width NI
// Work split between the issuers (NI = columns per issuer):
//   N = 128           column split: issuer W computes output columns [64 W, 64 W + 64) of EVERY k-block (NI = 64)
//   strided, N <= 64  each parity plane's k-blocks are halved; issuer W accumulates into its own pair, the epilogue adds the two
//   stride-1, N <= 64 by output-parity class: classes {0, 3} / {1, 2} (N = 32, four classes per item), class 0 / 1 of the pair (N = 64)
// Roles (416 threads, 544 in the stride-1 form; one persistent CTA per SM):
//   warp 0      halo TMA producer                                   -> h_full[s]
//   warps 4-7   lo pass: lo tile = raw - trunc(raw)                 -> h_lo[s]            (once per halo, NOT per tap)
//   warps 1, 12 weight producers (one bulk copy per 16 KB chunk)    -> w_full[W][t]
//   warps 2, 3  MMA issuers                                         -> w_empty[W][t], h_empty[s], acc_full[b]
//   warps 8-11 (+ 13-16 in the stride-1 form: one warpgroup per accumulator set, alternate items)
//               epilogue: tcgen05.ld, sum of the accumulators (RN), bias / frozen-BN / activation, smem transpose, coalesced
//               row stores                                          -> acc_empty[b]
// The tensor core adds into its fp32 accumulator with truncation (round 1, measured): no accumulator pair takes more than
// ~1600 products per output (N = 128 strided form: channel blocks alternate between G = 2 pairs per issuer).
// Variants of the same kernel (template flags of Cfg, each explained there):
//   FAST    UAD_MATH_TC_1XTF32 - one MMA per K-step, no lo pass, the lo tile's space = twice the halo stages, outputs stored rounded
//   PAIR    M-grids of exactly 8 x 8: two images per 128-row tile, their halos interleaved row by row by one 5-D TMA box
//   SLICED  N of n_total output columns per item (more, narrower items for the layers with too few tiles: the strided PAIR form)
// The epilogue transposes the RAW accumulators through shared memory first, so that a thread owns four channels of eight pixels
// and keeps bias / BN scale / shift / head weights in registers.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "uad_conv.cuh"
#include "uad_tc_ptx.cuh"

namespace {
using namespace uadptx;

constexpr int kTW = 8, kTH = 16;                             // tile: 16 rows x 8 pixels = 128 GEMM rows
constexpr int kHW = kTW + 2, kHH = kTH + 2;                  // halo: 18 rows x 10 pixels
constexpr uint32_t kHaloRows = kHW * kHH;                    // 180 rows of 128 bytes
constexpr uint32_t kHaloBytes = kHaloRows * 128u;            // 23040
constexpr uint32_t kHaloSlot = (kHaloBytes + 1023u) & ~1023u;   // 23552: raw tile, then the lo tile
constexpr uint32_t kHaloStage = 2u * kHaloSlot;
constexpr uint32_t kSbo = kHW * 128u;                        // 1280: one tile row down = one halo row down
constexpr int kTaps = 25;
constexpr uint32_t kSlot = 16384;                            // weight ring slot (one chunk)

template <int I, int E, class F>
__host__ __device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < E) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, E>(f);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Compile-time tap geometry (5 x 5 filter, stride 2, TF SAME: pad 1 before).  A work item's k-blocks (one filter tap x 32
// channels each) come in GROUPS:
//   FORM 0 (strided gather: conv forward, transposed-conv input gradient): group u = input parity plane (ph, pw) = (u >> 1, u & 1),
//          its own halo unit; tap (kh, kw) with kh - 1 = 2 (wh - 1) + ph reads window (wh, ww) of that plane's halo
//   FORM 1 (stride-1 gather, four output-parity classes: transposed-conv forward, conv input gradient): group = class (p, q);
//          tap kh = p + 3 - 2 wh reads window (wh, ww) of the ONE halo; an item covers CG consecutive classes
// Group g has (2 + g>>1) x (2 + g&1) k-blocks in both forms (4, 6, 6, 9); k-block i of it is window (ih, iw) = (i / nw, i % nw).
// RP = halo pixels between two vertically adjacent window positions: kHW, or 2 kHW in the image-pair layout (see Cfg::PAIR)
struct KbGeom { int a_off16, wt; };
__host__ __device__ constexpr int grp_nkb(int g) { return (2 + (g >> 1)) * (2 + (g & 1)); }
template <int FORM, int RP = kHW>
__host__ __device__ constexpr KbGeom kb_geom(int g, int i) {
  const int p = g >> 1, q = g & 1, nw = 2 + q, ih = i / nw, iw = i % nw;
  if (FORM == 0) {
    const int wh = 1 - p + ih, ww = 1 - q + iw;
    return KbGeom{(wh * RP + ww) * 8, (2 * wh + p - 1) * 5 + (2 * ww + q - 1)};
  }
  return KbGeom{(ih * RP + iw) * 8, (p + 3 - 2 * ih) * 5 + (q + 3 - 2 * iw)};
}

template <int FORM_, int N_, int CG_, int G_, bool FAST_ = false, bool PAIR_ = false, bool SLICED_ = false>
struct Cfg {
  static constexpr int FORM = FORM_, N = N_, CG = CG_, G = G_;
  // SLICED: the kernel computes N of HsParams::n_total output columns per item (HsParams::n_slices items per tile).  A template
  // flag because the unsliced kernels must keep N as a compile-time constant in their epilogues (measured: runtime row strides /
  // constant offsets cost the stride-1 N = 32 kernel 17 %)
  static constexpr bool SLICED = SLICED_;
  // PAIR: M-grids of exactly 8 x 8 pixels (the bottleneck layers).  A 128-row tile is TWO images; their halos (10 x 10 pixels each)
  // are loaded by ONE 5-D TMA box whose dimension order (channel, W, image, H) interleaves the images row by row in shared memory:
  // halo row r of image j lies at (2 r + j) * 1280 bytes, so the 16 eight-pixel groups of an operand descriptor (tile row r of
  // image j = group 2 r + j) are still SBO = 1280 bytes apart, and one window step down is 2 x 10 halo pixels.
  static constexpr bool PAIR = PAIR_;
  static constexpr int RP = PAIR ? 2 * kHW : kHW;            // halo pixels per vertical window step
  static constexpr uint32_t HALO_BYTES = PAIR ? 2u * (kTW + 2) * kHW * 128u : kHaloBytes;   // 25600 | 23040
  static constexpr uint32_t HALO_SLOT = (HALO_BYTES + 1023u) & ~1023u;
  // FAST = UAD_MATH_TC_1XTF32: ONE tf32 MMA per K-step (operands rounded to nearest tf32, fp32 accumulation), no lo images, no
  // correction accumulators - the arithmetic of a bf16 / tf32 training step, NOT the fp32-accurate default.  There is no lo pass at
  // all: the tensor core truncates whatever fp32 word it is given (E1), and every conv epilogue of this mode stores its output
  // ALREADY rounded to nearest tf32, so tensors that come from a conv block are exact operands; the others (gradients written by
  // the elementwise kernels, the first layer's output) are truncated.  The lo tile's space becomes extra halo stages.
  static constexpr bool FAST = FAST_;
  static constexpr bool COLSPLIT = N == 128;
  static constexpr int NI = COLSPLIT ? 64 : N;               // output columns per issuer MMA
  static constexpr int PW = FAST ? NI : 2 * NI;              // accumulator columns of one [main | corr] pair (main only when FAST)
  static constexpr int NVAR = FORM == 0 ? 1 : 4 / CG;        // item = tile * NVAR + variant (class group)
  static constexpr int NGRP = FORM == 0 ? 4 : CG;            // k-block groups per (item, channel block)
  static constexpr int NUNITS = FORM == 0 ? 4 : 1;           // halo loads per (item, channel block)
  static constexpr uint32_t WI_BYTES = (FAST ? 1u : 2u) * NI * 128u;   // one k-block's {hi, lo} (FAST: hi) weight image of ONE issuer
  static constexpr int CH = kSlot / WI_BYTES;                // k-blocks per weight chunk (2 at NI = 32, else 1; twice that when FAST)
  // the strided form turns a halo over every 4 .. 9 k-blocks and a refill (TMA + lo pass) takes ~2500 cycles: three stages there,
  // paid for with one weight slot; the stride-1 form keeps a halo for a whole (item, channel block)
  // (the pair layout's larger halos: one halo stage less in the strided form, one weight slot less in the stride-1 form)
  static constexpr int HS = (FORM == 0 ? (PAIR ? 2 : 3) : 2) * (FAST ? 2 : 1), WS = (FORM == 0 || PAIR) ? 2 : 3;   // halo stages, weight slots per issuer
  static constexpr uint32_t HSTAGE = (FAST ? 1u : 2u) * HALO_SLOT;       // bytes per halo stage (raw [+ lo])
  static constexpr int ACC_COLS = FORM == 0 ? (COLSPLIT ? 2 * G * PW : 2 * PW) : (COLSPLIT ? 2 * PW : CG * PW);
  static constexpr int ACC_BUFS = 2 * ACC_COLS <= 512 ? 2 : 1;
  // the stride-1 form writes up to four classes per item: a second epilogue warpgroup (one per accumulator set, alternate items)
  static constexpr int EPI_WG = (FORM == 1 && ACC_BUFS == 2) ? 2 : 1;
  static constexpr int THREADS = 416 + (EPI_WG - 1) * 128;
  static_assert(ACC_COLS <= 512, "TMEM budget");
  static_assert(G == 1 || (FORM == 0 && COLSPLIT), "second accumulator pair per issuer: strided form at N = 128 only");

  // k-blocks [i0, i1) of group g (absolute plane / class index) that issuer W processes
  __host__ __device__ static constexpr int i0(int W, int g) {
    if (COLSPLIT) return 0;
    if (FORM == 0) return W == 0 ? 0 : (grp_nkb(g) + 1) / 2;
    return 0;
  }
  __host__ __device__ static constexpr int i1(int W, int g) {
    if (COLSPLIT) return grp_nkb(g);
    if (FORM == 0) return W == 0 ? (grp_nkb(g) + 1) / 2 : grp_nkb(g);
    const int owner = CG == 4 ? ((g == 0 || g == 3) ? 0 : 1) : (g & 1);
    return owner == W ? grp_nkb(g) : 0;
  }
  // position of issuer W's first k-block of group g in its weight-image sequence of one channel block
  __host__ __device__ static constexpr int ord_base(int W, int g) {
    int s = 0;
    for (int gg = 0; gg < g; ++gg) s += i1(W, gg) - i0(W, gg);
    return s;
  }
  __host__ __device__ static constexpr int cnt(int W) { return ord_base(W, 4); }
  // first / last group (index within the variant) in which issuer W has work - brackets its use of a stride-1 halo unit
  __host__ __device__ static constexpr int first_gi(int W, int V) {
    for (int gi = 0; gi < NGRP; ++gi) { const int g = FORM == 0 ? gi : V * CG + gi; if (i1(W, g) > i0(W, g)) return gi; }
    return -1;
  }
  __host__ __device__ static constexpr int last_gi(int W, int V) {
    int r = -1;
    for (int gi = 0; gi < NGRP; ++gi) { const int g = FORM == 0 ? gi : V * CG + gi; if (i1(W, g) > i0(W, g)) r = gi; }
    return r;
  }
};

struct HsOrder { unsigned char wt[2][kTaps]; int cnt[2]; };  // weight tap of the o-th k-block image of issuer W

struct HsParams {
  int tiles_w, tiles_h;
  int B, C, Cblks;
  int OH, OW, osh;
  int n_items;
  // column slices: the kernel computes N (template) output columns of an n_total-wide layer per item, item = (tile x variant) x
  // n_slices + slice - more, narrower items where a layer has too few tiles to occupy the SMs (the 8 x 8 M-grids); 1 otherwise
  int n_total, n_slices;
  int debug;           // developer timing switches (UAD_HS_DEBUG): 1 = no lo pass, 2 = no MMAs, 4 = no global stores, 8 = no weight loads, 16 = no halo loads, 32 = no epilogue at all, 64 = epilogue reads the accumulators only
  float* z_out;
  float* a_out;
  const float* bias;
  const float* gamma;
  const float* beta;
  float bn_c, alpha;
  int act;
  const float* wimg;   // [Cblks][issuer][k-blocks in the issuer's order][2][NI][32] pre-swizzled {hi, lo} weight images
  // optional fused 1x1 head (stride-1 form, N = 32 only): head_out[pixel] = sum_n a[pixel][n] * head_w[n] + head_b[0]
  const float* head_w;
  const float* head_b;
  float* head_out;
};

struct Ring {                                                // ring position + phase parity
  uint32_t i, ph;
  __device__ __forceinline__ void next(uint32_t n) { if (++i == n) { i = 0; ph ^= 1; } }
};

__device__ __noinline__ float act_slow(float t, int act, float alpha) { return uad_act(t, act, alpha); }
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
// one branch on the fast path; the bounded spin (trap instead of hanging the GPU) lives out of line
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity) {
  if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);
}

struct Bars {
  uint32_t hfull, hlo, hempty, wfull, wempty, accfull, accempty;   // wfull / wempty: [issuer][slot], 8 bytes each, 4 per issuer
};

// ===================================================================== weight producer of issuer W: one bulk copy per chunk
template <class CF, int W>
__device__ __forceinline__ void weight_producer(const HsParams& p, const Bars& bars, uint32_t w_base) {
  constexpr int NVAR = CF::NVAR, NGRP = CF::NGRP, CH = CF::CH, FORM = CF::FORM, CG = CF::CG;
  constexpr uint32_t WI = CF::WI_BYTES;
  const uint32_t full = bars.wfull + W * 32, empty = bars.wempty + W * 32, ring = w_base + W * CF::WS * kSlot;
  const size_t cb_floats = (size_t)kTaps * (CF::FAST ? 1 : 2) * CF::N * 32;
  const float* img0 = p.wimg + (W ? (size_t)CF::cnt(0) * (WI / 4) : 0);
  Ring ws{0, 0};
  const bool no_load = (p.debug & 8) != 0;
  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
    const int nsl = CF::SLICED ? p.n_slices : 1;
    const int slice = CF::SLICED ? item % nsl : 0, var = NVAR > 1 ? (item / nsl) % NVAR : 0;
    for (int cb = 0; cb < p.Cblks; ++cb) {
      const float* cb_img = img0 + ((size_t)slice * p.Cblks + cb) * cb_floats;
      static_for<0, NVAR>([&](auto VI) {
        constexpr int V = decltype(VI)::value;
        if (NVAR > 1 && var != V) return;
        static_for<0, NGRP>([&](auto GI) {
          constexpr int g = FORM == 0 ? decltype(GI)::value : V * CG + decltype(GI)::value;
          constexpr int i0 = CF::i0(W, g), nown = CF::i1(W, g) - i0, ob = CF::ord_base(W, g);
          static_for<0, (nown + CH - 1) / CH>([&](auto CI) {
            constexpr int c0 = decltype(CI)::value * CH;
            constexpr int nk = nown - c0 < CH ? nown - c0 : CH;
            mbar_wait_fast(empty + 8 * ws.i, ws.ph ^ 1);
            if (no_load) {
              mbar_arrive(full + 8 * ws.i);
            } else {
              mbar_expect_tx(full + 8 * ws.i, nk * WI);
              bulk_load(ring + ws.i * kSlot, cb_img + (size_t)(ob + c0) * (WI / 4), nk * WI, full + 8 * ws.i);
            }
            ws.next(CF::WS);
          });
        });
      });
    }
  }
}

// ===================================================================== MMA issuer W (whole warp converged, one elected lane issues)
template <class CF, int W>
__device__ __forceinline__ void mma_issuer(const HsParams& p, const Bars& bars, uint32_t smem_base, uint32_t w_base, uint32_t tmem_base) {
  constexpr int NVAR = CF::NVAR, NGRP = CF::NGRP, CH = CF::CH, FORM = CF::FORM, CG = CF::CG, NI = CF::NI, G = CF::G, PW = CF::PW;
  constexpr uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint32_t idescN = idesc_base | ((uint32_t)(NI >> 3) << 17);
  constexpr uint32_t idesc2N = idesc_base | ((uint32_t)((2 * NI) >> 3) << 17);
  constexpr uint32_t HSTAGE_U = CF::HSTAGE >> 4, LO_U = CF::HALO_SLOT >> 4, SLOT_U = kSlot >> 4, W_U = CF::WI_BYTES >> 4;
  const uint32_t full = bars.wfull + W * 32, empty = bars.wempty + W * 32;
  const uint64_t adesc0 = make_kmajor_sw128_desc(smem_base, kSbo);                       // raw tile of halo stage 0
  const uint64_t bdesc0 = make_kmajor_sw128_desc(w_base + W * CF::WS * kSlot, 1024u);       // this issuer's weight slot 0
  const bool no_mma = (p.debug & 2) != 0;
  const int Cblks = p.Cblks;
  Ring hs{0, 0}, ws{0, 0}, ab{0, 0};
  uint32_t next_ok = 0;
  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
    const int var = NVAR > 1 ? (item / (CF::SLICED ? p.n_slices : 1)) % NVAR : 0;
    mbar_wait_fast(bars.accempty + 8 * ab.i, ab.ph ^ 1);        // the epilogue has drained this accumulator set
    tc_fence_after();
    const uint32_t acc0 = tmem_base + ab.i * CF::ACC_COLS;
    for (int cb = 0; cb < Cblks; ++cb) {
      // this issuer's accumulator pair(s); accum0 = 0 -> the pair's first MMA overwrites (zero-initialises) main | corr
      uint32_t acc_w, accum0;
      if (FORM == 0 && CF::COLSPLIT) { acc_w = acc0 + W * (G * PW) + (G == 2 ? (uint32_t)(cb & 1) * (uint32_t)PW : 0u); accum0 = cb >= G ? 1u : 0u; }
      else if (FORM == 0) { acc_w = acc0 + W * PW; accum0 = cb > 0 ? 1u : 0u; }
      else { acc_w = acc0 + (CF::COLSPLIT ? W * PW : 0); accum0 = cb > 0 ? 1u : 0u; }
      const bool last_cb = cb == Cblks - 1;
      uint64_t a_base = 0;
      static_for<0, NVAR>([&](auto VI) {
        constexpr int V = decltype(VI)::value;
        if (NVAR > 1 && var != V) return;
        constexpr int gi_first = CF::first_gi(W, V), gi_last = CF::last_gi(W, V);
        static_for<0, NGRP>([&](auto GI) {
          constexpr int gi = decltype(GI)::value;
          constexpr int g = FORM == 0 ? gi : V * CG + gi;
          constexpr int i0 = CF::i0(W, g), nown = CF::i1(W, g) - i0;
          if constexpr (nown > 0) {
            constexpr int nchunk = (nown + CH - 1) / CH;
            constexpr bool unit_first = FORM == 0 || gi == gi_first, unit_last = FORM == 0 || gi == gi_last;
            if (unit_first) {
              mbar_wait_fast(bars.hfull + 8 * hs.i, hs.ph);     // the issuing thread observes the TMA completion itself
              if constexpr (!CF::FAST) mbar_wait_fast(bars.hlo + 8 * hs.i, hs.ph);   // ... and the lo tile written from it
              a_base = adesc0 + (uint64_t)(hs.i * HSTAGE_U);
            }
            const uint32_t d_pair = acc_w + ((FORM == 1 && !CF::COLSPLIT) ? (uint32_t)(gi * PW) : 0u);
            static_for<0, nchunk>([&](auto CI) {
              constexpr int c0 = decltype(CI)::value * CH;
              constexpr int nk = nown - c0 < CH ? nown - c0 : CH;
              constexpr bool chunk_last = decltype(CI)::value == nchunk - 1;
              if (!next_ok) mbar_wait_slow(full + 8 * ws.i, ws.ph);   // next_ok: the probe issued behind the previous chunk's MMAs
              tc_fence_after();
              const uint64_t b_base = bdesc0 + (uint64_t)(ws.i * SLOT_U);
              if (elect_one()) {
                if (!no_mma) {
                  static_for<0, nk>([&](auto KI) {
                    constexpr int ki = decltype(KI)::value, i = i0 + c0 + ki;
                    constexpr KbGeom kg = kb_geom<FORM, CF::RP>(g, i);
                    // first k-block this issuer adds to the pair within the channel block
                    constexpr bool first_kb = (c0 + ki == 0) && (FORM == 1 || gi == 0);
                    const uint64_t a_raw = a_base + (uint64_t)kg.a_off16, b_img = b_base + (uint64_t)(ki * W_U);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {               // K = 8 tf32 per instruction = 32 bytes along the 128-byte row
                      if constexpr (CF::FAST) {
                        mma_tf32_ss(d_pair, a_raw + 2 * j, b_img + 2 * j, idescN, (first_kb && j == 0) ? accum0 : 1u);
                      } else {
                        mma_tf32_ss(d_pair, a_raw + 2 * j, b_img + 2 * j, idesc2N, (first_kb && j == 0) ? accum0 : 1u);
                        mma_tf32_ss(d_pair + NI, a_raw + LO_U + 2 * j, b_img + 2 * j, idescN, 1u);
                      }
                    }
                  });
                }
                tc_commit(empty + 8 * ws.i);                    // weight slot reusable once these MMAs retire
                if (unit_last && chunk_last) {
                  tc_commit(bars.hempty + 8 * hs.i);            // so is the halo stage (second arrival: the other issuer's)
                  if ((FORM == 1 || gi == NGRP - 1) && last_cb) tc_commit(bars.accfull + 8 * ab.i);
                }
              }
              __syncwarp();
              ws.next(CF::WS);
              // probe the NEXT chunk's barrier now: the ~90-cycle try_wait round trip overlaps the MMAs just queued
              next_ok = mbar_try(full + 8 * ws.i, ws.ph);
            });
            if (unit_last) hs.next(CF::HS);
          }
        });
      });
    }
    ab.next(CF::ACC_BUFS);
  }
}

template <class CF>
__global__ void __launch_bounds__(CF::THREADS, 1)
conv_halo_ss(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ HsParams p) {
  constexpr int N = CF::N, G = CF::G, FORM = CF::FORM, CG = CF::CG, NVAR = CF::NVAR, NI = CF::NI;
  constexpr int ACC_COLS = CF::ACC_COLS, ACC_BUFS = CF::ACC_BUFS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t w_base = smem_base + CF::HS * CF::HSTAGE;
  constexpr uint32_t misc_off = CF::HS * CF::HSTAGE + 2 * CF::WS * kSlot;
  const uint32_t misc = smem_base + misc_off;
  Bars bars;
  bars.hfull = misc;                      // 8 x 8
  bars.hlo = misc + 64;                   // 8 x 8
  bars.hempty = misc + 128;               // 8 x 8
  bars.wfull = misc + 192;                // 2 x 4 x 8
  bars.wempty = misc + 256;               // 2 x 4 x 8
  bars.accfull = misc + 320;              // 2 x 8
  bars.accempty = misc + 336;             // 2 x 8
  const uint32_t tmem_slot = misc + 352;
  static_assert(CF::HS <= 8 && CF::WS <= 4, "barrier layout");
  float* epi = reinterpret_cast<float*>(smem_gen + misc_off + 384);            // bias[N], scale[N], shift[N]
  float* stg_base = epi + (N == 32 ? 4 : 3) * (CF::SLICED ? p.n_total : N);                       // N = 32: [3 NT, 4 NT) head weights; then 4 warps x 32 x 36 staging

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Cblks = p.Cblks;

  if (threadIdx.x == 0) {
    for (int i = 0; i < CF::HS; ++i) { mbar_init(bars.hfull + 8 * i, 1); mbar_init(bars.hlo + 8 * i, 128); mbar_init(bars.hempty + 8 * i, 2); }
    for (int w = 0; w < 2; ++w)
      for (int i = 0; i < CF::WS; ++i) { mbar_init(bars.wfull + w * 32 + 8 * i, 1); mbar_init(bars.wempty + w * 32 + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bars.accfull + 8 * i, 2); mbar_init(bars.accempty + 8 * i, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 3) {
    tmem_alloc(tmem_slot, 512u);
    const int NT = CF::SLICED ? p.n_total : N;                               // constants of ALL columns (a CTA may serve several slices)
    for (int n = lane; n < NT; n += 32) {
      const float bias = p.bias ? p.bias[n] : 0.f, scale = p.gamma ? p.gamma[n] * p.bn_c : 1.f;
      epi[n] = bias;
      epi[NT + n] = scale;
      epi[2 * NT + n] = scale * bias + (p.beta ? p.beta[n] : 0.f);        // a = act(scale * acc + shift')
      if (N == 32 && n < N) epi[3 * NT + n] = p.head_out ? p.head_w[n] : 0.f;   // (the fused head exists for unsliced N = 32 only)
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================================================================== halo producer
    if (lane == 0) {
      prefetch_tmap(&tmap);
      Ring hs{0, 0};
      const bool no_load = (p.debug & 16) != 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int tile = item / (NVAR * (CF::SLICED ? p.n_slices : 1));
        const int twi = tile % p.tiles_w, thi = (tile / p.tiles_w) % p.tiles_h, b = tile / (p.tiles_w * p.tiles_h);
        const int s0 = twi * kTW - 1, r0 = thi * kTH - 1;     // halo origin (the zero fill outside the tensor == SAME padding)
        // PAIR: tile = image pair (2 tile, 2 tile + 1); an image index past the batch is zero-filled like the padding
        for (int cb = 0; cb < Cblks; ++cb) {
#pragma unroll
          for (int u = 0; u < CF::NUNITS; ++u) {
            const int c_plane = (FORM == 0 ? (u & 1) * p.C : 0) + cb * 32, h_plane = FORM == 0 ? (u >> 1) : 0;
            mbar_wait_fast(bars.hempty + 8 * hs.i, hs.ph ^ 1);
            if (no_load) {
              mbar_arrive(bars.hfull + 8 * hs.i);
            } else {
              mbar_expect_tx(bars.hfull + 8 * hs.i, CF::HALO_BYTES);
              if constexpr (CF::PAIR) tma_load_5d(smem_base + hs.i * CF::HSTAGE, &tmap, bars.hfull + 8 * hs.i, c_plane, -1, 2 * tile, -1, h_plane);
              else tma_load_5d(smem_base + hs.i * CF::HSTAGE, &tmap, bars.hfull + 8 * hs.i, c_plane, s0, h_plane, r0, b);
            }
            hs.next(CF::HS);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) weight_producer<CF, 0>(p, bars, w_base);
  } else if (warp == 12) {
    if (lane == 0) weight_producer<CF, 1>(p, bars, w_base);
  } else if (warp == 2) {
    mma_issuer<CF, 0>(p, bars, smem_base, w_base, tmem_base);
  } else if (warp == 3) {
    mma_issuer<CF, 1>(p, bars, smem_base, w_base, tmem_base);
  } else if (warp >= 4 && warp < 8 && !CF::FAST) {
    // ===================================================================== lo pass (once per halo tile, elementwise, same byte offsets)
    const int tid = threadIdx.x - 128;
    const bool skip = (p.debug & 1) != 0;
    const int units_per_item = Cblks * CF::NUNITS;
    Ring hs{0, 0};
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      for (int uu = 0; uu < units_per_item; ++uu) {
        mbar_wait_fast(bars.hfull + 8 * hs.i, hs.ph);
        if (!skip) {
          float4* raw = reinterpret_cast<float4*>(smem_gen + hs.i * CF::HSTAGE);
          float4* lo = reinterpret_cast<float4*>(smem_gen + hs.i * CF::HSTAGE + CF::HALO_SLOT);
#pragma unroll 4
          for (int i = tid; i < (int)(CF::HALO_BYTES / 16); i += 128) {
            const float4 v = raw[i];
            float4 h, l;
            h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
            l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
            raw[i] = h;                                         // hi = round-to-nearest tf32 (in place), lo = x - hi
            lo[i] = l;
          }
          fence_proxy_async();                                  // generic-proxy stores -> visible to the tensor core's operand reads
        }
        mbar_arrive(bars.hlo + 8 * hs.i);
        hs.next(CF::HS);
      }
    }
  } else if ((warp >= 8 && warp < 12) || warp >= 13) {
    // ===================================================================== epilogue (warpgroup wg owns accumulator set wg when there are two)
    const int wg = warp >= 13 ? 1 : 0;
    const int q = warp & 3;                                     // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;                              // tile row == TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float* stg = stg_base + (size_t)(wg * 4 + q) * 32 * 36;     // this warp's 32 x (32+4) staging rows
    constexpr int nchunks = N >> 5;
    constexpr int NCLS = FORM == 0 ? 1 : CG;
    // accumulator pairs that hold partial sums of the same output columns, and how far apart they are
    constexpr int NP = FORM == 0 ? (CF::COLSPLIT ? G : 2) : 1;
    constexpr int PSTRIDE = FORM == 0 ? CF::PW : 0;
    const bool no_store = (p.debug & 4) != 0;
    const int tw = row & (kTW - 1), th = CF::PAIR ? (row >> 4) : (row >> 3);   // PAIR: group 2 r + j = tile row r of image j
    const int act = p.act;
    // LeakyReLU / ReLU / identity as one select (the other activations take the generic path)
    const bool piecewise = act == UAD_ACT_NONE || act == UAD_ACT_LEAKY || act == UAD_ACT_RELU;
    const float slope = act == UAD_ACT_LEAKY ? p.alpha : (act == UAD_ACT_RELU ? 0.f : 1.f);
    Ring ab{(uint32_t)(CF::EPI_WG == 2 ? wg : 0), 0};
    for (int item = blockIdx.x + (CF::EPI_WG == 2 ? wg * gridDim.x : 0); item < p.n_items; item += CF::EPI_WG * gridDim.x) {
      const int nsl = CF::SLICED ? p.n_slices : 1;
      const int slice = CF::SLICED ? item % nsl : 0, var = NVAR > 1 ? (item / nsl) % NVAR : 0, tile = item / (NVAR * nsl);
      const int NT = CF::SLICED ? p.n_total : N, ncol0 = CF::SLICED ? slice * N : 0;   // output row stride, first output column of this item
      const int twi = tile % p.tiles_w, thi = (tile / p.tiles_w) % p.tiles_h;
      const int b = CF::PAIR ? 2 * tile + ((row >> 3) & 1) : tile / (p.tiles_w * p.tiles_h);
      const int s0 = CF::PAIR ? 0 : twi * kTW, r0 = CF::PAIR ? 0 : thi * kTH;
      const uint32_t acc_base = lane_base + ab.i * ACC_COLS;
      mbar_wait_fast(bars.accfull + 8 * ab.i, ab.ph);
      tc_fence_after();
      if (p.debug & 32) { tc_fence_before(); mbar_arrive(bars.accempty + 8 * ab.i); if (CF::EPI_WG == 2) ab.ph ^= 1; else ab.next(ACC_BUFS); continue; }
      const int cq = (lane & 7) * 4;
#pragma unroll 1
      for (int cls = 0; cls < NCLS; ++cls) {
        const int c = FORM == 0 ? 0 : var * CG + cls;           // output-parity class (p, q) = (c >> 1, c & 1)
        const long long my_off = (CF::PAIR && b >= p.B) ? -1ll :       // second image of the last pair of an odd batch: no output
            (((long long)b * p.OH + ((r0 + th) * p.osh + (c >> 1))) * p.OW + ((s0 + tw) * p.osh + (c & 1))) * (long long)NT + ncol0;
        long long offs[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) offs[it] = __shfl_sync(0xffffffffu, my_off, it * 4 + (lane >> 3));
#pragma unroll 1
        for (int ch = 0; ch < nchunks; ++ch) {
          const int c0 = ch * 32;
          // TMEM column of output column c0 in the first pair that holds it: main there, corr NI columns further
          uint32_t col;
          if (CF::COLSPLIT) col = (uint32_t)((c0 >> 6) * (FORM == 0 ? G * CF::PW : CF::PW) + (c0 & 63));
          else col = (uint32_t)((FORM == 1 ? cls * CF::PW : 0) + c0);
          uint32_t v[32], u[32];
          tmem_ld32(acc_base + col, v);
          if constexpr (CF::FAST) {
            if (NP == 2) tmem_ld32(acc_base + col + PSTRIDE, u);
            tmem_wait_ld();
            if (NP == 2) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
            }
          } else {
            tmem_ld32(acc_base + col + NI, u);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
            if (NP == 2) {
              uint32_t w2[32];
              tmem_ld32(acc_base + col + PSTRIDE, u);
              tmem_ld32(acc_base + col + PSTRIDE + NI, w2);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 32; ++j)
                v[j] = __float_as_uint(__uint_as_float(v[j]) + (__uint_as_float(u[j]) + __uint_as_float(w2[j])));
            }
          }
          if (cls + 1 == NCLS && ch + 1 == nchunks) {           // last TMEM read of this thread for the item: hand the set back
            tc_fence_before();
            mbar_arrive(bars.accempty + 8 * ab.i);
          }
          // transpose the RAW accumulators once through the warp's staging rows: afterwards a thread holds channels
          // c0 + cq .. + 3 of eight pixels, so bias / scale / shift / head weights are 16 registers (the per-element shared-memory
          // reads of the pre-transpose form were ~100 LDS per chunk and thread - the epilogue's largest instruction group)
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(stg + lane * 36 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          __syncwarp();
          const float4 e_bias = *reinterpret_cast<const float4*>(epi + ncol0 + c0 + cq);
          const float4 e_scale = *reinterpret_cast<const float4*>(epi + NT + ncol0 + c0 + cq);
          const float4 e_shift = *reinterpret_cast<const float4*>(epi + 2 * NT + ncol0 + c0 + cq);
          const float b4[4] = {e_bias.x, e_bias.y, e_bias.z, e_bias.w}, s4[4] = {e_scale.x, e_scale.y, e_scale.z, e_scale.w},
                      h4[4] = {e_shift.x, e_shift.y, e_shift.z, e_shift.w};
#pragma unroll 1
          for (int pass = 0; pass < 2; ++pass) {                // z then a from the SAME staged accumulators
            float* out = pass == 0 ? p.z_out : p.a_out;
            if (!out || (p.debug & 64)) continue;
            const bool plain = pass == 0;
            const bool with_head = FORM == 1 && N == 32 && !plain && p.head_out;
            // fused 1x1 head: x_hat[pixel] = sum_n a[pixel][n] * head_w[n] + head_b; a pixel's 32 channels lie in the 8 lanes that
            // share lane >> 3 (fixed order of the partial sums: deterministic)
            float4 hw = make_float4(0.f, 0.f, 0.f, 0.f);
            float hb = 0.f;
            if (with_head) { hw = *reinterpret_cast<const float4*>(epi + 3 * NT + cq); hb = p.head_b[0]; }
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const float4 r = *reinterpret_cast<const float4*>(stg + (it * 4 + (lane >> 3)) * 36 + cq);
              const float r4[4] = {r.x, r.y, r.z, r.w};
              float o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float t = plain ? r4[e] + b4[e] : fmaf(s4[e], r4[e], h4[e]);
                o[e] = (plain || piecewise) ? (t > 0.f || plain ? t : slope * t) : act_slow(t, act, p.alpha);
              }
              if (FORM == 1 && N == 32) {
                if (with_head) {
                  float h = fmaf(o[3], hw.w, fmaf(o[2], hw.z, fmaf(o[1], hw.y, o[0] * hw.x)));
                  h += __shfl_xor_sync(0xffffffffu, h, 1);
                  h += __shfl_xor_sync(0xffffffffu, h, 2);
                  h += __shfl_xor_sync(0xffffffffu, h, 4);
                  if ((lane & 7) == 0 && (!CF::PAIR || offs[it] >= 0)) p.head_out[offs[it] / N] = h + hb;
                }
              }
              if constexpr (CF::FAST) {                         // the next conv reads this tensor as a tf32 operand: store it rounded
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = tf32_rn(o[e]);
              }
              if (!no_store && (!CF::PAIR || offs[it] >= 0)) *reinterpret_cast<float4*>(out + offs[it] + c0 + cq) = make_float4(o[0], o[1], o[2], o[3]);
            }
          }
          __syncwarp();                                         // the staging rows are rewritten by the next chunk
        }
      }
      if (CF::EPI_WG == 2) ab.ph ^= 1; else ab.next(ACC_BUFS);   // a warpgroup that owns its set sees every one of its phases
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// raw weights -> per (32-channel block, issuer, k-block in the issuer's order): {hi, lo} images of [NI rows][32 k] fp32 in the
// SWIZZLE_128B byte order the descriptor expects (16-byte chunk index XOR (row & 7)).  transposed=false: raw[t][c][n];
// true: raw[t][n][c].  One thread per (cb, issuer-image, row, k); a column-split issuer W holds output columns [64 W, 64 W + 64).
// With column slices (NT = total output columns > N): the images of slice q (output columns [q N, q N + N)) follow those of slice q - 1.
__global__ void hs_weight_image_kernel(const float* __restrict__ w, float* __restrict__ img, HsOrder order, int C, int N, int NI,
                                       int transposed, int parts, int NT) {   // parts = 2: {hi, lo} images; 1: hi only (1xTF32)
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nimg = order.cnt[0] + order.cnt[1];                // images per channel block (25, or 50 halves when column-split)
  const size_t per_slice = (size_t)(C / 32) * nimg * NI * 32;
  const size_t total = per_slice * (NT / N);
  if (i >= total) return;
  const int slice = (int)(i / per_slice);
  const size_t is = i % per_slice;
  const int k = is % 32;
  const int r = (is / 32) % NI;
  const int im = (is / ((size_t)32 * NI)) % nimg;
  const int cb = is / ((size_t)32 * NI * nimg);
  const int W = im >= order.cnt[0] ? 1 : 0, o = W ? im - order.cnt[0] : im;
  const int t = order.wt[W][o];
  const int c = cb * 32 + k;
  const int n = slice * N + (NI < N ? W * NI : 0) + r;
  const float v = transposed ? w[((size_t)t * NT + n) * C + c] : w[((size_t)t * C + c) * NT + n];
  const uint32_t h = __float_as_uint(tf32_rn(v));
  const float lo = v - __uint_as_float(h);
  const size_t base = (((size_t)slice * (C / 32) + cb) * nimg + im) * parts * NI * 32;
  const int pos = r * 32 + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3));
  img[base + pos] = __uint_as_float(h);
  if (parts == 2) img[base + (size_t)NI * 32 + pos] = lo;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn hs_encode_fn() {
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

// weight-image order of a configuration + cross-check of the compiled geometry against the caller's runtime tap tables
template <class CF>
int build_order(const GatherParams& g, HsOrder& order) {
  memset(&order, 0, sizeof(order));
  for (int W = 0; W < 2; ++W) {
    order.cnt[W] = CF::cnt(W);
    for (int grp = 0; grp < 4; ++grp)
      for (int i = CF::i0(W, grp); i < CF::i1(W, grp); ++i) {
        const KbGeom kg = kb_geom<CF::FORM, CF::RP>(grp, i);
        order.wt[W][CF::ord_base(W, grp) + (i - CF::i0(W, grp))] = (unsigned char)kg.wt;
        const TapSet& ts = g.taps[CF::FORM == 0 ? 0 : grp];
        bool found = false;
        for (int t = 0; t < ts.n && !found; ++t) {
          if (ts.wt[t] != kg.wt) continue;
          const int dh = ts.dh[t], dw = ts.dw[t];
          int wh, wwd;
          if (CF::FORM == 0) {
            if ((dh & 1) != (grp >> 1) || (dw & 1) != (grp & 1)) continue;
            wh = (dh >> 1) + 1; wwd = (dw >> 1) + 1;
          } else {
            wh = dh + 1; wwd = dw + 1;
          }
          found = (wh * CF::RP + wwd) * 8 == kg.a_off16;
        }
        UAD_REQUIRE(found, "conv_halo_ss: tap table does not match the compiled geometry (group %d, k-block %d)", grp, i);
      }
  }
  UAD_REQUIRE(order.cnt[0] + order.cnt[1] == (CF::COLSPLIT ? 2 : 1) * kTaps, "conv_halo_ss: k-block ownership does not cover the filter");
  if (CF::FORM == 1)
    for (int c = 0; c < 4; ++c)
      UAD_REQUIRE(g.taps[c].oh0 == (c >> 1) && g.taps[c].ow0 == (c & 1) && g.taps[c].n == grp_nkb(c), "conv_halo_ss: class table mismatch");
  return 0;
}

template <class CF>
int launch_cfg(const GatherParams& g, const CUtensorMap& tmap, HsParams& p, const float* w_raw, bool weights_transposed, cudaStream_t st) {
  HsOrder order;
  if (int rc = build_order<CF>(g, order)) return rc;
  {
    const size_t total = (size_t)kTaps * p.C * p.n_total;
    hs_weight_image_kernel<<<uad_cdiv(total, 256), 256, 0, st>>>(w_raw, const_cast<float*>(p.wimg), order, p.C, CF::N, CF::NI,
                                                                 weights_transposed ? 1 : 0, CF::FAST ? 1 : 2, p.n_total);
    UAD_LAUNCH_CHECK("hs_weight_image");
  }
  // shared memory: halo stages (raw + lo), two weight rings, barriers / constants / staging
  UAD_REQUIRE(CF::SLICED || (p.n_total == CF::N && p.n_slices == 1), "conv_halo_ss: column slices need the sliced kernel");
  const size_t tail = 384 + (CF::N == 32 ? 4 : 3) * p.n_total * sizeof(float) + CF::EPI_WG * 4 * 32 * 36 * sizeof(float) + 64;
  const size_t smem = 1024 + CF::HS * CF::HSTAGE + 2 * CF::WS * kSlot + tail;
  UAD_REQUIRE(smem <= 227 * 1024, "conv_halo_ss: shared-memory budget exceeded");
  static bool attr = false;
  if (!attr) {
    UAD_CUDA(cudaFuncSetAttribute(conv_halo_ss<CF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  p.n_items = (CF::PAIR ? (p.B + 1) / 2 : p.tiles_w * p.tiles_h * p.B) * CF::NVAR * p.n_slices;
  const int grid = p.n_items < UAD_NUM_SMS ? p.n_items : UAD_NUM_SMS;
  conv_halo_ss<CF><<<grid, CF::THREADS, smem, st>>>(tmap, p);
  UAD_LAUNCH_CHECK("conv_halo_ss");
  return 0;
}

}  // namespace

int uad_hs_gather_supported(int Cin, int N, int lgMH, int lgMW, int nclasses) {
  if (Cin % 32 != 0 || Cin < 32) return 0;
  if (!(N == 32 || N == 64 || N == 128)) return 0;
  if (lgMW < 3 || lgMH < 3) return 0;          // M-grid at least 16 x 8 (one tile never spans two images) ...
  if (lgMH == 3 && lgMW != 3) return 0;        // ... or exactly 8 x 8: the image-pair layout (Cfg::PAIR)
  if (nclasses == 4 && Cin > 128) return 0;    // stride-1 form: one accumulator pair per class (K <= 9 * 128 per pair)
  return 1;
}

size_t uad_hs_gather_ws_bytes(int ksize, int Cin, int N) {
  if (Cin % 32 != 0) return 0;
  return (size_t)ksize * ksize * Cin * N * 2 * sizeof(float) + 1024;
}

int uad_launch_gather_hs(const GatherParams& g, int nclasses, int ksize, bool weights_transposed, const float* w_raw, bool fast,
                         void* ws, size_t ws_bytes, cudaStream_t st) {
  const int C = g.Cin, N = g.N;
  const size_t need = uad_hs_gather_ws_bytes(ksize, C, N);
  UAD_REQUIRE(ksize == 5, "conv_halo_ss: 5 x 5 filters only");
  UAD_REQUIRE(ws && ws_bytes >= need, "conv_halo_ss: workspace too small (%zu < %zu)", ws_bytes, need);
  UAD_REQUIRE(((uintptr_t)ws % 128) == 0 && ((uintptr_t)g.in % 16) == 0, "conv_halo_ss: unaligned buffers");
  UAD_REQUIRE(nclasses == 1 || nclasses == 4, "conv_halo_ss: %d tap classes", nclasses);
  UAD_REQUIRE((g.sh == 2) == (nclasses == 1), "conv_halo_ss: strided form has one class, stride-1 form four");
  EncodeTiledFn encode = hs_encode_fn();
  UAD_REQUIRE(encode != nullptr, "conv_halo_ss: cuTensorMapEncodeTiled entry point unavailable");
  const int form = g.sh == 2 ? 0 : 1;

  HsParams p;
  memset(&p, 0, sizeof(p));
  const int MW = 1 << g.lgMW, MH = 1 << g.lgMH;
  const bool pair = MH == 8 && MW == 8;
  p.tiles_w = MW / kTW;
  p.tiles_h = pair ? 1 : MH / kTH;
  p.B = g.B; p.C = C; p.Cblks = C / 32;
  // 8 x 8 M-grids: ceil(B / 2) tiles cannot occupy 148 SMs - the layer is computed in 32-column slices (N / 32 times the items)
  static int slice_pairs = -1;
  if (slice_pairs < 0) { const char* e = getenv("UAD_HS_SLICES"); slice_pairs = e ? atoi(e) : 1; }
  // (measured, 256^2 B = 64, N = C = 128: strided form 0.078 -> 0.046 ms sliced; stride-1 form 0.044 unsliced vs 0.049 sliced - its
  // items already come in four output-parity variants - so only the strided form is sliced)
  const bool sliced = pair && form == 0 && N > 32 && slice_pairs && !g.head_out;
  p.n_total = N; p.n_slices = sliced ? N / 32 : 1;
  p.OH = g.OH; p.OW = g.OW; p.osh = g.osh;
  p.z_out = g.z_out; p.a_out = g.a_out; p.bias = g.bias; p.gamma = g.gamma; p.beta = g.beta;
  p.bn_c = g.bn_c; p.alpha = g.alpha; p.act = g.act;
  p.head_w = g.head_w; p.head_b = g.head_b; p.head_out = g.head_out;
  UAD_REQUIRE(!g.head_out || (form == 1 && N == 32 && g.a_out && g.head_w && g.head_b), "conv_halo_ss: the fused 1x1 head needs the stride-1 form with N = 32");
  p.wimg = reinterpret_cast<float*>(ws);
  { const char* dbg = getenv("UAD_HS_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
  UAD_REQUIRE(p.z_out || p.a_out, "conv_halo_ss: no output requested");

  // 5-D tensor map over the NHWC input: (channel [x column parity], W, row parity / 1, H, B); box = (32 ch, 10, 1, 18, 1)
  CUtensorMap tmap;
  cuuint64_t dims[5], strides[4];
  const cuuint64_t e = sizeof(float);
  if (pair && form == 0) {          // (channel x column parity, W / 2, image, H / 2, row parity): box = both images' plane halos, row-interleaved
    dims[0] = 2ull * C; dims[1] = g.IW / 2; dims[2] = g.B; dims[3] = g.IH / 2; dims[4] = 2;
    strides[0] = 2ull * C * e; strides[1] = (cuuint64_t)g.IH * g.IW * C * e; strides[2] = 2ull * g.IW * C * e;
    strides[3] = (cuuint64_t)g.IW * C * e;
  } else if (pair) {                // (channel, W, image, H, 1)
    dims[0] = C; dims[1] = g.IW; dims[2] = g.B; dims[3] = g.IH; dims[4] = 1;
    strides[0] = (cuuint64_t)C * e; strides[1] = (cuuint64_t)g.IH * g.IW * C * e; strides[2] = (cuuint64_t)g.IW * C * e;
    strides[3] = (cuuint64_t)g.IH * g.IW * C * e;
  } else if (form == 0) {
    dims[0] = 2ull * C; dims[1] = g.IW / 2; dims[2] = 2; dims[3] = g.IH / 2; dims[4] = g.B;
    strides[0] = 2ull * C * e; strides[1] = (cuuint64_t)g.IW * C * e; strides[2] = 2ull * g.IW * C * e;
    strides[3] = (cuuint64_t)g.IH * g.IW * C * e;
  } else {
    dims[0] = C; dims[1] = g.IW; dims[2] = 1; dims[3] = g.IH; dims[4] = g.B;
    strides[0] = (cuuint64_t)C * e; strides[1] = (cuuint64_t)g.IW * C * e; strides[2] = (cuuint64_t)g.IW * C * e;
    strides[3] = (cuuint64_t)g.IH * g.IW * C * e;
  }
  cuuint32_t box[5] = {32u, (cuuint32_t)kHW, 1u, (cuuint32_t)kHH, 1u};
  if (pair) { box[2] = 2u; box[3] = (cuuint32_t)(kTW + 2); }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(g.in), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UAD_REQUIRE(cr == CUDA_SUCCESS, "conv_halo_ss: cuTensorMapEncodeTiled failed (%d)", (int)cr);

  if (pair) {   // 8 x 8 M-grids: two images per tile
    if (form == 1) UAD_REQUIRE(C <= 128, "conv_halo_ss: stride-1 form supports at most 128 input channels");
#define UAD_HS_PAIR(FAST)                                                                                                   \
    if (form == 0) {                                                                                                         \
      if (N == 32) return launch_cfg<Cfg<0, 32, 1, 1, FAST, true>>(g, tmap, p, w_raw, weights_transposed, st);               \
      if (N == 64) return launch_cfg<Cfg<0, 64, 1, 1, FAST, true>>(g, tmap, p, w_raw, weights_transposed, st);               \
      if (!FAST && p.Cblks >= 2) return launch_cfg<Cfg<0, 128, 1, 2, false, true>>(g, tmap, p, w_raw, weights_transposed, st); \
      return launch_cfg<Cfg<0, 128, 1, 1, FAST, true>>(g, tmap, p, w_raw, weights_transposed, st);                           \
    }                                                                                                                        \
    if (N == 32) return launch_cfg<Cfg<1, 32, 4, 1, FAST, true>>(g, tmap, p, w_raw, weights_transposed, st);                 \
    if (N == 64) return launch_cfg<Cfg<1, 64, 2, 1, FAST, true>>(g, tmap, p, w_raw, weights_transposed, st);                 \
    return launch_cfg<Cfg<1, 128, 1, 1, FAST, true>>(g, tmap, p, w_raw, weights_transposed, st);
    if (sliced)
      return fast ? launch_cfg<Cfg<0, 32, 1, 1, true, true, true>>(g, tmap, p, w_raw, weights_transposed, st)
                  : launch_cfg<Cfg<0, 32, 1, 1, false, true, true>>(g, tmap, p, w_raw, weights_transposed, st);
    if (fast) { UAD_HS_PAIR(true) }
    UAD_HS_PAIR(false)
#undef UAD_HS_PAIR
  }
  if (fast) {   // UAD_MATH_TC_1XTF32
    if (form == 0) {
      if (N == 32) return launch_cfg<Cfg<0, 32, 1, 1, true>>(g, tmap, p, w_raw, weights_transposed, st);
      if (N == 64) return launch_cfg<Cfg<0, 64, 1, 1, true>>(g, tmap, p, w_raw, weights_transposed, st);
      return launch_cfg<Cfg<0, 128, 1, 1, true>>(g, tmap, p, w_raw, weights_transposed, st);
    }
    UAD_REQUIRE(C <= 128, "conv_halo_ss: stride-1 form supports at most 128 input channels");
    if (N == 32) return launch_cfg<Cfg<1, 32, 4, 1, true>>(g, tmap, p, w_raw, weights_transposed, st);
    if (N == 64) return launch_cfg<Cfg<1, 64, 2, 1, true>>(g, tmap, p, w_raw, weights_transposed, st);
    return launch_cfg<Cfg<1, 128, 1, 1, true>>(g, tmap, p, w_raw, weights_transposed, st);
  }
  if (form == 0) {
    if (N == 32) return launch_cfg<Cfg<0, 32, 1, 1>>(g, tmap, p, w_raw, weights_transposed, st);
    if (N == 64) return launch_cfg<Cfg<0, 64, 1, 1>>(g, tmap, p, w_raw, weights_transposed, st);
    return p.Cblks >= 2 ? launch_cfg<Cfg<0, 128, 1, 2>>(g, tmap, p, w_raw, weights_transposed, st)
                        : launch_cfg<Cfg<0, 128, 1, 1>>(g, tmap, p, w_raw, weights_transposed, st);
  }
  UAD_REQUIRE(C <= 128, "conv_halo_ss: stride-1 form supports at most 128 input channels");
  if (N == 32) return launch_cfg<Cfg<1, 32, 4, 1>>(g, tmap, p, w_raw, weights_transposed, st);
  if (N == 64) return launch_cfg<Cfg<1, 64, 2, 1>>(g, tmap, p, w_raw, weights_transposed, st);
  return launch_cfg<Cfg<1, 128, 1, 1>>(g, tmap, p, w_raw, weights_transposed, st);
}
