// Halo-resident SS-form implicit GEMM on tcgen05 for the conv family (Form F and Form T of uad_conv.cuh), sm_100a only.
//
// Built on three hardware answers of tools/ubench/operand_probe.cu (profiles/r2_operand_probe.txt):
//   E1      kind::tf32 TRUNCATES fp32 operand words  -> a TMA-loaded fp32 tile IS the tf32 'hi' operand; only lo = x - trunc(x)
//           has to be produced (one elementwise pass over the tile, same byte offsets)
//   E6/E8   the 128-byte swizzle is a function of the absolute shared-memory address -> a K-major descriptor may start at ANY
//           128-byte row of a TMA-written tile and its 8-row groups may be any number of rows apart (SBO = halo pitch)
// and on one measurement of this kernel's first version (profiles/r2_hs_issue_loop.md): with a table-driven issue loop (tap
// tables in constant memory, ~100 dependent SASS instructions per k-block) the MMA-issuing warp was instruction-latency bound -
// 880 cycles per k-block with the tensor pipe 25 % busy, and switching off the loads, the lo pass, the stores AND the MMAs
// still left 55 % of the time.  The tap geometry is therefore COMPILE-TIME here: the kernel is a template over the form and
// every k-block of a work item is unrolled, so window offsets, accumulator columns and weight-slot offsets are immediates and a
// k-block costs 8 UTCHMMA + ~12 uniform adds.
//
// Tile = 16 x 8 pixels of the M-grid (M = 128 rows; one 8-pixel tile row = one 8-row descriptor group).  Per 32-channel block
// the (16+2) x (8+2) pixel halo of the tile (stride-1 form) or of one stride-2 parity plane (strided form) is TMA-loaded ONCE
// as 180 rows of 128 bytes (SWIZZLE_128B); every filter tap that reads from it is a descriptor START ADDRESS
// (halo + (wh * 10 + ww) * 128 bytes, SBO = 1280).  fp32 parity = 3xTF32 with the paired-B trick:
//     acc[main | corr] (+)= A_raw . [B_hi ; B_lo]^T          one MMA of width 2N  (hi*hi -> main, hi*lo -> corr)
//     acc[corr]         +=  A_lo  .  B_hi^T                  one MMA of width N
// Roles (384 threads, one persistent CTA per SM):
//   warp 0      halo TMA producer                                   -> h_full[s]
//   warps 4-7   lo pass: lo tile = raw - trunc(raw)                 -> h_lo[s]            (once per halo, NOT per tap)
//   warp 1      weight producer (one bulk copy per CHUNK of k-blocks) -> w_full[t]
//   warp 2      MMA issuer: 8 x tcgen05.mma per k-block, commits    -> w_empty[t], h_empty[s], acc_full[b]
//   warp 3      TMEM allocation, epilogue constants
//   warps 8-11  epilogue: tcgen05.ld, sum of the accumulators (RN), bias / frozen-BN / activation, smem transpose, coalesced
//               row stores                                          -> acc_empty[b]
// The tensor core adds into its fp32 accumulator with truncation (round 1, measured): the strided form deals its 32-channel
// blocks round-robin over G = 2 accumulator pairs when there are several; the epilogue sums them in registers.
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
constexpr int kThreads = 384;
constexpr int kTaps = 25;

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
// The weight images are laid out in exactly this order ([channel block][group][k-block]) so a chunk of consecutive k-blocks
// is one bulk copy.
struct KbGeom { int a_off16, wt; };
__host__ __device__ constexpr int grp_nkb(int g) { return (2 + (g >> 1)) * (2 + (g & 1)); }
__host__ __device__ constexpr int grp_base(int g) { return g == 0 ? 0 : (g == 1 ? 4 : (g == 2 ? 10 : 16)); }
template <int FORM>
__host__ __device__ constexpr KbGeom kb_geom(int g, int i) {
  const int p = g >> 1, q = g & 1, nw = 2 + q, ih = i / nw, iw = i % nw;
  if (FORM == 0) {
    const int wh = 1 - p + ih, ww = 1 - q + iw;
    return KbGeom{(wh * kHW + ww) * 8, (2 * wh + p - 1) * 5 + (2 * ww + q - 1)};
  }
  return KbGeom{(ih * kHW + iw) * 8, (p + 3 - 2 * ih) * 5 + (q + 3 - 2 * iw)};
}

struct HsOrder { unsigned char wt[kTaps]; };                 // weight tap of the o-th k-block image of a channel block

template <int FORM_, int N_, int CG_, int G_>
struct Cfg {
  static constexpr int FORM = FORM_, N = N_, CG = CG_, G = G_;
  static constexpr int NVAR = FORM == 0 ? 1 : 4 / CG;        // item = tile * NVAR + variant (class group)
  static constexpr int NGRP = FORM == 0 ? 4 : CG;            // k-block groups per (item, channel block)
  static constexpr int NUNITS = FORM == 0 ? 4 : 1;           // halo loads per (item, channel block)
  static constexpr int CH = N == 32 ? 2 : 1;                 // k-blocks per weight chunk (one barrier round trip per chunk)
  static constexpr uint32_t W_BYTES = 2u * N * 128u;         // one k-block's {hi, lo} weight image
  static constexpr uint32_t SLOT = CH * W_BYTES;             // weight ring slot
  static constexpr int ACC_COLS = (FORM == 0 ? 1 : CG) * G * 2 * N;
  static constexpr int ACC_BUFS = 2 * ACC_COLS <= 512 ? 2 : 1;
  static_assert(ACC_COLS <= 512, "TMEM budget");
};

struct HsParams {
  int tiles_w, tiles_h;
  int B, C, Cblks;
  int OH, OW, osh;
  int h_stages, w_stages;
  int n_items;
  int debug;           // developer timing switches (UAD_HS_DEBUG): 1 = no lo pass, 2 = no MMAs, 4 = no global stores, 8 = no weight loads, 16 = no halo loads
  float* z_out;
  float* a_out;
  const float* bias;
  const float* gamma;
  const float* beta;
  float bn_c, alpha;
  int act;
  const float* wimg;   // [Cblks][25 k-blocks in group order][2][N][32] pre-swizzled {hi, lo} weight images
};

struct Ring {                                                // ring position + phase parity
  uint32_t i, ph;
  __device__ __forceinline__ void next(uint32_t n) { if (++i == n) { i = 0; ph ^= 1; } }
};

template <class CF>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_ss(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ HsParams p) {
  constexpr int N = CF::N, G = CF::G, FORM = CF::FORM, CG = CF::CG, NVAR = CF::NVAR, NGRP = CF::NGRP, CH = CF::CH;
  constexpr uint32_t W_BYTES = CF::W_BYTES, SLOT = CF::SLOT;
  constexpr int ACC_COLS = CF::ACC_COLS, ACC_BUFS = CF::ACC_BUFS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t HS = p.h_stages, WS = p.w_stages;
  const uint32_t w_base = smem_base + HS * kHaloStage;
  const uint32_t misc_off = HS * kHaloStage + WS * SLOT;
  const uint32_t misc = smem_base + misc_off;
  const uint32_t bar_hfull = misc;                      // 4 x 8
  const uint32_t bar_hlo = misc + 32;                   // 4 x 8
  const uint32_t bar_hempty = misc + 64;                // 4 x 8
  const uint32_t bar_wfull = misc + 96;                 // 8 x 8
  const uint32_t bar_wempty = misc + 160;               // 8 x 8
  const uint32_t bar_accfull = misc + 224;              // 2 x 8
  const uint32_t bar_accempty = misc + 240;             // 2 x 8
  const uint32_t tmem_slot = misc + 256;
  float* epi = reinterpret_cast<float*>(smem_gen + misc_off + 320);            // bias[N], scale[N], shift[N]
  float* stg_base = epi + 3 * N;                                               // 4 warps x 32 x 36 staging

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Cblks = p.Cblks;

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < HS; ++i) { mbar_init(bar_hfull + 8 * i, 1); mbar_init(bar_hlo + 8 * i, 128); mbar_init(bar_hempty + 8 * i, 1); }
    for (uint32_t i = 0; i < WS; ++i) { mbar_init(bar_wfull + 8 * i, 1); mbar_init(bar_wempty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_accfull + 8 * i, 1); mbar_init(bar_accempty + 8 * i, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 3) {
    tmem_alloc(tmem_slot, 512u);
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
    // ===================================================================== halo producer
    if (lane == 0) {
      prefetch_tmap(&tmap);
      Ring hs{0, 0};
      const bool no_load = (p.debug & 16) != 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int tile = item / NVAR;
        const int twi = tile % p.tiles_w, thi = (tile / p.tiles_w) % p.tiles_h, b = tile / (p.tiles_w * p.tiles_h);
        const int s0 = twi * kTW - 1, r0 = thi * kTH - 1;     // halo origin (the zero fill outside the tensor == SAME padding)
        for (int cb = 0; cb < Cblks; ++cb) {
#pragma unroll
          for (int u = 0; u < CF::NUNITS; ++u) {
            const int c_plane = (FORM == 0 ? (u & 1) * p.C : 0) + cb * 32, h_plane = FORM == 0 ? (u >> 1) : 0;
            mbar_wait(bar_hempty + 8 * hs.i, hs.ph ^ 1);
            if (no_load) {
              mbar_arrive(bar_hfull + 8 * hs.i);
            } else {
              mbar_expect_tx(bar_hfull + 8 * hs.i, kHaloBytes);
              tma_load_5d(smem_base + hs.i * kHaloStage, &tmap, bar_hfull + 8 * hs.i, c_plane, s0, h_plane, r0, b);
            }
            hs.next(HS);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== weight producer: one bulk copy per chunk
    if (lane == 0) {
      Ring ws{0, 0};
      const bool no_load = (p.debug & 8) != 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int var = NVAR > 1 ? item % NVAR : 0;
        for (int cb = 0; cb < Cblks; ++cb) {
          const float* cb_img = p.wimg + (size_t)cb * kTaps * (W_BYTES / 4);
          static_for<0, NVAR>([&](auto VI) {
            constexpr int V = decltype(VI)::value;
            if (NVAR > 1 && var != V) return;
            static_for<0, NGRP>([&](auto GI) {
              constexpr int g = FORM == 0 ? decltype(GI)::value : V * CG + decltype(GI)::value;
              constexpr int nkb = grp_nkb(g);
              static_for<0, (nkb + CH - 1) / CH>([&](auto CI) {
                constexpr int c0 = decltype(CI)::value * CH;
                constexpr int nk = nkb - c0 < CH ? nkb - c0 : CH;
                mbar_wait(bar_wempty + 8 * ws.i, ws.ph ^ 1);
                if (no_load) {
                  mbar_arrive(bar_wfull + 8 * ws.i);
                } else {
                  mbar_expect_tx(bar_wfull + 8 * ws.i, nk * W_BYTES);
                  bulk_load(w_base + ws.i * SLOT, cb_img + (size_t)(grp_base(g) + c0) * (W_BYTES / 4), nk * W_BYTES, bar_wfull + 8 * ws.i);
                }
                ws.next(WS);
              });
            });
          });
        }
      }
    }
  } else if (warp == 2) {
    // ===================================================================== MMA issuer (whole warp converged, one elected lane issues)
    constexpr uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idescN = idesc_base | ((uint32_t)(N >> 3) << 17);
    constexpr uint32_t idesc2N = idesc_base | ((uint32_t)((2 * N) >> 3) << 17);
    const uint64_t adesc0 = make_kmajor_sw128_desc(smem_base, kSbo);           // raw tile of halo stage 0
    const uint64_t bdesc0 = make_kmajor_sw128_desc(w_base, 1024u);             // weight slot 0: N hi rows, N lo rows per k-block
    constexpr uint32_t HSTAGE_U = kHaloStage >> 4, LO_U = kHaloSlot >> 4, SLOT_U = SLOT >> 4, W_U = W_BYTES >> 4;
    const bool no_mma = (p.debug & 2) != 0;
    Ring hs{0, 0}, ws{0, 0}, ab{0, 0};
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int var = NVAR > 1 ? item % NVAR : 0;
      mbar_wait(bar_accempty + 8 * ab.i, ab.ph ^ 1);           // the epilogue has drained this accumulator set
      tc_fence_after();
      const uint32_t acc0 = tmem_base + ab.i * ACC_COLS;
      for (int cb = 0; cb < Cblks; ++cb) {
        const uint32_t acc_g = acc0 + (G == 2 ? (uint32_t)(cb & 1) * 2 * N : 0u);   // this channel block's accumulator pair
        const uint32_t accum0 = cb >= G ? 1u : 0u;             // 0 -> the pair's first MMA overwrites (zero-initialises) main | corr
        const bool last_cb = cb == Cblks - 1;
        uint64_t a_base = 0;
        static_for<0, NVAR>([&](auto VI) {
          constexpr int V = decltype(VI)::value;
          if (NVAR > 1 && var != V) return;
          static_for<0, NGRP>([&](auto GI) {
            constexpr int gi = decltype(GI)::value;
            constexpr int g = FORM == 0 ? gi : V * CG + gi;
            constexpr int nkb = grp_nkb(g);
            constexpr int nchunk = (nkb + CH - 1) / CH;
            constexpr bool unit_first = FORM == 0 || gi == 0, unit_last = FORM == 0 || gi == NGRP - 1;
            if (unit_first) {
              mbar_wait(bar_hfull + 8 * hs.i, hs.ph);          // the issuing thread observes the TMA completion itself
              mbar_wait(bar_hlo + 8 * hs.i, hs.ph);            // ... and the lo tile written from it
              a_base = adesc0 + (uint64_t)(hs.i * HSTAGE_U);
            }
            const uint32_t d_pair = acc_g + (FORM == 0 ? 0u : (uint32_t)(gi * G * 2 * N));
            static_for<0, nchunk>([&](auto CI) {
              constexpr int c0 = decltype(CI)::value * CH;
              constexpr int nk = nkb - c0 < CH ? nkb - c0 : CH;
              constexpr bool chunk_last = decltype(CI)::value == nchunk - 1;
              mbar_wait(bar_wfull + 8 * ws.i, ws.ph);
              tc_fence_after();
              const uint64_t b_base = bdesc0 + (uint64_t)(ws.i * SLOT_U);
              if (elect_one()) {
                if (!no_mma) {
                  static_for<0, nk>([&](auto KI) {
                    constexpr int ki = decltype(KI)::value, i = c0 + ki;
                    constexpr KbGeom kg = kb_geom<FORM>(g, i);
                    constexpr bool first_kb = i == 0 && (FORM == 1 || gi == 0);   // first k-block of this accumulator pair in the channel block
                    const uint64_t a_raw = a_base + (uint64_t)kg.a_off16, a_lo = a_raw + LO_U, b_img = b_base + (uint64_t)(ki * W_U);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {               // K = 8 tf32 per instruction = 32 bytes along the 128-byte row
                      mma_tf32_ss(d_pair, a_raw + 2 * j, b_img + 2 * j, idesc2N, (first_kb && j == 0) ? accum0 : 1u);
                      mma_tf32_ss(d_pair + N, a_lo + 2 * j, b_img + 2 * j, idescN, 1u);
                    }
                  });
                }
                tc_commit(bar_wempty + 8 * ws.i);               // weight slot reusable once these MMAs retire
                if (unit_last && chunk_last) {
                  tc_commit(bar_hempty + 8 * hs.i);             // so is the halo stage after the unit's last k-block
                  if ((FORM == 1 || gi == NGRP - 1) && last_cb) tc_commit(bar_accfull + 8 * ab.i);
                }
              }
              __syncwarp();
              ws.next(WS);
            });
            if (unit_last) hs.next(HS);
          });
        });
      }
      ab.next(ACC_BUFS);
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================================================================== lo pass (once per halo tile, elementwise, same byte offsets)
    const int tid = threadIdx.x - 128;
    const bool skip = (p.debug & 1) != 0;
    const int units_per_item = Cblks * CF::NUNITS;
    Ring hs{0, 0};
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      for (int uu = 0; uu < units_per_item; ++uu) {
        mbar_wait(bar_hfull + 8 * hs.i, hs.ph);
        if (!skip) {
          const float4* raw = reinterpret_cast<const float4*>(smem_gen + hs.i * kHaloStage);
          float4* lo = reinterpret_cast<float4*>(smem_gen + hs.i * kHaloStage + kHaloSlot);
#pragma unroll 4
          for (int i = tid; i < (int)(kHaloBytes / 16); i += 128) {
            const float4 v = raw[i];
            float4 l;
            l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
            l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
            l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
            l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
            lo[i] = l;
          }
          fence_proxy_async();                                  // generic-proxy stores -> visible to the tensor core's operand reads
        }
        mbar_arrive(bar_hlo + 8 * hs.i);
        hs.next(HS);
      }
    }
  } else if (warp >= 8) {
    // ===================================================================== epilogue
    const int row = threadIdx.x - 256;                          // tile row == TMEM lane
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float* stg = stg_base + (size_t)q * 32 * 36;                // this warp's 32 x (32+4) staging rows
    constexpr int nchunks = N >> 5;
    constexpr int NCLS = FORM == 0 ? 1 : CG;
    const bool no_store = (p.debug & 4) != 0;
    const int tw = row & (kTW - 1), th = row >> 3;
    const int act = p.act;
    const float alpha = p.alpha;
    Ring ab{0, 0};
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int var = NVAR > 1 ? item % NVAR : 0, tile = item / NVAR;
      const int twi = tile % p.tiles_w, thi = (tile / p.tiles_w) % p.tiles_h, b = tile / (p.tiles_w * p.tiles_h);
      const int s0 = twi * kTW, r0 = thi * kTH;
      const uint32_t acc_base = lane_base + ab.i * ACC_COLS;
      mbar_wait(bar_accfull + 8 * ab.i, ab.ph);
      tc_fence_after();
      const int cq = (lane & 7) * 4;
#pragma unroll 1
      for (int cls = 0; cls < NCLS; ++cls) {
        const int c = FORM == 0 ? 0 : var * CG + cls;           // output-parity class (p, q) = (c >> 1, c & 1)
        const long long my_off =
            (((long long)b * p.OH + ((r0 + th) * p.osh + (c >> 1))) * p.OW + ((s0 + tw) * p.osh + (c & 1))) * (long long)N;
        long long offs[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) offs[it] = __shfl_sync(0xffffffffu, my_off, it * 4 + (lane >> 3));
        const uint32_t cls_base = acc_base + cls * G * 2 * N;
#pragma unroll 1
        for (int ch = 0; ch < nchunks; ++ch) {
          const int c0 = ch * 32;
          uint32_t v[32], u[32];
          tmem_ld32(cls_base + c0, v);                          // pair 0: main columns [0, N), corr columns [N, 2N)
          tmem_ld32(cls_base + N + c0, u);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
          if (G == 2) {
            uint32_t w2[32];
            tmem_ld32(cls_base + 2 * N + c0, u);
            tmem_ld32(cls_base + 3 * N + c0, w2);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = __float_as_uint(__uint_as_float(v[j]) + (__uint_as_float(u[j]) + __uint_as_float(w2[j])));
          }
          if (cls + 1 == NCLS && ch + 1 == nchunks) {           // last TMEM read of this thread for the item: hand the set back
            tc_fence_before();
            mbar_arrive(bar_accempty + 8 * ab.i);
          }
#pragma unroll 1
          for (int pass = 0; pass < 2; ++pass) {                // z then a from the SAME registers
            float* out = pass == 0 ? p.z_out : p.a_out;
            if (!out) continue;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int n = c0 + j + e;
                const float z = __uint_as_float(v[j + e]) + epi[n];
                o[e] = pass == 0 ? z : uad_act(epi[N + n] * z + epi[2 * N + n], act, alpha);
              }
              *reinterpret_cast<float4*>(stg + lane * 36 + j) = make_float4(o[0], o[1], o[2], o[3]);
            }
            __syncwarp();
            float4 vals[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) vals[it] = *reinterpret_cast<const float4*>(stg + (it * 4 + (lane >> 3)) * 36 + cq);
            if (!no_store) {
#pragma unroll
              for (int it = 0; it < 8; ++it) *reinterpret_cast<float4*>(out + offs[it] + c0 + cq) = vals[it];
            }
            __syncwarp();
          }
        }
      }
      ab.next(ACC_BUFS);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// raw weights -> per (32-channel block, k-block in group order): {hi, lo} images of [N rows][32 k] fp32 in the SWIZZLE_128B byte
// order the descriptor expects (16-byte chunk index XOR (row & 7)).  transposed=false: raw[t][c][n]; true: raw[t][n][c].
__global__ void hs_weight_image_kernel(const float* __restrict__ w, float* __restrict__ img, HsOrder order, int C, int N, int transposed) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)kTaps * C * N;
  if (i >= total) return;
  const int k = i % 32;
  const int n = (i / 32) % N;
  const int o = (i / ((size_t)32 * N)) % kTaps;
  const int cb = i / ((size_t)32 * N * kTaps);
  const int t = order.wt[o];
  const int c = cb * 32 + k;
  const float v = transposed ? w[((size_t)t * N + n) * C + c] : w[((size_t)t * C + c) * N + n];
  const uint32_t h = __float_as_uint(v) & 0xffffe000u;
  const float lo = v - __uint_as_float(h);
  const size_t base = ((size_t)(cb * kTaps + o)) * 2 * N * 32;
  const int pos = n * 32 + ((((k >> 2) ^ (n & 7)) << 2) | (k & 3));
  img[base + pos] = __uint_as_float(h);
  img[base + (size_t)N * 32 + pos] = lo;
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

template <class CF>
int launch_cfg(const CUtensorMap& tmap, HsParams& p, cudaStream_t st) {
  // shared-memory budget: halo stages (raw + lo), weight slots, barriers / constants / staging
  const size_t tail = 320 + 3 * CF::N * sizeof(float) + 4 * 32 * 36 * sizeof(float) + 64;
  const size_t budget = 227 * 1024 - 1024 - tail;
  // the stride-1 form loads one halo per 4 .. 25 k-blocks and N = 128 spends >= 3000 cycles per halo: two stages cover the
  // refill there and leave room for a deeper weight ring; the strided form at N <= 64 turns a halo over every 4 .. 9 short k-blocks
  p.h_stages = (CF::FORM == 1 || CF::N == 128) ? 2 : 3;
  p.w_stages = (int)((budget - p.h_stages * kHaloStage) / CF::SLOT);
  if (p.w_stages > 8) p.w_stages = 8;
  UAD_REQUIRE(p.w_stages >= 2, "conv_halo_ss: shared-memory budget exceeded");
  const size_t smem = 1024 + p.h_stages * kHaloStage + p.w_stages * CF::SLOT + tail;
  static bool attr = false;
  if (!attr) {
    UAD_CUDA(cudaFuncSetAttribute(conv_halo_ss<CF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  const int grid = p.n_items < UAD_NUM_SMS ? p.n_items : UAD_NUM_SMS;
  conv_halo_ss<CF><<<grid, kThreads, smem, st>>>(tmap, p);
  UAD_LAUNCH_CHECK("conv_halo_ss");
  return 0;
}

}  // namespace

int uad_hs_gather_supported(int Cin, int N, int lgMH, int lgMW, int nclasses) {
  if (Cin % 32 != 0 || Cin < 32) return 0;
  if (!(N == 32 || N == 64 || N == 128)) return 0;
  if (lgMH < 4 || lgMW < 3) return 0;          // M-grid at least 16 x 8: one tile never spans two images
  if (nclasses == 4 && Cin > 128) return 0;    // stride-1 form: one accumulator pair per class (K <= 9 * 128 per pair)
  return 1;
}

size_t uad_hs_gather_ws_bytes(int ksize, int Cin, int N) {
  if (Cin % 32 != 0) return 0;
  return (size_t)ksize * ksize * Cin * N * 2 * sizeof(float) + 1024;
}

int uad_launch_gather_hs(const GatherParams& g, int nclasses, int ksize, bool weights_transposed, const float* w_raw,
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

  // the compile-time tap geometry must be the caller's tap tables (uad_conv_api.cu: taps_full / taps_parity)
  HsOrder order;
  for (int grp = 0; grp < 4; ++grp)
    for (int i = 0; i < grp_nkb(grp); ++i) {
      const KbGeom kg = form == 0 ? kb_geom<0>(grp, i) : kb_geom<1>(grp, i);
      order.wt[grp_base(grp) + i] = (unsigned char)kg.wt;
      const TapSet& ts = g.taps[form == 0 ? 0 : grp];
      bool found = false;
      for (int t = 0; t < ts.n && !found; ++t) {
        if (ts.wt[t] != kg.wt) continue;
        const int dh = ts.dh[t], dw = ts.dw[t];
        int wh, wwd;
        if (form == 0) {
          if ((dh & 1) != (grp >> 1) || (dw & 1) != (grp & 1)) continue;
          wh = (dh >> 1) + 1; wwd = (dw >> 1) + 1;
        } else {
          wh = dh + 1; wwd = dw + 1;
        }
        found = (wh * kHW + wwd) * 8 == kg.a_off16;
      }
      UAD_REQUIRE(found, "conv_halo_ss: tap table does not match the compiled geometry (group %d, k-block %d)", grp, i);
    }
  if (form == 1)
    for (int c = 0; c < 4; ++c)
      UAD_REQUIRE(g.taps[c].oh0 == (c >> 1) && g.taps[c].ow0 == (c & 1) && g.taps[c].n == grp_nkb(c), "conv_halo_ss: class table mismatch");

  float* img = reinterpret_cast<float*>(ws);
  {
    const size_t total = (size_t)kTaps * C * N;
    hs_weight_image_kernel<<<uad_cdiv(total, 256), 256, 0, st>>>(w_raw, img, order, C, N, weights_transposed ? 1 : 0);
    UAD_LAUNCH_CHECK("hs_weight_image");
  }

  HsParams p;
  memset(&p, 0, sizeof(p));
  const int MW = 1 << g.lgMW, MH = 1 << g.lgMH;
  p.tiles_w = MW / kTW;
  p.tiles_h = MH / kTH;
  p.B = g.B; p.C = C; p.Cblks = C / 32;
  p.OH = g.OH; p.OW = g.OW; p.osh = g.osh;
  p.z_out = g.z_out; p.a_out = g.a_out; p.bias = g.bias; p.gamma = g.gamma; p.beta = g.beta;
  p.bn_c = g.bn_c; p.alpha = g.alpha; p.act = g.act;
  p.wimg = img;
  { const char* dbg = getenv("UAD_HS_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
  UAD_REQUIRE(p.z_out || p.a_out, "conv_halo_ss: no output requested");

  // 5-D tensor map over the NHWC input: (channel [x column parity], W, row parity / 1, H, B); box = (32 ch, 10, 1, 18, 1)
  CUtensorMap tmap;
  cuuint64_t dims[5], strides[4];
  const cuuint64_t e = sizeof(float);
  if (form == 0) {
    dims[0] = 2ull * C; dims[1] = g.IW / 2; dims[2] = 2; dims[3] = g.IH / 2; dims[4] = g.B;
    strides[0] = 2ull * C * e; strides[1] = (cuuint64_t)g.IW * C * e; strides[2] = 2ull * g.IW * C * e;
    strides[3] = (cuuint64_t)g.IH * g.IW * C * e;
  } else {
    dims[0] = C; dims[1] = g.IW; dims[2] = 1; dims[3] = g.IH; dims[4] = g.B;
    strides[0] = (cuuint64_t)C * e; strides[1] = (cuuint64_t)g.IW * C * e; strides[2] = (cuuint64_t)g.IW * C * e;
    strides[3] = (cuuint64_t)g.IH * g.IW * C * e;
  }
  cuuint32_t box[5] = {32u, (cuuint32_t)kHW, 1u, (cuuint32_t)kHH, 1u};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(g.in), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UAD_REQUIRE(cr == CUDA_SUCCESS, "conv_halo_ss: cuTensorMapEncodeTiled failed (%d)", (int)cr);

  const int tiles = p.tiles_w * p.tiles_h * g.B;
  const bool g2 = p.Cblks >= 2;
  if (form == 0) {
    p.n_items = tiles;
    if (N == 32) return g2 ? launch_cfg<Cfg<0, 32, 1, 2>>(tmap, p, st) : launch_cfg<Cfg<0, 32, 1, 1>>(tmap, p, st);
    if (N == 64) return g2 ? launch_cfg<Cfg<0, 64, 1, 2>>(tmap, p, st) : launch_cfg<Cfg<0, 64, 1, 1>>(tmap, p, st);
    return g2 ? launch_cfg<Cfg<0, 128, 1, 2>>(tmap, p, st) : launch_cfg<Cfg<0, 128, 1, 1>>(tmap, p, st);
  }
  UAD_REQUIRE(C <= 128, "conv_halo_ss: stride-1 form supports at most 128 input channels");
  if (N == 32) { p.n_items = tiles; return launch_cfg<Cfg<1, 32, 4, 1>>(tmap, p, st); }
  if (N == 64) { p.n_items = tiles * 2; return launch_cfg<Cfg<1, 64, 2, 1>>(tmap, p, st); }
  p.n_items = tiles * 4;
  return launch_cfg<Cfg<1, 128, 1, 1>>(tmap, p, st);
}
