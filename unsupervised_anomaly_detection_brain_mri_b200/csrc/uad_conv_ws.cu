// Form W (filter gradients of the 5 x 5 stride-2 conv / transposed conv) on tcgen05 with BOTH operands read MN-major straight
// from TMA-written NHWC tiles - no transposition, no converter warps.  sm_100a only.
//
//   dW[(kh, kw)][cg][co] = sum_{b, r, s} G[b, 2r + kh - 1, 2s + kw - 1, cg] * O[b, r, s, co]        (uad_conv.cuh, WgradParams)
//
// Hardware facts it rests on (tools/ubench/operand_probe.cu E2-E5, tools/ubench/mn_probe.cu, tools/ubench/tma_swz_dump.cu;
// outputs under profiles/r2_*):
//   * a [pixel][32 channels] tile written by TMA with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B is a valid MN-major tf32 operand
//     (descriptor layout type 1): K = pixels (8 per instruction = two 4-row groups SBO = 512 bytes apart), M / N = channels in
//     groups of 32 that lie LBO bytes apart; the swizzle is a function of the absolute shared-memory address, so a descriptor may
//     start at ANY pixel row and its groups may overlap;
//   * kind::tf32 truncates fp32 operand words: the raw tile is its own 'hi' operand, lo = x - trunc(x) is one elementwise pass.
// Two-sided shift.  With P = one stride-2 parity plane (ph, pw) of G and window offsets (wh, ww) of a tap inside it,
//     D[(g, ci)][(j, co)] = sum_{pixels of a block} P[r + a_h, s + a_w0 + g][ci] * O[r + b_h(j), s][co]
// summed over all blocks of the (extended) pixel domain is the gradient of the tap with window offset (a_h - b_h(j), a_w0 + g):
// the M groups of ONE descriptor are the plane's halo shifted by one pixel each (LBO = 128 bytes: the 2 or 3 horizontal taps),
// the N groups of ONE descriptor are the O tile shifted by one row each (LBO = one tile row: the 2 or 3 vertical taps), so a
// parity plane's 4 / 6 / 9 taps are three MMAs per 8 pixels (raw x raw -> main; raw x lo, lo x raw -> corr).  Rows the shifted O
// reads above a block belong to its halo; the rows it never reaches at the bottom of an image are covered by extending the
// pixel domain by two rows (O is zero there by TMA out-of-bounds fill, G by SAME padding).
// One CTA = (pixel-block range, CTA type = the planes / row shifts whose accumulators share its 512 TMEM columns, 32-channel
// block of G); split-K partials are reduced by uad_launch_splitk_reduce (deterministic).
// Roles (256 threads): warp 0 TMA producer, warp 2 MMA issuer (everything unrolled at compile time), warp 3 TMEM, warps 4-7 lo
// pass, then the epilogue (TMEM -> partials).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "uad_conv.cuh"
#include "uad_tc_ptx.cuh"

namespace {
using namespace uadptx;

constexpr int kBH = 4;                                       // block rows
constexpr int kStages = 3;
constexpr int kThreads = 256;

template <int I, int E, class F>
__host__ __device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < E) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, E>(f);
  }
}

// one accumulator region of a CTA type: parity plane (ph, pw), row shifts [j0, j0 + nj)
struct Region { int ph, pw, j0, nj; };

template <int CO_, bool FAST_ = false>
struct WsCfg {
  static constexpr int CO = CO_;
  // FAST = UAD_MATH_TC_1XTF32: raw x raw only, no lo pass (the tensor core truncates the fp32 words it reads; activations written by
  // this mode's conv epilogues are already rounded to nearest tf32), no correction accumulators; the lo images' space = more stages
  static constexpr bool FAST = FAST_;
  static constexpr int STAGES = FAST ? 2 * kStages : kStages;
  static constexpr int NCB = CO / 32;                        // 32-channel blocks of O
  static constexpr int BW = CO == 128 ? 8 : 16;              // block width (pixels)
  static constexpr int HW = BW + 2;                          // plane halo width
  static constexpr int KQ = BW / 8;                          // K = 8 steps per block row
  static constexpr int NTYPES = CO == 32 ? 2 : (CO == 64 ? 4 : 6);
  static constexpr int MAXREG = 2;
  __host__ __device__ static constexpr int nreg(int t) { return CO == 32 ? 2 : 1; }
  __host__ __device__ static constexpr Region region(int t, int i) {
    if (CO == 32) return t == 0 ? (i == 0 ? Region{1, 1, 0, 3} : Region{0, 0, 0, 2}) : (i == 0 ? Region{1, 0, 0, 3} : Region{0, 1, 0, 2});
    if (CO == 64) return t == 0 ? Region{1, 1, 0, 3} : (t == 1 ? Region{1, 0, 0, 3} : (t == 2 ? Region{0, 1, 0, 2} : Region{0, 0, 0, 2}));
    return t == 0 ? Region{1, 1, 0, 2} : (t == 1 ? Region{1, 1, 2, 1} : (t == 2 ? Region{1, 0, 0, 2} : (t == 3 ? Region{1, 0, 2, 1}
           : (t == 4 ? Region{0, 1, 0, 2} : Region{0, 0, 0, 2}))));
  }
  // distinct planes a type loads (regions of one type never share a plane)
  __host__ __device__ static constexpr int nplanes(int t) { return nreg(t); }
  // TMEM column of region i of type t: [main (nj * CO) | corr (nj * CO)] per region, regions side by side
  __host__ __device__ static constexpr int col(int t, int i) { int c = 0; for (int k = 0; k < i; ++k) c += 2 * region(t, k).nj * CO; return c; }
  static constexpr uint32_t X_BYTES = kBH * HW * 128u;                         // one plane halo (raw)
  static constexpr uint32_t O_BYTES = (kBH + 2) * NCB * BW * 128u;             // the O tile with its two halo rows (raw)
  static constexpr uint32_t HALF = (CO == 32 ? 2 : 1) * X_BYTES + O_BYTES;     // raw images; the lo images follow at + HALF
  static constexpr uint32_t STAGE = (FAST ? 1 : 2) * HALF;
  static_assert(STAGE % 1024 == 0 && X_BYTES % 512 == 0 && O_BYTES % 512 == 0, "tile alignment");
};

struct WsParams {
  int B, Cg, ncb_g;             // images, channels of G, its 32-channel blocks
  int nbr, nbc;                 // block rows (extended domain) / block columns per image
  int nblocks, bpc;             // blocks in all images, blocks per chunk
  int Mp;                       // 25 * Cg
  int debug;                    // UAD_WGRAD_DEBUG: 1 = no lo pass, 2 = no MMAs
  float* partial;               // [chunks][Mp][CO]
};

template <class CF, int TYPE>
__device__ __forceinline__ void wgrad_ss_body(const CUtensorMap& tmap_g, const CUtensorMap& tmap_o, const WsParams& p, int cgb) {
  constexpr int CO = CF::CO, NCB = CF::NCB, BW = CF::BW, HW = CF::HW, KQ = CF::KQ, NREG = CF::nreg(TYPE);
  constexpr uint32_t X_BYTES = CF::X_BYTES, HALF = CF::HALF, STAGE = CF::STAGE;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int kSt = CF::STAGES;
  const uint32_t misc = smem_base + kSt * STAGE;
  const uint32_t bar_full = misc, bar_lo = misc + 64, bar_empty = misc + 128, bar_acc = misc + 192, tmem_slot = misc + 200;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blk0 = blockIdx.y * p.bpc;
  const int blk1 = blk0 + p.bpc < p.nblocks ? blk0 + p.bpc : p.nblocks;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSt; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_lo + 8 * i, 128); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 3) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      prefetch_tmap(&tmap_g);
      prefetch_tmap(&tmap_o);
      uint32_t s = 0, ph = 0;
      for (int blk = blk0; blk < blk1; ++blk) {
        const int bc = blk % p.nbc, br = (blk / p.nbc) % p.nbr, b = blk / (p.nbc * p.nbr);
        const int s0 = bc * BW, r0 = br * kBH;
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        mbar_expect_tx(bar_full + 8 * s, HALF);
        const uint32_t st = smem_base + s * STAGE;
#pragma unroll
        for (int i = 0; i < NREG; ++i) {
          const Region rg = CF::region(TYPE, i);
          // plane (ph, pw): rows r + a_h with a_h = -1 (ph = 1) / 0, columns s + a_w0 + g with a_w0 = -1 (pw = 1) / 0
          tma_load_5d(st + i * X_BYTES, &tmap_g, bar_full + 8 * s, rg.pw * p.Cg + cgb * 32, s0 - rg.pw, rg.ph, r0 - rg.ph, b);
        }
        tma_load_5d(st + NREG * X_BYTES, &tmap_o, bar_full + 8 * s, 0, s0, 0, r0 - 2, b);   // rows r0 - 2 .. r0 + 3
        if (++s == kSt) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ===================================================================== MMA issuer
    constexpr uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 4) << 24);
    // MN-major SWIZZLE_128B_BASE32B descriptors: A groups one pixel apart, B groups one (tile row, channel block) apart
    constexpr uint64_t da_bits = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
    constexpr uint64_t db_bits = ((uint64_t)((BW * 128) >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
    const bool no_mma = (p.debug & 2) != 0;
    uint32_t s = 0, ph = 0;
    for (int blk = blk0; blk < blk1; ++blk) {
      mbar_wait(bar_full + 8 * s, ph);
      if constexpr (!CF::FAST) mbar_wait(bar_lo + 8 * s, ph);
      tc_fence_after();
      const uint32_t st = smem_base + s * STAGE;
      const uint64_t xa = da_bits | (uint64_t)((st & 0x3FFFF) >> 4);
      const uint64_t ob = db_bits | (uint64_t)(((st + NREG * X_BYTES) & 0x3FFFF) >> 4);
      const uint32_t accum = blk > blk0 ? 1u : 0u;              // the CTA's first block overwrites the accumulators
      if (elect_one()) {
        if (!no_mma) {
          static_for<0, kBH>([&](auto RR) {
            constexpr int rr = decltype(RR)::value;
            static_for<0, KQ>([&](auto QQ) {
              constexpr int q = decltype(QQ)::value;
              static_for<0, NREG>([&](auto RI) {
                constexpr int i = decltype(RI)::value;
                constexpr Region rg = CF::region(TYPE, i);
                constexpr int nb = 2 + rg.ph;                   // row shifts of the plane; j <-> tile row rr + (3 - nb) + j
                constexpr uint32_t a_off = (uint32_t)(i * X_BYTES + (rr * HW + 8 * q) * 128) >> 4;
                constexpr uint32_t b_off = (uint32_t)(((rr + 3 - nb + rg.j0) * NCB * BW + 8 * q) * 128) >> 4;
                constexpr uint32_t idesc = idesc_base | ((uint32_t)((rg.nj * CO) >> 3) << 17);
                constexpr uint32_t main_c = CF::col(TYPE, i), corr_c = main_c + rg.nj * CO;
                const uint32_t acc = (rr == 0 && q == 0) ? accum : 1u;
                mma_tf32_ss(tmem_base + main_c, xa + a_off, ob + b_off, idesc, acc);                              // raw x raw
                if constexpr (!CF::FAST) {
                  mma_tf32_ss(tmem_base + corr_c, xa + a_off, ob + b_off + (HALF >> 4), idesc, acc);              // raw x lo
                  mma_tf32_ss(tmem_base + corr_c, xa + a_off + (HALF >> 4), ob + b_off, idesc, 1u);               // lo x raw
                }
              });
            });
          });
        }
        tc_commit(bar_empty + 8 * s);
        if (blk == blk1 - 1) tc_commit(bar_acc);
      }
      __syncwarp();
      if (++s == kSt) { s = 0; ph ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================================================================== lo pass: second half of the stage = first half - trunc
    const int tid = threadIdx.x - 128;
    const bool skip = (p.debug & 1) != 0;
    uint32_t s = 0, ph = 0;
    for (int blk = blk0; blk < (CF::FAST ? blk0 : blk1); ++blk) {
      mbar_wait(bar_full + 8 * s, ph);
      if (!skip) {
        float4* raw = reinterpret_cast<float4*>(smem_gen + s * STAGE);
        float4* lo = reinterpret_cast<float4*>(smem_gen + s * STAGE + HALF);
#pragma unroll 4
        for (int i = tid; i < (int)(HALF / 16); i += 128) {
          const float4 v = raw[i];
          float4 h, l;
          h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
          l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
          raw[i] = h;                                           // hi = round-to-nearest tf32 (in place), lo = x - hi
          lo[i] = l;
        }
        fence_proxy_async();
      }
      mbar_arrive(bar_lo + 8 * s);
      if (++s == kSt) { s = 0; ph ^= 1; }
    }
    // ===================================================================== epilogue: TMEM lane group g = horizontal tap g of each region
    if (blk1 > blk0) {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
      const int g = warp & 3;                                   // this warp reads TMEM lanes 32 g .. 32 g + 31 (row = channel ci = lane)
      float* part = p.partial + (size_t)blockIdx.y * p.Mp * CO;
      static_for<0, NREG>([&](auto RI) {
        constexpr int i = decltype(RI)::value;
        constexpr Region rg = CF::region(TYPE, i);
        constexpr int na = 2 + rg.pw;
        constexpr uint32_t main_c = CF::col(TYPE, i), corr_c = main_c + rg.nj * CO;
        if (g < na) {
          const int kw = 2 * g + (rg.pw ? 0 : 1);
#pragma unroll 1
          for (int jj = 0; jj < rg.nj; ++jj) {
            const int j = rg.j0 + jj;
            const int kh = (rg.ph ? 4 : 3) - 2 * j;             // window offset wh = 1 - j
            float* row = part + ((size_t)((kh * 5 + kw) * p.Cg + cgb * 32 + lane)) * CO;
#pragma unroll 1
            for (int c0 = 0; c0 < CO; c0 += 32) {
              uint32_t v[32], u[32];
              const uint32_t lane_addr = tmem_base + ((uint32_t)(g * 32) << 16);
              tmem_ld32(lane_addr + main_c + jj * CO + c0, v);
              if constexpr (CF::FAST) {
#pragma unroll
                for (int e = 0; e < 32; ++e) u[e] = 0u;
              } else {
                tmem_ld32(lane_addr + corr_c + jj * CO + c0, u);
              }
              tmem_wait_ld();
#pragma unroll
              for (int e = 0; e < 32; e += 4)
                *reinterpret_cast<float4*>(row + c0 + e) =
                    make_float4(__uint_as_float(v[e]) + __uint_as_float(u[e]), __uint_as_float(v[e + 1]) + __uint_as_float(u[e + 1]),
                                __uint_as_float(v[e + 2]) + __uint_as_float(u[e + 2]), __uint_as_float(v[e + 3]) + __uint_as_float(u[e + 3]));
            }
          }
        }
      });
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// grid = (CTA type x 32-channel block of G, pixel chunks): the CTAs that share a chunk's O tiles are scheduled together (L2 reuse)
template <class CF>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_ss(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ WsParams p) {
  const int type = blockIdx.x / p.ncb_g, cgb = blockIdx.x % p.ncb_g;
  static_for<0, CF::NTYPES>([&](auto T) {
    if (type == decltype(T)::value) wgrad_ss_body<CF, decltype(T)::value>(tmap_g, tmap_o, p, cgb);
  });
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn ws_encode_fn() {
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

struct WsPlan { int nbr, nbc, nblocks, bpc, nchunks, ntypes, BW; };

WsPlan ws_plan(int Cg, int Co, int B, int MH, int MW) {
  WsPlan pl;
  pl.BW = Co == 128 ? 8 : 16;
  pl.ntypes = Co == 32 ? 2 : (Co == 64 ? 4 : 6);
  pl.nbr = (MH + 2 + kBH - 1) / kBH;                          // two extra rows: the shifted O reaches every row of the image
  pl.nbc = MW / pl.BW;
  pl.nblocks = B * pl.nbr * pl.nbc;
  // about four CTAs per SM over all (type, channel block) columns of the grid, and no accumulator deeper than ~4096 pixels
  // (the tensor core adds into its fp32 accumulator with truncation)
  // CTAs per SM over the whole grid: four at Co = 32 (the 256^2 layers: tail balance wins), three above (fewer split-K partials to
  // reduce wins; measured at the VAE-256 shapes, profiles/r2_final_wgrad_waves.txt).  UAD_WS_WAVES overrides (developer switch).
  static int waves_env = -1;
  if (waves_env < 0) { const char* e = getenv("UAD_WS_WAVES"); waves_env = e ? atoi(e) : 0; }
  const int waves = waves_env > 0 ? waves_env : (Co == 32 ? 4 : 3);
  int target = (waves * UAD_NUM_SMS) / (pl.ntypes * (Cg / 32));
  if (target < 1) target = 1;
  const int px_per_block = kBH * pl.BW;
  static int depth = -1;                                       // developer switch UAD_WS_DEPTH: pixels per accumulator
  if (depth < 0) { const char* e = getenv("UAD_WS_DEPTH"); depth = e ? atoi(e) : 4096; }
  const int min_chunks = uad_cdiv((long long)pl.nblocks * px_per_block, depth);
  if (target < min_chunks) target = min_chunks;
  if (target > pl.nblocks) target = pl.nblocks;
  pl.bpc = uad_cdiv(pl.nblocks, target);
  pl.nchunks = uad_cdiv(pl.nblocks, pl.bpc);
  return pl;
}

template <class CF>
int launch_all_types(const CUtensorMap& tg, const CUtensorMap& to, const WsParams& p, int nchunks, cudaStream_t st) {
  const size_t smem = 1024 + CF::STAGES * CF::STAGE + 256;
  static bool attr = false;
  if (!attr) {
    UAD_CUDA(cudaFuncSetAttribute(wgrad_ss<CF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  UAD_REQUIRE(smem <= 227 * 1024, "wgrad_ss: shared-memory budget exceeded");
  dim3 grid(CF::NTYPES * p.ncb_g, nchunks);
  wgrad_ss<CF><<<grid, kThreads, smem, st>>>(tg, to, p);
  UAD_LAUNCH_CHECK("wgrad_ss");
  return 0;
}

}  // namespace

int uad_ws_wgrad_supported(int Cg, int Co, int lgMH, int lgMW) {
  if (Cg % 32 != 0 || Cg < 32) return 0;
  if (!(Co == 32 || Co == 64 || Co == 128)) return 0;
  const int BW = Co == 128 ? 8 : 16;
  if ((1 << lgMW) < BW || lgMH < 1) return 0;
  return 1;
}

size_t uad_ws_wgrad_ws_bytes(int Cg, int Co, int B, int MH, int MW) {
  const WsPlan pl = ws_plan(Cg, Co, B, MH, MW);
  return (size_t)pl.nchunks * 25 * Cg * Co * sizeof(float) + 1024;
}

int uad_launch_wgrad_ss(const WgradParams& w, float* out, int accumulate, bool fast, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int Cg = w.Cg, Co = w.Co;
  UAD_REQUIRE(w.sh == 2 && w.taps.n == 25, "wgrad_ss: only the 5x5 stride-2 gather is implemented");
  for (int t = 0; t < 25; ++t)
    UAD_REQUIRE(w.taps.dh[t] == t / 5 - 1 && w.taps.dw[t] == t % 5 - 1, "wgrad_ss: unexpected tap table");
  EncodeTiledFn encode = ws_encode_fn();
  UAD_REQUIRE(encode != nullptr, "wgrad_ss: cuTensorMapEncodeTiled entry point unavailable");
  const int MH = 1 << w.lgMH, MW = 1 << w.lgMW;
  const WsPlan pl = ws_plan(Cg, Co, w.B, MH, MW);
  const size_t need = (size_t)pl.nchunks * w.Mp * Co * sizeof(float);
  UAD_REQUIRE(ws && ws_bytes >= need, "wgrad_ss: workspace too small (%zu < %zu)", ws_bytes, need);
  UAD_REQUIRE(((uintptr_t)w.g % 16) == 0 && ((uintptr_t)w.o % 16) == 0 && ((uintptr_t)ws % 16) == 0, "wgrad_ss: unaligned buffers");

  WsParams p;
  memset(&p, 0, sizeof(p));
  p.B = w.B; p.Cg = Cg; p.ncb_g = Cg / 32;
  p.nbr = pl.nbr; p.nbc = pl.nbc; p.nblocks = pl.nblocks; p.bpc = pl.bpc; p.Mp = w.Mp;
  p.partial = reinterpret_cast<float*>(ws);
  { const char* dbg = getenv("UAD_WGRAD_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }

  const cuuint64_t e = sizeof(float);
  CUtensorMap tmap_g, tmap_o;
  {   // G [B, GH, GW, Cg] viewed as (2 * Cg [column parity x channel], GW / 2, 2 [row parity], GH / 2, B); box = one plane's block halo
    cuuint64_t dims[5] = {2ull * Cg, (cuuint64_t)w.GW / 2, 2, (cuuint64_t)w.GH / 2, (cuuint64_t)w.B};
    cuuint64_t strides[4] = {2ull * Cg * e, (cuuint64_t)w.GW * Cg * e, 2ull * w.GW * Cg * e, (cuuint64_t)w.GH * w.GW * Cg * e};
    cuuint32_t box[5] = {32, (cuuint32_t)(pl.BW + 2), 1, kBH, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult cr = encode(&tmap_g, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(w.g), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UAD_REQUIRE(cr == CUDA_SUCCESS, "wgrad_ss: cuTensorMapEncodeTiled(g) failed (%d)", (int)cr);
  }
  {   // O [B, MH, MW, Co] viewed as (32 [channel in block], MW, Co / 32 [block], MH, B): shared memory order [row][block][pixel][32]
    cuuint64_t dims[5] = {32, (cuuint64_t)MW, (cuuint64_t)(Co / 32), (cuuint64_t)MH, (cuuint64_t)w.B};
    cuuint64_t strides[4] = {(cuuint64_t)Co * e, 128, (cuuint64_t)MW * Co * e, (cuuint64_t)MH * MW * Co * e};
    cuuint32_t box[5] = {32, (cuuint32_t)pl.BW, (cuuint32_t)(Co / 32), kBH + 2, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult cr = encode(&tmap_o, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(w.o), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UAD_REQUIRE(cr == CUDA_SUCCESS, "wgrad_ss: cuTensorMapEncodeTiled(o) failed (%d)", (int)cr);
  }
  int rc;
  if (fast) {   // UAD_MATH_TC_1XTF32
    if (Co == 32) rc = launch_all_types<WsCfg<32, true>>(tmap_g, tmap_o, p, pl.nchunks, st);
    else if (Co == 64) rc = launch_all_types<WsCfg<64, true>>(tmap_g, tmap_o, p, pl.nchunks, st);
    else rc = launch_all_types<WsCfg<128, true>>(tmap_g, tmap_o, p, pl.nchunks, st);
  } else if (Co == 32) rc = launch_all_types<WsCfg<32>>(tmap_g, tmap_o, p, pl.nchunks, st);
  else if (Co == 64) rc = launch_all_types<WsCfg<64>>(tmap_g, tmap_o, p, pl.nchunks, st);
  else rc = launch_all_types<WsCfg<128>>(tmap_g, tmap_o, p, pl.nchunks, st);
  if (rc) return rc;
  return uad_launch_splitk_reduce(p.partial, pl.nchunks, (size_t)w.Mp * Co, out, accumulate, st);
}
