// Halo-resident SS-form implicit GEMM on tcgen05 for the conv family (Form F and Form T of uad_conv.cuh), sm_100a only.
//
// Round-2 replacement of the converter-warp kernels (uad_conv_tc.cu), built on three hardware answers of
// tools/ubench/operand_probe.cu (profiles/r2_operand_probe.txt):
//   E1      kind::tf32 TRUNCATES fp32 operand words  -> a TMA-loaded fp32 tile IS the tf32 'hi' operand; only lo = x - trunc(x)
//           has to be produced (one elementwise pass over the tile, same byte offsets)
//   E6/E8   the 128-byte swizzle is a function of the absolute shared-memory address -> a K-major descriptor may start at ANY
//           128-byte row of a TMA-written tile and its 8-row groups may be any number of rows apart (SBO = halo pitch)
//   (r2a)   shared-memory bandwidth is shared between TMA fills and MMA operand reads, and every per-k-block barrier hop
//           costs 150-200 cycles in a loaded kernel -> fill each input element ONCE per tile and hop per tile, not per tap
//
// Tile = 16 x 8 pixels of the M-grid (M = 128 rows; one 8-pixel tile row = one 8-row descriptor group).  Per 32-channel block
// the (16+2) x (8+2) pixel halo of the tile (stride-1 form) or of one stride-2 parity plane (strided form) is TMA-loaded ONCE
// as 180 rows of 128 bytes (SWIZZLE_128B); every filter tap that reads from it is a descriptor START ADDRESS
// (halo + ((dh+1) * 10 + (dw+1)) * 128 bytes, SBO = 1280).  fp32 parity = 3xTF32 with the paired-B trick:
//     acc[main | corr] (+)= A_raw . [B_hi ; B_lo]^T          one MMA of width 2N  (hi*hi -> main, hi*lo -> corr)
//     acc[corr]         +=  A_lo  .  B_hi^T                  one MMA of width N
// Roles (384 threads, one persistent CTA per SM):
//   warp 0      halo TMA producer                                   -> h_full[s]
//   warps 4-7   lo pass: lo tile = raw - trunc(raw)                 -> h_lo[s]            (once per halo, NOT per tap)
//   warp 1      weight-image producer (one bulk copy per k-block)   -> w_full[t]
//   warp 2      MMA issuer: 8 x tcgen05.mma per k-block, commits    -> w_empty[t], h_empty[s], acc_full[b]
//   warp 3      TMEM allocation, epilogue constants
//   warps 8-11  epilogue: tcgen05.ld, sum of the accumulators (RN), bias / frozen-BN / activation, smem transpose, coalesced
//               row stores                                          -> acc_empty[b]
// The tensor core adds into its fp32 accumulator with truncation (round 1, measured): deep reductions (K > 1152) deal the
// k-blocks round-robin over G = 2 accumulator pairs; the epilogue sums them in registers.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

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

struct HsKb { unsigned short a_off16; unsigned char wt, cls; };      // window start inside the halo (16-byte units), weight tap, class
struct HsUnit { int n, c_plane, h_plane; HsKb kb[UAD_MAX_TAPS]; };  // one halo load and the k-blocks that read from it

struct HsParams {
  int tiles_w, tiles_h;
  int B, C, Cblks, N;
  int OH, OW, osh;
  int h_stages, w_stages;
  int G;               // accumulator pairs per class (round-robin over k-blocks)
  int CG;              // output-parity classes per work item (stride-1 form), 1 for the strided form
  int nvar;            // work-item variants (class groups): item = tile * nvar + variant
  int n_units;         // halo units per (variant, channel block): 4 parity planes (strided form) or 1
  int acc_bufs;        // accumulator sets in TMEM
  int n_items;
  int debug;           // developer timing switches (UAD_HS_DEBUG): 1 = no lo pass, 2 = no MMAs, 4 = no global stores
  float* z_out;
  float* a_out;
  const float* bias;
  const float* gamma;
  const float* beta;
  float bn_c, alpha;
  int act;
  const float* wimg;   // [k*k][Cblks][2][N][32] pre-swizzled {hi, lo} weight images
  signed char oh0[4][4], ow0[4][4];      // [variant][class in group]
  HsUnit units[4];     // [variant * n_units + unit]
};

__global__ void __launch_bounds__(kThreads, 1)
conv_halo_ss(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ HsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int N = p.N;
  const uint32_t w_bytes = 2u * N * 128u;
  const int HS = p.h_stages, WS = p.w_stages;
  const uint32_t w_base = smem_base + HS * kHaloStage;
  const uint32_t misc_off = HS * kHaloStage + WS * w_bytes;
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
  const int acc_cols = p.CG * p.G * 2 * N;              // TMEM columns of one accumulator set
  const int units_per_item = p.Cblks * p.n_units;

  if (threadIdx.x == 0) {
    for (int i = 0; i < HS; ++i) { mbar_init(bar_hfull + 8 * i, 1); mbar_init(bar_hlo + 8 * i, 128); mbar_init(bar_hempty + 8 * i, 1); }
    for (int i = 0; i < WS; ++i) { mbar_init(bar_wfull + 8 * i, 1); mbar_init(bar_wempty + 8 * i, 1); }
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
      int s = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int var = item % p.nvar, tile = item / p.nvar;
        const int twi = tile % p.tiles_w, thi = (tile / p.tiles_w) % p.tiles_h, b = tile / (p.tiles_w * p.tiles_h);
        const int s0 = twi * kTW - 1, r0 = thi * kTH - 1;     // halo origin (the zero fill outside the tensor == SAME padding)
        for (int cb = 0; cb < p.Cblks; ++cb)
          for (int u = 0; u < p.n_units; ++u) {
            const HsUnit& un = p.units[var * p.n_units + u];
            mbar_wait(bar_hempty + 8 * s, ph ^ 1);
            mbar_expect_tx(bar_hfull + 8 * s, kHaloBytes);
            tma_load_5d(smem_base + s * kHaloStage, &tmap, bar_hfull + 8 * s, un.c_plane + cb * 32, s0, un.h_plane, r0, b);
            if (++s == HS) { s = 0; ph ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== weight-image producer
    if (lane == 0) {
      int t = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int var = item % p.nvar;
        for (int cb = 0; cb < p.Cblks; ++cb)
          for (int u = 0; u < p.n_units; ++u) {
            const HsUnit& un = p.units[var * p.n_units + u];
            for (int i = 0; i < un.n; ++i) {
              mbar_wait(bar_wempty + 8 * t, ph ^ 1);
              mbar_expect_tx(bar_wfull + 8 * t, w_bytes);
              const float* src = p.wimg + ((size_t)((int)un.kb[i].wt * p.Cblks + cb)) * 2 * N * 32;
              bulk_load(w_base + t * w_bytes, src, w_bytes, bar_wfull + 8 * t);
              if (++t == WS) { t = 0; ph ^= 1; }
            }
          }
      }
    }
  } else if (warp == 2) {
    // ===================================================================== MMA issuer (whole warp converged, one elected lane issues)
    const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idescN = idesc_base | ((uint32_t)(N >> 3) << 17);
    const uint32_t idesc2N = idesc_base | ((uint32_t)((2 * N) >> 3) << 17);
    const uint64_t adesc0 = make_kmajor_sw128_desc(smem_base, kSbo);           // raw tile of halo stage 0
    const uint64_t bdesc0 = make_kmajor_sw128_desc(w_base, 1024u);             // weight image of stage 0: N hi rows, N lo rows
    const uint32_t hstage_units = kHaloStage >> 4, lo_units = kHaloSlot >> 4, w_units = w_bytes >> 4;
    const bool no_mma = (p.debug & 2) != 0;
    int s = 0, t = 0, buf = 0;
    uint32_t phs = 0, pht = 0, phb = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int var = item % p.nvar;
      mbar_wait(bar_accempty + 8 * buf, phb ^ 1);               // the epilogue has drained this accumulator set
      tc_fence_after();
      const uint32_t acc0 = tmem_base + buf * acc_cols;
      uint32_t inited = 0, rr = 0;                              // per (class, g) "holds a partial sum" bits; per-class round-robin bit
      int units_left = units_per_item;
      for (int cb = 0; cb < p.Cblks; ++cb)
        for (int u = 0; u < p.n_units; ++u) {
          const HsUnit& un = p.units[var * p.n_units + u];
          mbar_wait(bar_hfull + 8 * s, phs);                    // the issuing thread observes the TMA completion itself
          mbar_wait(bar_hlo + 8 * s, phs);                      // ... and the lo tile written from it
          tc_fence_after();
          const uint64_t a_raw0 = adesc0 + (uint64_t)(s * hstage_units);
          --units_left;
          for (int i = 0; i < un.n; ++i) {
            mbar_wait(bar_wfull + 8 * t, pht);
            tc_fence_after();
            const HsKb kb = un.kb[i];
            const int g = (p.G == 2) ? (int)((rr >> kb.cls) & 1u) : 0;
            const int slot = kb.cls * p.G + g;
            const uint32_t first = (inited >> slot) & 1u;       // 0 -> the first MMA overwrites (zero-initialises) main | corr
            inited |= 1u << slot;
            rr ^= 1u << kb.cls;
            const uint32_t d_pair = acc0 + slot * 2 * N;
            const uint64_t a_raw = a_raw0 + kb.a_off16;
            const uint64_t a_lo = a_raw + lo_units;
            const uint64_t b_img = bdesc0 + (uint64_t)(t * w_units);
            if (elect_one()) {
              if (!no_mma) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {                   // K = 8 tf32 per instruction = 32 bytes along the 128-byte row
                  mma_tf32_ss(d_pair, a_raw + 2 * j, b_img + 2 * j, idesc2N, first | (j != 0));
                  mma_tf32_ss(d_pair + N, a_lo + 2 * j, b_img + 2 * j, idescN, 1u);
                }
              }
              tc_commit(bar_wempty + 8 * t);                    // weight stage reusable once these MMAs retire
              if (i == un.n - 1) {
                tc_commit(bar_hempty + 8 * s);                  // so is the halo stage after the unit's last k-block
                if (units_left == 0) tc_commit(bar_accfull + 8 * buf);
              }
            }
            __syncwarp();
            if (++t == WS) { t = 0; pht ^= 1; }
          }
          if (++s == HS) { s = 0; phs ^= 1; }
        }
      if (++buf == p.acc_bufs) { buf = 0; phb ^= 1; }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================================================================== lo pass (once per halo tile, elementwise, same byte offsets)
    const int tid = threadIdx.x - 128;
    const bool skip = (p.debug & 1) != 0;
    int s = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      for (int uu = 0; uu < units_per_item; ++uu) {
        mbar_wait(bar_hfull + 8 * s, ph);
        if (!skip) {
          const float4* raw = reinterpret_cast<const float4*>(smem_gen + s * kHaloStage);
          float4* lo = reinterpret_cast<float4*>(smem_gen + s * kHaloStage + kHaloSlot);
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
        mbar_arrive(bar_hlo + 8 * s);
        if (++s == HS) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 8) {
    // ===================================================================== epilogue
    const int row = threadIdx.x - 256;                          // tile row == TMEM lane
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float* stg = stg_base + (size_t)q * 32 * 36;                // this warp's 32 x (32+4) staging rows
    const int nchunks = N >> 5;
    const bool no_store = (p.debug & 4) != 0;
    const int tw = row & (kTW - 1), th = row >> 3;
    int buf = 0;
    uint32_t phb = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int var = item % p.nvar, tile = item / p.nvar;
      const int twi = tile % p.tiles_w, thi = (tile / p.tiles_w) % p.tiles_h, b = tile / (p.tiles_w * p.tiles_h);
      const int s0 = twi * kTW, r0 = thi * kTH;
      const uint32_t acc_base = lane_base + buf * acc_cols;
      mbar_wait(bar_accfull + 8 * buf, phb);
      tc_fence_after();
      const int cq = (lane & 7) * 4;
      for (int cls = 0; cls < p.CG; ++cls) {
        const long long my_off =
            (((long long)b * p.OH + ((r0 + th) * p.osh + p.oh0[var][cls])) * p.OW + ((s0 + tw) * p.osh + p.ow0[var][cls])) * (long long)N;
        long long offs[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) offs[it] = __shfl_sync(0xffffffffu, my_off, it * 4 + (lane >> 3));
        const uint32_t cls_base = acc_base + cls * p.G * 2 * N;
        for (int c = 0; c < nchunks; ++c) {
          const int c0 = c * 32;
          uint32_t v[32], u[32];
          tmem_ld32(cls_base + c0, v);                          // pair 0: main columns [0, N), corr columns [N, 2N)
          tmem_ld32(cls_base + N + c0, u);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
          if (p.G == 2) {
            uint32_t w2[32];
            tmem_ld32(cls_base + 2 * N + c0, u);
            tmem_ld32(cls_base + 3 * N + c0, w2);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = __float_as_uint(__uint_as_float(v[j]) + (__uint_as_float(u[j]) + __uint_as_float(w2[j])));
          }
          if (cls + 1 == p.CG && c + 1 == nchunks) {            // last TMEM read of this thread for the item: hand the set back
            tc_fence_before();
            mbar_arrive(bar_accempty + 8 * buf);
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
                o[e] = pass == 0 ? z : uad_act(epi[N + n] * z + epi[2 * N + n], p.act, p.alpha);
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
      if (++buf == p.acc_bufs) { buf = 0; phb ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// raw weights -> per (tap, 32-channel block): {hi, lo} images of [N rows][32 k] fp32 in the SWIZZLE_128B byte order the
// descriptor expects (16-byte chunk index XOR (row & 7)).  transposed=false: raw[t][c][n]; true: raw[t][n][c].
__global__ void hs_weight_image_kernel(const float* __restrict__ w, float* __restrict__ img, int taps, int C, int N, int transposed) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)taps * C * N;
  if (i >= total) return;
  const int k = i % 32;
  const int n = (i / 32) % N;
  const int cb = (i / ((size_t)32 * N)) % (C / 32);
  const int t = i / ((size_t)C * N);
  const int c = cb * 32 + k;
  const float v = transposed ? w[((size_t)t * N + n) * C + c] : w[((size_t)t * C + c) * N + n];
  const uint32_t h = __float_as_uint(v) & 0xffffe000u;
  const float lo = v - __uint_as_float(h);
  const size_t base = ((size_t)(t * (C / 32) + cb)) * 2 * N * 32;
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

}  // namespace

int uad_hs_gather_supported(int Cin, int N, int lgMH, int lgMW) {
  if (Cin % 32 != 0 || Cin < 32) return 0;
  if (!(N == 32 || N == 64 || N == 128)) return 0;
  if (lgMH < 4 || lgMW < 3) return 0;          // M-grid at least 16 x 8: one tile never spans two images
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
  UAD_REQUIRE(ws && ws_bytes >= need, "conv_halo_ss: workspace too small (%zu < %zu)", ws_bytes, need);
  UAD_REQUIRE(((uintptr_t)ws % 128) == 0 && ((uintptr_t)g.in % 16) == 0, "conv_halo_ss: unaligned buffers");
  UAD_REQUIRE(nclasses == 1 || nclasses == 4, "conv_halo_ss: %d tap classes", nclasses);
  UAD_REQUIRE((g.sh == 2) == (nclasses == 1), "conv_halo_ss: strided form has one class, stride-1 form four");
  EncodeTiledFn encode = hs_encode_fn();
  UAD_REQUIRE(encode != nullptr, "conv_halo_ss: cuTensorMapEncodeTiled entry point unavailable");

  float* img = reinterpret_cast<float*>(ws);
  {
    const size_t total = (size_t)ksize * ksize * C * N;
    hs_weight_image_kernel<<<uad_cdiv(total, 256), 256, 0, st>>>(w_raw, img, ksize * ksize, C, N, weights_transposed ? 1 : 0);
    UAD_LAUNCH_CHECK("hs_weight_image");
  }

  HsParams p;
  memset(&p, 0, sizeof(p));
  const int MW = 1 << g.lgMW, MH = 1 << g.lgMH;
  p.tiles_w = MW / kTW;
  p.tiles_h = MH / kTH;
  p.B = g.B; p.C = C; p.Cblks = C / 32; p.N = N;
  p.OH = g.OH; p.OW = g.OW; p.osh = g.osh;
  int max_taps = 0;
  for (int c = 0; c < nclasses; ++c) max_taps = g.taps[c].n > max_taps ? g.taps[c].n : max_taps;
  p.G = (max_taps * C > 1152) ? 2 : 1;
  if (g.sh == 2) {
    // strided form: four parity planes of the input, each loaded once per channel block; tap (dh, dw) reads plane (dh & 1, dw & 1)
    // at window offset (dh >> 1, dw >> 1) (floor)
    p.CG = 1; p.nvar = 1; p.n_units = 4;
    const TapSet& ts = g.taps[0];
    for (int u = 0; u < 4; ++u) {
      HsUnit& un = p.units[u];
      const int ph = u >> 1, pw = u & 1;
      un.n = 0; un.c_plane = pw * C; un.h_plane = ph;
      for (int t = 0; t < ts.n; ++t) {
        const int dh = ts.dh[t], dw = ts.dw[t];
        if ((dh & 1) != ph || (dw & 1) != pw) continue;
        const int wh = (dh >> 1) + 1, wwd = (dw >> 1) + 1;
        UAD_REQUIRE(wh >= 0 && wh <= 2 && wwd >= 0 && wwd <= 2, "conv_halo_ss: tap (%d, %d) outside the halo", dh, dw);
        HsKb& kb = un.kb[un.n++];
        kb.a_off16 = (unsigned short)((wh * kHW + wwd) * 8);
        kb.wt = (unsigned char)ts.wt[t];
        kb.cls = 0;
      }
    }
  } else {
    // stride-1 form: one halo per channel block serves every class of the item
    p.CG = N == 32 ? 4 : (N == 64 ? 2 : 1);
    p.nvar = 4 / p.CG; p.n_units = 1;
    for (int v = 0; v < p.nvar; ++v) {
      HsUnit& un = p.units[v];
      un.n = 0; un.c_plane = 0; un.h_plane = 0;
      for (int cl = 0; cl < p.CG; ++cl) {
        const TapSet& ts = g.taps[v * p.CG + cl];
        p.oh0[v][cl] = (signed char)ts.oh0;
        p.ow0[v][cl] = (signed char)ts.ow0;
        for (int t = 0; t < ts.n; ++t) {
          const int wh = ts.dh[t] + 1, wwd = ts.dw[t] + 1;
          UAD_REQUIRE(wh >= 0 && wh <= 2 && wwd >= 0 && wwd <= 2, "conv_halo_ss: tap (%d, %d) outside the halo", ts.dh[t], ts.dw[t]);
          UAD_REQUIRE(un.n < UAD_MAX_TAPS, "conv_halo_ss: too many k-blocks per unit");
          HsKb& kb = un.kb[un.n++];
          kb.a_off16 = (unsigned short)((wh * kHW + wwd) * 8);
          kb.wt = (unsigned char)ts.wt[t];
          kb.cls = (unsigned char)cl;
        }
      }
    }
  }
  const int acc_cols = p.CG * p.G * 2 * N;
  UAD_REQUIRE(acc_cols <= 512, "conv_halo_ss: TMEM budget exceeded (%d columns)", acc_cols);
  p.acc_bufs = (2 * acc_cols <= 512) ? 2 : 1;
  p.n_items = p.tiles_w * p.tiles_h * g.B * p.nvar;
  p.z_out = g.z_out; p.a_out = g.a_out; p.bias = g.bias; p.gamma = g.gamma; p.beta = g.beta;
  p.bn_c = g.bn_c; p.alpha = g.alpha; p.act = g.act;
  p.wimg = img;
  { const char* dbg = getenv("UAD_HS_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
  UAD_REQUIRE(p.z_out || p.a_out, "conv_halo_ss: no output requested");

  // shared-memory budget: halo stages (raw + lo), weight stages, barriers / constants / staging
  const size_t w_bytes = 2u * N * 128u;
  const size_t tail = 320 + 3 * N * sizeof(float) + 4 * 32 * 36 * sizeof(float) + 64;
  const size_t budget = 227 * 1024 - 1024 - tail;
  p.h_stages = 3;
  p.w_stages = (int)((budget - p.h_stages * kHaloStage) / w_bytes);
  if (p.w_stages < 4) {
    p.h_stages = 2;
    p.w_stages = (int)((budget - p.h_stages * kHaloStage) / w_bytes);
  }
  if (p.w_stages > 8) p.w_stages = 8;
  UAD_REQUIRE(p.w_stages >= 2, "conv_halo_ss: shared-memory budget exceeded");
  const size_t smem = 1024 + p.h_stages * kHaloStage + p.w_stages * w_bytes + tail;

  // 5-D tensor map over the NHWC input: (channel [x column parity], W, row parity / 1, H, B); box = (32 ch, 10, 1, 18, 1)
  CUtensorMap tmap;
  cuuint64_t dims[5], strides[4];
  const cuuint64_t e = sizeof(float);
  if (g.sh == 2) {
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

  static bool attr = false;
  if (!attr) {
    UAD_CUDA(cudaFuncSetAttribute(conv_halo_ss, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  const int grid = p.n_items < UAD_NUM_SMS ? p.n_items : UAD_NUM_SMS;
  conv_halo_ss<<<grid, kThreads, smem, st>>>(tmap, p);
  UAD_LAUNCH_CHECK("conv_halo_ss");
  return 0;
}
