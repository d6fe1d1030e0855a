// Index arithmetic of the plane-resident Form-W kernel (wgrad_tc2, csrc/uad_conv_tc.cu).  Host/device so that
// tests/test_wgrad_tiles.py can run it with g++ against the tap table of the shipped kernel.
//
// A 4 x 8 block of coarse pixels gathers, for the 5 x 5 stride-2 filter, from a 6 x 10 halo of each of the four stride-2 parity
// planes (ph, pw) of the fine tensor: tap (kh, kw) has fine offset (dh, dw) = (kh - 1, kw - 1), lives in plane (dh & 1, dw & 1)
// and reads the 4 x 8 window that starts at halo row oh = (dh >> 1) + 1, column ow = (dw >> 1) + 1 (arithmetic shifts).
// The kernel keeps each plane's halo in tensor memory (lane = 32 * plane + channel, column = 10 * halo row + halo column) and
// indexes its accumulator tiles by the WINDOW (oh, ow) in {0,1,2}^2: lane group `plane` of tile (oh, ow) accumulates tap
// (2 * oh + ph - 1, 2 * ow + pw - 1) when that is a filter tap (4 full, 4 half, 1 quarter tile = 25 taps).
#pragma once
#if defined(__CUDACC__)
#define UAD_WT_HD __host__ __device__ __forceinline__
#else
#define UAD_WT_HD static inline
#endif

#define UAD_WT_HALO_W 10          // halo columns per plane (8 + 2)
#define UAD_WT_HALO_H 6           // halo rows per plane (4 + 2)
#define UAD_WT_TILES 9

// filter tap index kh * 5 + kw accumulated by lane group `plane` = (ph << 1) | pw of window tile `tile` = oh * 3 + ow; -1 if none
UAD_WT_HD int uad_wt_tap(int tile, int plane) {
  const int oh = tile / 3, ow = tile % 3;
  const int kh = 2 * oh + (plane >> 1) - 1, kw = 2 * ow + (plane & 1) - 1;
  return (kh < 0 || kh > 4 || kw < 0 || kw > 4) ? -1 : kh * 5 + kw;
}

// first tensor-memory column (relative to the plane copy) of the K = 8 window of block row r (0..3) for window tile `tile`
UAD_WT_HD int uad_wt_a_column(int tile, int r) { return (tile / 3 + r) * UAD_WT_HALO_W + tile % 3; }
