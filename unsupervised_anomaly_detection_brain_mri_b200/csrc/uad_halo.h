// Index arithmetic of the halo-resident gather candidate (gather_gemm_tc_np_halo, csrc/uad_conv_tc.cu).  Host/device so that
// tests/test_halo_window.py can run it with g++.  The halo of a TH x TW pixel tile is the (TH + 2) x (TW + 2) pixel box that
// starts one pixel above / left of the tile (TMA zero-fills outside the image); rows of 128 bytes, one per halo pixel.
#pragma once
#if defined(__CUDACC__)
#define UAD_HALO_HD __host__ __device__ __forceinline__
#else
#define UAD_HALO_HD static inline
#endif

// halo row read by tile pixel (th, tw) for a tap with input offset (dh, dw), dh, dw in [-1, 1]
UAD_HALO_HD int uad_halo_row(int th, int tw, int dh, int dw, int TW) { return (th + dh + 1) * (TW + 2) + tw + dw + 1; }
