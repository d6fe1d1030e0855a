// Index arithmetic of the epilogue's per-warp staging tile (32 rows x 32 fp32 columns) of gather_gemm_tc2: each lane WRITES its
// own row (one float4 per 4 columns), then the warp READS the tile back row-major so that a quarter-warp stores one 128-byte
// output row.  Two layouts: padded (36 floats per row, shipped) and swizzled (32 floats per row, the 16-byte column group XOR-ed
// with row & 7).  Host/device so that tests/test_staging_layout.py can run it with g++ (the device code includes this file).
#pragma once
#if defined(__CUDACC__)
#define UAD_STG_HD __host__ __device__ __forceinline__
#else
#define UAD_STG_HD static inline
#endif

// floats per staging row
template <bool kSwz>
UAD_STG_HD int uad_stg_ld() { return kSwz ? 32 : 36; }

// float index at which `lane` (== tile row) writes columns j..j+3 (j a multiple of 4)
template <bool kSwz>
UAD_STG_HD int uad_stg_write_index(int lane, int j) {
  return lane * uad_stg_ld<kSwz>() + (kSwz ? (((j >> 2) ^ (lane & 7)) << 2) : j);
}

// row that `lane` reads in iteration `it` (0..7), and the float index of its columns (lane & 7) * 4 .. + 3 in that row
UAD_STG_HD int uad_stg_read_row(int lane, int it) { return it * 4 + (lane >> 3); }

template <bool kSwz>
UAD_STG_HD int uad_stg_read_index(int lane, int it) {
  const int r = uad_stg_read_row(lane, it);
  const int cq = (lane & 7) * 4;
  return r * uad_stg_ld<kSwz>() + (kSwz ? (cq ^ ((r & 7) << 2)) : cq);
}
