// tcgen05 / TMA implicit-GEMM conv kernels (placeholder until the tensor-core path lands).
#include "uad_conv.cuh"

int uad_tc_gather_supported(int, int, int, int) { return 0; }
size_t uad_tc_gather_ws_bytes(int, int, int) { return 0; }
int uad_launch_gather_tc(const GatherParams&, int, int, bool, const float*, int, void*, size_t, cudaStream_t) {
  return uad_set_error("tcgen05 conv path not built");
}
