// Data-parallel optimiser step as ONE kernel over NVLink peer memory: reduce-scatter of the flat gradient buffers, TF-form Adam on
// the rank's own shard, all-gather of the updated parameters.
//
// The reference's step ends with `optimizer.minimize` on one device (trainers/DLMODEL.py:112-131); data parallelism adds one sum of
// the per-rank gradients in front of it.  Calling NCCL for that sum and then the Adam kernel is the baseline (engine.train_step does
// exactly that by default): three launches, 2 x 8.8 MB through the collective, Adam over the whole buffer on every rank.  Here each
// rank r owns the shard [r * chunk, (r + 1) * chunk) of the flat buffers:
//     phase 0   tell every peer "my gradients are complete" (a flag in THEIR memory), wait for theirs
//     phase 1   for the own shard: g = sum_j grads_j[i] read straight from the peers' buffers in rank order (every element is summed
//               by exactly one rank -> every rank sees bit-identical parameters), Adam on the local m / v shard, the new parameter
//               written into EVERY rank's parameter buffer
//     phase 2   tell every peer "my shard of your parameters is written", wait for theirs
// so 2 x (W - 1) / W x 8.8 MB cross NVLink per rank as in a ring all-reduce, Adam touches 1 / W of the state per rank, and nothing
// is launched between the backward pass and the next forward pass but this kernel (CUDA-graph capturable: no host state - the flag
// values are a sequence number kept in device memory).
// Buffers: each rank allocates ONE region [params | grads | flags] with uad_peer_alloc (cudaMalloc, so it has an IPC handle),
// exchanges uad_peer_ipc_handle blobs through the host-side process group and maps the others with uad_peer_ipc_open.
// Spins are bounded (~2 min): a missing peer traps instead of hanging the GPU for good.
#include <stdint.h>
#include <string.h>

#include "uad_common.cuh"

#define UAD_PEER_MAX 16
#define UAD_PEER_FLAG_WORDS 64          // per rank: [0,16) ready[src], [16,32) done[src], 32 block counter, 33 sequence number, 34 blocks so far

struct PeerPtrs {
  float* params[UAD_PEER_MAX];
  const float* grads[UAD_PEER_MAX];
  unsigned long long* flags[UAD_PEER_MAX];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_f4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void spin_until(const unsigned long long* p, unsigned long long want) {
  const long long t0 = clock64();
  while (ld_acquire_sys(p) < want) {
    if (clock64() - t0 > (1ll << 38)) __trap();               // ~2 min at 2 GHz: a peer never arrived (host-side skew between
                                                              // ranks - a rank that evaluates or logs - is seconds and must not trip it)
    __nanosleep(64);
  }
}

__global__ void __launch_bounds__(256) peer_rs_adam_ag_kernel(const __grid_constant__ PeerPtrs pp, int rank, int world,
                                                              float* __restrict__ m, float* __restrict__ v, size_t n, size_t chunk,
                                                              float lr, float b1, float b2, float eps, float gs,
                                                              const long long* __restrict__ step_dev) {
  unsigned long long* myflags = pp.flags[rank];
  const unsigned long long seq = myflags[33] + 1ull;          // written by this rank's previous invocation only
  const unsigned long long blocks_done = myflags[34] + (unsigned long long)gridDim.x;   // block counter value once all my blocks have arrived
  // ---- phase 0: my gradients are complete (stream order) -> tell the peers; wait until theirs are
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(pp.flags[threadIdx.x] + rank, seq);
  }
  if (threadIdx.x < world) spin_until(myflags + threadIdx.x, seq);
  __syncthreads();

  // ---- phase 1: own shard
  float lr_t = lr;
  if (step_dev) {
    const double t = (double)(*step_dev);
    lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
  }
  const size_t lo = (size_t)rank * chunk;
  const size_t hi = lo + chunk < n ? lo + chunk : n;
  float* myp = pp.params[rank];
  for (size_t i = lo + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < hi; i += (size_t)gridDim.x * blockDim.x * 4) {
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < world; ++j) {                          // fixed rank order: deterministic
      const float4 t4 = ld_relaxed_sys_f4(pp.grads[j] + i);
      g4.x += t4.x; g4.y += t4.y; g4.z += t4.z; g4.w += t4.w;
    }
    const float4 p4 = *reinterpret_cast<const float4*>(myp + i);
    float4 m4 = *reinterpret_cast<float4*>(m + i), v4 = *reinterpret_cast<float4*>(v + i);
    float pw[4] = {p4.x, p4.y, p4.z, p4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
    float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {                              // the arithmetic of adam_tf_kernel (uad_elementwise.cu), term by term
      const float gj = gg[e] * gs;
      mm[e] = b1 * mm[e] + (1.f - b1) * gj;
      vv[e] = b2 * vv[e] + (1.f - b2) * gj * gj;
      pw[e] = pw[e] - lr_t * mm[e] / (sqrtf(vv[e]) + eps);
    }
    *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    const float4 out = make_float4(pw[0], pw[1], pw[2], pw[3]);
    for (int j = 0; j < world; ++j) st_relaxed_sys_f4(pp.params[j] + i, out);
  }

  // ---- phase 2: all my blocks have written -> tell the peers; wait until every shard of MY parameters has arrived
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(myflags + 32, 1ull);
    if (blockIdx.x == 0) {
      spin_until(myflags + 32, blocks_done);
      __threadfence_system();
      for (int j = 0; j < world; ++j) st_release_sys(pp.flags[j] + 16 + rank, seq);
    }
  }
  if (threadIdx.x < world) spin_until(myflags + 16 + threadIdx.x, seq);
  __syncthreads();
  // every block has read seq / blocks_done (they all passed the counter) before the two values move on
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    spin_until(myflags + 32, blocks_done);
    myflags[33] = seq;
    myflags[34] = blocks_done;
  }
}

extern "C" size_t uad_peer_region_bytes(size_t numel) {
  return 2 * ((numel * sizeof(float) + 255) & ~(size_t)255) + UAD_PEER_FLAG_WORDS * sizeof(unsigned long long);
}

extern "C" int uad_peer_alloc(size_t bytes, void** out) {
  UAD_REQUIRE(out && bytes > 0, "uad_peer_alloc: bad arguments");
  UAD_CUDA(cudaMalloc(out, bytes));
  UAD_CUDA(cudaMemset(*out, 0, bytes));
  UAD_CUDA(cudaDeviceSynchronize());
  return 0;
}

extern "C" int uad_peer_free(void* p) {
  if (p) UAD_CUDA(cudaFree(p));
  return 0;
}

extern "C" int uad_peer_ipc_handle(void* region, void* handle64) {
  UAD_REQUIRE(region && handle64, "uad_peer_ipc_handle: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  UAD_CUDA(cudaIpcGetMemHandle(&h, region));
  memcpy(handle64, &h, sizeof(h));
  return 0;
}

extern "C" int uad_peer_ipc_open(const void* handle64, void** out) {
  UAD_REQUIRE(handle64 && out, "uad_peer_ipc_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  UAD_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int uad_peer_ipc_close(void* p) {
  if (p) UAD_CUDA(cudaIpcCloseMemHandle(p));
  return 0;
}

// regions[j] = base of rank j's region as mapped in THIS process (own region for j == rank); layout [params | grads | flags] with
// `numel` floats per buffer.  The step covers the slice [offset, offset + count) of the flat index space (one optimiser's variables:
// f-AnoGAN runs three Adam optimisers on scope-contiguous slices); m / v point at the slice's first element.
extern "C" int uad_peer_adam_step(void* const* regions, int rank, int world, size_t numel, size_t offset, size_t count, float* m,
                                  float* v, float lr, float b1, float b2, float eps, float grad_scale, const int64_t* step_dev,
                                  void* stream) {
  UAD_REQUIRE(regions && world >= 1 && world <= UAD_PEER_MAX && rank >= 0 && rank < world, "uad_peer_adam_step: bad rank / world");
  UAD_REQUIRE(m && v && ((uintptr_t)m % 16 == 0) && ((uintptr_t)v % 16 == 0), "uad_peer_adam_step: unaligned Adam state");
  UAD_REQUIRE(numel % 4 == 0 && offset % 4 == 0 && count % 4 == 0 && count > 0 && offset + count <= numel,
              "uad_peer_adam_step: the slice must lie inside the flat buffer and be a multiple of 4 floats");
  const size_t half = (numel * sizeof(float) + 255) & ~(size_t)255;
  PeerPtrs pp;
  memset(&pp, 0, sizeof(pp));
  for (int j = 0; j < world; ++j) {
    UAD_REQUIRE(regions[j] != nullptr, "uad_peer_adam_step: region %d is not mapped", j);
    char* base = (char*)regions[j];
    pp.params[j] = (float*)base + offset;
    pp.grads[j] = (const float*)(base + half) + offset;
    pp.flags[j] = (unsigned long long*)(base + 2 * half);
  }
  size_t chunk = (count + world - 1) / world;
  chunk = (chunk + 3) & ~(size_t)3;
  long long blocks = (long long)((chunk / 4 + 255) / 256);
  if (blocks > UAD_NUM_SMS) blocks = UAD_NUM_SMS;               // one wave: the in-kernel block counter needs every block resident
  if (blocks < 1) blocks = 1;
  peer_rs_adam_ag_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(pp, rank, world, m, v, count, chunk, lr, b1, b2, eps,
                                                                       grad_scale, (const long long*)step_dev);
  UAD_LAUNCH_CHECK("peer_rs_adam_ag");
  return 0;
}
