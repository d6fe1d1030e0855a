// GMVAE latent block on the device: one thread per sample (B <= a few hundred, ~dz*dc = 1152 terms each: microseconds, no
// inter-thread communication - the arithmetic lives in uad_gmvae_latent.h, which the CPU test-suite compiles and checks too).
#include "uad_common.cuh"
#include "uad_gmvae_latent.h"

__global__ void gmvae_latent_fwd_kernel(const float* __restrict__ z_mu, const float* __restrict__ z_ls, const float* __restrict__ z_s,
                                        const float* __restrict__ M, const float* __restrict__ S, float* __restrict__ pc,
                                        float* __restrict__ con, float* __restrict__ closs, int B, int dz, int dc, float c_lambda) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const size_t o = (size_t)b * dz, oc = (size_t)b * dz * dc;
  uad_gmvae_latent_fwd_sample(z_mu + o, z_ls + o, z_s + o, M + oc, S + oc, dz, dc, c_lambda, pc ? pc + (size_t)b * dc : nullptr, con + b,
                              closs + b);
}

__global__ void gmvae_latent_bwd_kernel(const float* __restrict__ z_mu, const float* __restrict__ z_ls, const float* __restrict__ z_s,
                                        const float* __restrict__ M, const float* __restrict__ S, float scale, float* __restrict__ dz_mu,
                                        float* __restrict__ dz_ls, float* __restrict__ dz_s, float* __restrict__ dM,
                                        float* __restrict__ dS, int B, int dz, int dc, float c_lambda) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const size_t o = (size_t)b * dz, oc = (size_t)b * dz * dc;
  uad_gmvae_latent_bwd_sample(z_mu + o, z_ls + o, z_s + o, M + oc, S + oc, dz, dc, c_lambda, scale, dz_mu + o, dz_ls + o, dz_s + o,
                              dM + oc, dS + oc);
}

extern "C" int uad_gmvae_latent_fwd(const float* z_mu, const float* z_ls, const float* z_s, const float* M, const float* S, float* pc,
                                    float* con, float* closs, int B, int dz, int dc, float c_lambda, void* stream) {
  UAD_REQUIRE(B > 0 && dz > 0 && dc > 0 && dc <= UAD_GMVAE_MAX_C, "uad_gmvae_latent_fwd: B=%d dz=%d dc=%d (dc <= %d)", B, dz, dc,
              UAD_GMVAE_MAX_C);
  gmvae_latent_fwd_kernel<<<uad_cdiv(B, 32), 32, 0, (cudaStream_t)stream>>>(z_mu, z_ls, z_s, M, S, pc, con, closs, B, dz, dc, c_lambda);
  UAD_LAUNCH_CHECK("uad_gmvae_latent_fwd");
  return 0;
}

extern "C" int uad_gmvae_latent_bwd(const float* z_mu, const float* z_ls, const float* z_s, const float* M, const float* S, float scale,
                                    float* dz_mu, float* dz_ls, float* dz_s, float* dM, float* dS, int B, int dz, int dc,
                                    float c_lambda, void* stream) {
  UAD_REQUIRE(B > 0 && dz > 0 && dc > 0 && dc <= UAD_GMVAE_MAX_C, "uad_gmvae_latent_bwd: B=%d dz=%d dc=%d (dc <= %d)", B, dz, dc,
              UAD_GMVAE_MAX_C);
  gmvae_latent_bwd_kernel<<<uad_cdiv(B, 32), 32, 0, (cudaStream_t)stream>>>(z_mu, z_ls, z_s, M, S, scale, dz_mu, dz_ls, dz_s, dM, dS, B,
                                                                         dz, dc, c_lambda);
  UAD_LAUNCH_CHECK("uad_gmvae_latent_bwd");
  return 0;
}
