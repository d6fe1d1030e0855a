// Mixture-of-Gaussians latent block of the GMVAE (reference models/gaussian_mixture_variational_autoencoder.py:64-71,
// trainers/GMVAE.py:66-88), one SAMPLE per call - plain C so that the very same arithmetic is compiled for the device
// (uad_gmvae.cu: one thread per sample, no inter-thread communication) and, by tests/test_gmvae_latent.py, for the host,
// where it is checked against float64 autograd.
//
// Inputs of a sample:  z_mu[dz], z_ls[dz] (log-variance of q(z|x)), z_s[dz] (the reparameterised sample),
//                      M[dz*dc] = z_wc_mus, S[dz*dc] = z_wc_log_sigma_invs, both laid out [j*dc + c].
//   logit_c = sum_j ( -0.5 (z_s_j - M_jc)^2 e^{S_jc} - S_jc + log(pi) ),   pc = softmax_c(logit)
//   kl_jc   = 0.5 ( (e^{z_ls_j} + (z_mu_j - M_jc)^2) (e^{S_jc} + 1e-6) - S_jc - z_ls_j - 1 )
//   con     = sum_c pc_c sum_j kl_jc                                  (conditional_prior_loss, per sample)
//   closs1  = sum_c pc_c log(pc_c * dc + 1e-8),   c_loss = max(closs1, c_lambda)          (c_prior_loss, per sample)
// All sums run in double; inputs / outputs are float.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define UAD_HD __host__ __device__
#else
#define UAD_HD
#endif

#define UAD_GMVAE_MAX_C 32

// pc[dc], K[dc] (= sum_j kl_jc) of one sample; returns closs1
UAD_HD static inline double uad_gmvae_responsibilities(const float* z_mu, const float* z_ls, const float* z_s, const float* M,
                                                       const float* S, int dz, int dc, double* pc, double* K) {
  const double log_pi = 1.1447298858494002;
  double logit[UAD_GMVAE_MAX_C];
  for (int c = 0; c < dc; ++c) { logit[c] = 0.0; K[c] = 0.0; }
  for (int j = 0; j < dz; ++j) {
    const double zs = z_s[j], zm = z_mu[j], zl = z_ls[j], ezl = exp(zl);
    for (int c = 0; c < dc; ++c) {
      const double m = M[j * dc + c], s = S[j * dc + c], es = exp(s);
      const double d = zs - m, dm = zm - m;
      logit[c] += -0.5 * d * d * es - s + log_pi;
      K[c] += 0.5 * ((ezl + dm * dm) * (es + 1e-6) - s - zl - 1.0);
    }
  }
  double mx = logit[0];
  for (int c = 1; c < dc; ++c) mx = logit[c] > mx ? logit[c] : mx;
  double den = 0.0;
  for (int c = 0; c < dc; ++c) { pc[c] = exp(logit[c] - mx); den += pc[c]; }
  double closs1 = 0.0;
  for (int c = 0; c < dc; ++c) { pc[c] /= den; closs1 += pc[c] * log(pc[c] * dc + 1e-8); }
  return closs1;
}

UAD_HD static inline void uad_gmvae_latent_fwd_sample(const float* z_mu, const float* z_ls, const float* z_s, const float* M,
                                                      const float* S, int dz, int dc, float c_lambda, float* pc_out, float* con_out,
                                                      float* closs_out) {
  double pc[UAD_GMVAE_MAX_C], K[UAD_GMVAE_MAX_C];
  const double closs1 = uad_gmvae_responsibilities(z_mu, z_ls, z_s, M, S, dz, dc, pc, K);
  double con = 0.0;
  for (int c = 0; c < dc; ++c) { con += pc[c] * K[c]; if (pc_out) pc_out[c] = (float)pc[c]; }
  *con_out = (float)con;
  *closs_out = (float)(closs1 > (double)c_lambda ? closs1 : (double)c_lambda);
}

// gradients of scale * (con + c_loss) of one sample w.r.t. z_mu, z_ls, z_s (each dz) and M, S (each dz*dc); z_s is treated as an
// independent input (the caller chains the reparameterisation).  tf.maximum passes the gradient to closs1 when closs1 >= c_lambda.
UAD_HD static inline void uad_gmvae_latent_bwd_sample(const float* z_mu, const float* z_ls, const float* z_s, const float* M,
                                                      const float* S, int dz, int dc, float c_lambda, float scale, float* dz_mu,
                                                      float* dz_ls, float* dz_s, float* dM, float* dS) {
  double pc[UAD_GMVAE_MAX_C], K[UAD_GMVAE_MAX_C], g[UAD_GMVAE_MAX_C], dlogit[UAD_GMVAE_MAX_C];
  const double closs1 = uad_gmvae_responsibilities(z_mu, z_ls, z_s, M, S, dz, dc, pc, K);
  const double gate = closs1 >= (double)c_lambda ? 1.0 : 0.0;
  double dot = 0.0;
  for (int c = 0; c < dc; ++c) {
    const double q = pc[c] * dc + 1e-8;
    g[c] = K[c] + gate * (log(q) + pc[c] * dc / q);            // d(con + c_loss) / d pc_c
    dot += pc[c] * g[c];
  }
  for (int c = 0; c < dc; ++c) dlogit[c] = pc[c] * (g[c] - dot);   // softmax backward
  for (int j = 0; j < dz; ++j) {
    const double zs = z_s[j], zm = z_mu[j], zl = z_ls[j], ezl = exp(zl);
    double a_mu = 0.0, a_ls = 0.0, a_s = 0.0;
    for (int c = 0; c < dc; ++c) {
      const double m = M[j * dc + c], s = S[j * dc + c], es = exp(s);
      const double d = zs - m, dm = zm - m;
      a_s += dlogit[c] * (-d * es);
      a_mu += pc[c] * dm * (es + 1e-6);
      a_ls += pc[c] * 0.5 * (ezl * (es + 1e-6) - 1.0);
      dM[j * dc + c] = (float)(scale * (dlogit[c] * d * es - pc[c] * dm * (es + 1e-6)));
      dS[j * dc + c] = (float)(scale * (dlogit[c] * (-0.5 * d * d * es - 1.0) + pc[c] * 0.5 * ((ezl + dm * dm) * es - 1.0)));
    }
    dz_mu[j] = (float)(scale * a_mu);
    dz_ls[j] = (float)(scale * a_ls);
    dz_s[j] = (float)(scale * a_s);
  }
}
