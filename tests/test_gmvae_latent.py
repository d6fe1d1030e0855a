"""The GMVAE latent block's arithmetic (csrc/uad_gmvae_latent.h) compiled for the HOST with gcc and checked against a float64
torch-autograd restatement of reference models/gaussian_mixture_variational_autoencoder.py:64-71 + trainers/GMVAE.py:66-88.
The device kernel (csrc/uad_gmvae.cu) runs this same header with one thread per sample and no inter-thread communication, so
what remains unverified without a GPU is only its launch, not its math."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, 'unsupervised_anomaly_detection_brain_mri_b200', 'csrc')

SHIM = r'''
#include "uad_gmvae_latent.h"
void fwd(const float* z_mu, const float* z_ls, const float* z_s, const float* M, const float* S, float* pc, float* con, float* closs,
         int B, int dz, int dc, float c_lambda) {
  for (int b = 0; b < B; ++b)
    uad_gmvae_latent_fwd_sample(z_mu + b * dz, z_ls + b * dz, z_s + b * dz, M + b * dz * dc, S + b * dz * dc, dz, dc, c_lambda,
                                pc + b * dc, con + b, closs + b);
}
void bwd(const float* z_mu, const float* z_ls, const float* z_s, const float* M, const float* S, float scale, float* dz_mu, float* dz_ls,
         float* dz_s, float* dM, float* dS, int B, int dz, int dc, float c_lambda) {
  for (int b = 0; b < B; ++b)
    uad_gmvae_latent_bwd_sample(z_mu + b * dz, z_ls + b * dz, z_s + b * dz, M + b * dz * dc, S + b * dz * dc, dz, dc, c_lambda, scale,
                                dz_mu + b * dz, dz_ls + b * dz, dz_s + b * dz, dM + b * dz * dc, dS + b * dz * dc);
}
'''


@pytest.fixture(scope='module')
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp('gmvae')
    src = d / 'shim.c'
    src.write_text(SHIM)
    so = d / 'shim.so'
    subprocess.check_call(['gcc', '-O2', '-shared', '-fPIC', '-std=c99', '-I', HDR, str(src), '-o', str(so), '-lm'])
    return C.CDLL(str(so))


def reference(z_mu, z_ls, z_s, M, S, dc, c_lambda):
    """The reference's graph nodes, literally (float64)."""
    z_sample = z_s.unsqueeze(-1).expand(-1, -1, dc)
    loglh = -0.5 * ((z_sample - M) ** 2 * torch.exp(S)) - S + np.log(np.pi)
    pc = torch.softmax(loglh.sum(1), dim=-1)
    zmu = z_mu.unsqueeze(-1).expand(-1, -1, dc)
    zlv = z_ls.unsqueeze(-1).expand(-1, -1, dc)
    d_var = (torch.exp(zlv) + (zmu - M) ** 2) * (torch.exp(S) + 1e-6)
    kl = (d_var - (S + zlv) - 1) * 0.5
    con = torch.matmul(kl, pc.unsqueeze(-1)).squeeze(-1).sum(1)
    closs1 = (pc * torch.log(pc * dc + 1e-8)).sum(1)
    return pc, con, torch.maximum(closs1, torch.full_like(closs1, c_lambda)), closs1


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize('B,dz,dc,c_lambda', [(5, 128, 9, 1.0), (3, 1, 6, 1.0), (4, 16, 9, 0.05), (2, 7, 32, 0.0), (4, 128, 9, 100.0)])
def test_latent_block_matches_float64_autograd(lib, B, dz, dc, c_lambda):
    rng = np.random.default_rng(dz * 100 + dc)
    z_mu = rng.standard_normal((B, dz)).astype(np.float32)
    z_ls = (0.3 * rng.standard_normal((B, dz)) - 0.5).astype(np.float32)
    z_s = (z_mu + 0.5 * rng.standard_normal((B, dz))).astype(np.float32)
    M = (0.8 * rng.standard_normal((B, dz, dc))).astype(np.float32)
    S = (0.3 * rng.standard_normal((B, dz, dc)) + 0.1).astype(np.float32)
    if dz >= 16:                     # spread the clusters so the softmax is neither uniform nor one-hot
        M *= 0.2
    pc, con, closs = np.zeros((B, dc), np.float32), np.zeros(B, np.float32), np.zeros(B, np.float32)
    lib.fwd(_p(z_mu), _p(z_ls), _p(z_s), _p(M), _p(S), _p(pc), _p(con), _p(closs), B, dz, dc, C.c_float(c_lambda))
    t = [torch.from_numpy(a).double().requires_grad_(True) for a in (z_mu, z_ls, z_s, M, S)]
    rpc, rcon, rcloss, rcloss1 = reference(*t, dc, c_lambda)
    np.testing.assert_allclose(pc, rpc.detach().numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(con, rcon.detach().numpy(), rtol=2e-6)
    np.testing.assert_allclose(closs, rcloss.detach().numpy(), rtol=2e-6, atol=1e-7)
    scale = 1.0 / B
    grads = torch.autograd.grad(scale * (rcon + rcloss).sum(), t)
    outs = [np.zeros_like(a) for a in (z_mu, z_ls, z_s, M, S)]
    lib.bwd(_p(z_mu), _p(z_ls), _p(z_s), _p(M), _p(S), C.c_float(scale), *[_p(o) for o in outs], B, dz, dc, C.c_float(c_lambda))
    for name, o, g in zip(('z_mu', 'z_ls', 'z_s', 'M', 'S'), outs, grads):
        ref = g.numpy()
        err = np.abs(o - ref).max() / max(np.abs(ref).max(), 1e-30)
        assert err < 2e-6, (name, err)
    # the tf.maximum gate: c_lambda = 100 closes it for every sample, 0.0 opens it for every sample
    if c_lambda == 100.0:
        assert (rcloss1.detach().numpy() < c_lambda).all() and (closs == 100.0).all()
    if c_lambda == 0.0:
        assert (rcloss1.detach().numpy() >= 0.0).all()
