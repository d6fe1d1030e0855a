"""GPU (first green hardware run: round 2, gpurun call r2d): first hardware check of the GMVAE pieces written after round 1's GPU budget was spent - the latent
kernel pair (uad_gmvae_latent_fwd / _bwd; its arithmetic header already matches float64 autograd in a host build,
tests/test_gmvae_latent.py), the GMVAE train step and one restoration iteration (both already verified on CPU through the ABI emulator)."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import gmvae_cpu as GO  # noqa: E402
from oracle import tf_graph_cpu as O  # noqa: E402


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))


@pytest.mark.parametrize('B,dz,dc,c_lambda', [(64, 128, 9, 1.0), (5, 1, 6, 0.0), (33, 16, 32, 100.0)])
def test_latent_kernels(B, dz, dc, c_lambda):
    from test_gmvae_latent import reference
    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    rng = np.random.default_rng(1)
    host = [rng.standard_normal((B, dz)).astype(np.float32), (0.3 * rng.standard_normal((B, dz)) - 0.5).astype(np.float32),
            rng.standard_normal((B, dz)).astype(np.float32), (0.2 * rng.standard_normal((B, dz, dc))).astype(np.float32),
            (0.3 * rng.standard_normal((B, dz, dc)) + 0.1).astype(np.float32)]
    dev = [torch.from_numpy(a).cuda() for a in host]
    pc, con, closs = torch.empty(B, dc, device='cuda'), torch.empty(B, device='cuda'), torch.empty(B, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    abi.call('uad_gmvae_latent_fwd', *[d.data_ptr() for d in dev], pc.data_ptr(), con.data_ptr(), closs.data_ptr(), B, dz, dc, float(c_lambda), st)
    t = [torch.from_numpy(a).double().requires_grad_(True) for a in host]
    rpc, rcon, rcloss, _ = reference(*t, dc, c_lambda)
    assert _rel(pc.cpu().numpy(), rpc.detach().numpy()) < 1e-5 and _rel(con.cpu().numpy(), rcon.detach().numpy()) < 1e-5
    assert _rel(closs.cpu().numpy(), rcloss.detach().numpy()) < 1e-5
    outs = [torch.empty_like(d) for d in dev]
    abi.call('uad_gmvae_latent_bwd', *[d.data_ptr() for d in dev], 1.0 / B, *[o.data_ptr() for o in outs], B, dz, dc, float(c_lambda), st)
    for o, g in zip(outs, torch.autograd.grad((rcon + rcloss).sum() / B, t)):
        assert _rel(o.cpu().numpy(), g.numpy()) < 1e-5


def _setup(S, B, rate, mode, dz=128, dw=1, dc=9, c_lambda=0.01):
    from unsupervised_anomaly_detection_brain_mri_b200.engine import GMVAE, ConvAutoencoderEngine
    P = GO.perturb(GO.init_params(S, dim_z=dz, dim_w=dw, dim_c=dc, seed=1))
    eng = ConvAutoencoderEngine(GMVAE, S, zDim=dz, batch=B, math_mode=mode, dim_w=dw, dim_c=dc, c_lambda=c_lambda)
    eng.fp.load(P)
    rng = np.random.default_rng(9)
    x = O.synthetic_slices(B, S, seed=31)
    eps_w, eps_z = rng.standard_normal((B, dw)).astype(np.float32), rng.standard_normal((B, dz)).astype(np.float32)
    mk = lambda n: (rng.uniform(size=(B, n)) >= rate).astype(np.float32)   # noqa: E731
    masks = {'w_mu': mk(dw), 'w_ls': mk(dw), 'z_mu': mk(dz), 'dec': mk(eng.flat)}
    eng.set_inputs(x)
    eng.br[0].eps_w.copy_(torch.from_numpy(eps_w))
    eng.set_noise(eps_z, {'wmu': masks['w_mu'], 'wls': masks['w_ls'], 'mu': masks['z_mu'], 'dec': masks['dec']})
    return eng, P, x, eps_w, eps_z, masks


@pytest.mark.parametrize('mode', [0, 1])
@pytest.mark.parametrize('S,B', [(64, 4), (128, 2)])
def test_gmvae_train_step(S, B, mode):
    rate, lr = 0.2, 1e-3
    eng, P, x, eps_w, eps_z, masks = _setup(S, B, rate, mode)
    eng._keep = 1.0 / (1.0 - rate)
    eng.forward(training=True, dropout_rate=rate)
    sgn = np.sign(eng.br[0].xhat.cpu().numpy().astype(np.float64) - x)
    eng.train_step(lr, beta1=0.5, dropout_rate=rate, dropout=True, parity_noise=True)
    torch.cuda.synchronize()
    o, L, G = GO.loss_and_grads(P, x, eps_w, eps_z, masks, rate, True, 9, 0.01, torch.float64, l1_sign=sgn)
    assert _rel(eng.br[0].xhat.cpu().numpy(), o['xz_mu'].numpy()) < 1e-4 and _rel(eng.br[0].pc.cpu().numpy(), o['pc'].numpy()) < 1e-4
    got = eng.losses()
    for k in got:
        assert abs(got[k] - float(L[k])) <= 1e-4 * max(abs(float(L[k])), 1e-6), (k, got[k], float(L[k]))
    grads = eng.fp.to_numpy(eng.fp.grads)
    for k in P:
        assert _rel(grads[k], G[k].numpy()) < 5e-4, (k, _rel(grads[k], G[k].numpy()))


def test_gmvae_restoration_graph_equals_eager():
    outs = []
    for use_graph in (False, True):
        eng, *_ = _setup(64, 4, 0.0, 1)
        for br in eng.br:
            br.masks = {k: None for k in br.masks}
        torch.manual_seed(0)
        eng.rng_ctr.zero_()
        eng.restore(4, 1e-3, 1.5, use_graph=use_graph)
        torch.cuda.synchronize()
        outs.append(eng.br[0].x.cpu().numpy().copy())
    assert np.isfinite(outs[0]).all() and np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize('mode', [0, 1])
def test_gmvae_spatial_train_step(mode):
    from unsupervised_anomaly_detection_brain_mri_b200.engine import GMVAES, ConvAutoencoderEngine
    S, B, dz, dw, dc, c_lambda, lr = 64, 4, 1, 1, 9, 0.01, 1e-3
    P = GO.perturb(GO.init_params_spatial(S, dim_z=dz, dim_w=dw, dim_c=dc, seed=1))
    eng = ConvAutoencoderEngine(GMVAES, S, zDim=dz, batch=B, math_mode=mode, dim_w=dw, dim_c=dc, c_lambda=c_lambda)
    eng.fp.load(P)
    rng = np.random.default_rng(9)
    x = O.synthetic_slices(B, S, seed=31)
    eps_w, eps_z = rng.standard_normal((B, 8, 8, dw)).astype(np.float32), rng.standard_normal((B, 8, 8, dz)).astype(np.float32)
    eng.set_inputs(x)
    eng.br[0].eps_w.copy_(torch.from_numpy(eps_w.reshape(-1, dw)))
    eng.br[0].eps.copy_(torch.from_numpy(eps_z.reshape(-1, dz)))
    eng._keep = 1.0
    eng.forward(training=True, dropout_rate=0.0)
    sgn = np.sign(eng.br[0].xhat.cpu().numpy().astype(np.float64) - x)
    eng.train_step(lr, beta1=0.5, dropout_rate=0.0, dropout=False, parity_noise=True)
    torch.cuda.synchronize()
    o, L, G = GO.loss_and_grads_spatial(P, x, eps_w, eps_z, dc, c_lambda, torch.float64, l1_sign=sgn)
    assert _rel(eng.br[0].xhat.cpu().numpy(), o['xz_mu'].numpy()) < 1e-4
    got = eng.losses()
    for k in got:
        assert abs(got[k] - float(L[k])) <= 1e-4 * max(abs(float(L[k])), 1e-6), (k, got[k], float(L[k]))
    grads = eng.fp.to_numpy(eng.fp.grads)
    for k in P:
        assert _rel(grads[k], G[k].numpy()) < 5e-4, (k, _rel(grads[k], G[k].numpy()))
