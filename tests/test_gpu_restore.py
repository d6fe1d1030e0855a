"""Iterative MAP restoration (reference trainers/VAE_You.py:53-54,125-147) on the device: the TV seed kernel is bit-exact
against its numpy statement, one iteration's gradient matches the float64 oracle (tf.gradients restated with autograd),
graph replay == eager, and the VAE_You trainer surface behaves like the reference's."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tf_graph_cpu as O  # noqa: E402

TOL = 1e-4


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(float(np.abs(b).max()), 1e-30))


def _tv_seed_numpy(x, xh, lam):
    d = (x - xh).astype(np.float32)
    T = np.zeros_like(d)
    dv = np.sign(d[:, 1:, :] - d[:, :-1, :])
    dh = np.sign(d[:, :, 1:] - d[:, :, :-1])
    T[:, 1:, :] += dv
    T[:, :-1, :] -= dv
    T[:, :, 1:] += dh
    T[:, :, :-1] -= dh
    g = np.sign(-d) - np.float32(lam) * T
    tv = np.abs(d[:, 1:, :] - d[:, :-1, :]).astype(np.float64).sum((1, 2)) + np.abs(d[:, :, 1:] - d[:, :, :-1]).astype(np.float64).sum((1, 2))
    return g.astype(np.float32), tv


@pytest.mark.parametrize('B,H,W,lam', [(3, 40, 72, 1.8), (2, 8, 32, -0.5), (1, 256, 256, 1.0), (2, 33, 5, 0.0)])
def test_tv_restore_seed_bitexact(B, H, W, lam):
    from gpu_util import call, dev, dptr, empty, ptr, st, sync, workspace
    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    rng = np.random.default_rng(B * H + W)
    x = rng.random((B, H, W), dtype=np.float32)
    xh = rng.random((B, H, W), dtype=np.float32)
    x[:, : H // 3] = 0.0                       # flat regions: sign(0) = 0 on both the L1 and the TV terms
    xh[:, : H // 4] = 0.0
    wsb = abi.lib().uad_tv_restore_workspace_bytes(B, H, W)
    ws = workspace(wsb)
    g, tv = empty(B, H, W), empty(B)
    call('uad_tv_restore_seed', dptr(x), dptr(xh), lam, ptr(g), ptr(tv), B, H, W, ptr(ws), wsb, st())
    sync()
    g_ref, tv_ref = _tv_seed_numpy(x, xh, lam)
    assert np.array_equal(g.cpu().numpy(), g_ref)
    assert _rel(tv.cpu().numpy(), tv_ref) < 1e-5
    # the update kernel:  x <- x - lr*(gx - g)
    gx = rng.standard_normal((B, H, W)).astype(np.float32)
    xd, grads = dev(x), empty(B, H, W)
    call('uad_restore_update', ptr(xd), dptr(gx), ptr(g), 1e-3, ptr(grads), x.size, st())
    sync()
    assert np.array_equal(grads.cpu().numpy(), gx - g_ref)
    assert np.allclose(xd.cpu().numpy(), x - np.float32(1e-3) * (gx - g_ref), rtol=0, atol=1e-7)


@pytest.mark.parametrize('mode', [0, 1])
@pytest.mark.parametrize('arch,S,B,lam', [(O.VAE, 64, 3, 1.8), (O.VAE, 128, 2, 0.0), (O.AE, 64, 2, 0.7)])
def test_restore_gradient_matches_oracle(arch, S, B, lam, mode):
    from unsupervised_anomaly_detection_brain_mri_b200.engine import ConvAutoencoderEngine
    P = O.perturb_params(O.init_params(arch, S, seed=1))
    x = O.synthetic_slices(B, S, seed=11)
    eps = np.random.default_rng(4).standard_normal((B, 128)).astype(np.float32)
    eng = ConvAutoencoderEngine(arch, S, batch=B, math_mode=mode)
    eng.fp.load(P)
    eng.set_inputs(x)
    eng.set_noise(eps)
    lr = 1e-3
    eng.restore_step(lr, lam, parity_noise=True, keep_grads=True)
    torch.cuda.synchronize()
    xh = eng.br[0].xhat.cpu().numpy()
    g_ref, out, tv_ref = O.restore_gradient(arch, P, x, eps=eps, tv_lambda=lam, dtype=torch.float64, sign_from=xh)
    assert _rel(xh, out['x_hat'].numpy()) < TOL
    assert _rel(eng.tv.cpu().numpy(), tv_ref.numpy()) < TOL
    got = eng.restore_grads.cpu().numpy()
    assert _rel(got, g_ref.numpy()) < 5 * TOL
    # literal |.| (no fixed sign pattern) differs on almost no pixel
    g_lit, _, _ = O.restore_gradient(arch, P, x, eps=eps, tv_lambda=lam, dtype=torch.float64)
    assert (np.abs(g_lit.numpy() - g_ref.numpy()) > 1e-6).mean() < 2e-3
    assert np.allclose(eng.br[0].x.cpu().numpy(), x - np.float32(lr) * got, rtol=0, atol=1e-6)


def test_restore_loop_graph_equals_eager_and_tracks_oracle():
    from unsupervised_anomaly_detection_brain_mri_b200.engine import ConvAutoencoderEngine
    arch, S, B, steps, lr, lam = O.VAE, 64, 2, 6, 1e-3, 1.0
    P = O.perturb_params(O.init_params(arch, S, seed=2))
    x = O.synthetic_slices(B, S, seed=5)
    res = []
    for use_graph in (False, True):
        eng = ConvAutoencoderEngine(arch, S, batch=B, math_mode=1, rng_seed=77)
        eng.fp.load(P)
        eng.set_inputs(x)
        eng.restore(steps, lr, lam, use_graph=use_graph)
        torch.cuda.synchronize()
        assert (getattr(eng, '_restore_graph', None) is not None) == use_graph
        res.append(eng.br[0].x.cpu().numpy())
    assert np.array_equal(res[0], res[1])
    assert not np.array_equal(res[0], x)
    # explicit eps per iteration: the trajectory follows the oracle's (sign flips at kinks move single pixels by <= 2*lr*(1+4*lam))
    rng = np.random.default_rng(9)
    eps_list = [rng.standard_normal((B, 128)).astype(np.float32) for _ in range(steps)]
    eng = ConvAutoencoderEngine(arch, S, batch=B, math_mode=1)
    eng.fp.load(P)
    eng.set_inputs(x)
    for k in range(steps):
        eng.set_noise(eps_list[k])
        eng.restore_step(lr, lam, parity_noise=True)
    torch.cuda.synchronize()
    ref = O.restore(arch, P, x, steps=steps, restore_lr=lr, tv_lambda=lam, eps_list=eps_list, dtype=torch.float64)
    diff = np.abs(eng.br[0].x.cpu().numpy() - ref)
    assert float(diff.max()) <= steps * 2 * lr * (1 + 4 * lam) + 1e-6
    assert (diff > 1e-5).mean() < 0.02
    moved = np.abs(ref - x).max()
    assert moved > 10 * np.median(diff)


def test_vae_you_trainer_surface(tmp_path):
    from unsupervised_anomaly_detection_brain_mri_b200.models.variational_autoencoder import variational_autoencoder
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.VAE_You import VAE_You
    from unsupervised_anomaly_detection_brain_mri_b200.utils.default_config_setup import get_config, get_datasets, get_options
    c0 = VAE_You.Config()
    assert (c0.restore_lr, c0.restore_steps, c0.tv_lambda, c0.modelname) == (1e-3, 150, 1.8, 'VAE_You')
    cfgjson = {'CHECKPOINTDIR': str(tmp_path / 'ckpt'), 'SAMPLEDIR': str(tmp_path / 'samples'), 'BRAINWEBDIR': ''}
    options = get_options(batchsize=8, learningrate=1e-4, numEpochs=1, zDim=128, outputWidth=32, outputHeight=32, slices_start=20,
                          slices_end=70, config=cfgjson)
    options['data']['numPatients'] = 2
    options['data']['numTestPatients'] = 1
    hc, _ = get_datasets(options)
    config = get_config(VAE_You, options, 'ADAM', [8, 8], 0.1, hc)
    config.useTensorboard = False
    config.verbose = False
    config.restore_steps = 5
    config.tv_lambda = -1.0
    model = VAE_You(None, config, network=variational_autoencoder)
    model.train(hc)                                    # ends with determine_best_lambda (tv_lambda == -1, VAE_You.py:88-93)
    assert 0.0 <= model.tv_lambda_value <= 1.9
    x = hc.next_batch(6, set='VAL')[0]
    r = model.reconstruct(x)
    assert r['reconstruction'].shape == x.shape and np.isfinite(r['reconstruction']).all()
    assert np.isclose(r['l1err'], np.abs(x - r['reconstruction']).sum(), rtol=1e-5)
    assert not np.array_equal(r['reconstruction'], x)
    r1 = model.reconstruct(x[0])                       # single [H,W,C] slice, as utils/Evaluation.py calls it
    assert r1['reconstruction'].shape == (1,) + x.shape[1:]
    model.restore_steps = 0                            # no iterations: the restored image IS the input (VAE_You.py:130)
    assert np.array_equal(model.reconstruct(x)['reconstruction'], x)
