"""GPU (first green hardware run: round 2, gpurun call r2d): first hardware check of the AnoVAEGAN engine / trainer, written after round 1's GPU budget was
spent.  Its call sequences already match the oracle on CPU through the ABI emulator (tests/test_engine_emulated.py); what
remains to be seen on the B200 is the same comparison with the real kernels, CUDA-graph replay and the device RNG streams."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import anovaegan_cpu as AO  # noqa: E402
from oracle import fanogan_cpu as FO  # noqa: E402
from oracle import tf_graph_cpu as O  # noqa: E402

TOL, GTOL = 1e-4, 1e-4


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))


def _setup(S, B, rate, mode, zDim=128):
    from unsupervised_anomaly_detection_brain_mri_b200.anovaegan_engine import AnoVaeGanEngine
    P = FO.perturb(AO.init_params(S, zDim=zDim, seed=1))
    eng = AnoVaeGanEngine(S, zDim=zDim, batch=B, math_mode=mode, kl_weight=1.0, scale=10.0)
    rng = np.random.default_rng(21)
    x = O.synthetic_slices(B, S, seed=21)
    eps = rng.standard_normal((B, zDim)).astype(np.float32)
    alpha = rng.random((B, 1), dtype=np.float32)
    masks = {'mu': (rng.uniform(size=(B, zDim)) >= rate).astype(np.float32), 'ls': (rng.uniform(size=(B, zDim)) >= rate).astype(np.float32),
             'dec': (rng.uniform(size=(B, eng.flat)) >= rate).astype(np.float32)}
    eng.fp.load(P)
    eng.enable_training()
    eng.set_inputs(x)
    eng.set_noise(eps)
    eng.alpha.copy_(torch.from_numpy(alpha.reshape(-1)))
    eng.mask_mu.copy_(torch.from_numpy(masks['mu']))
    eng.mask_ls.copy_(torch.from_numpy(masks['ls']))
    eng.mask_gen.copy_(torch.from_numpy(masks['dec']))
    return eng, P, x, eps, alpha, masks


def _signs(eng, which, keep):
    from unsupervised_anomaly_detection_brain_mri_b200 import abi

    def pat(ts):
        return [(t > 0).cpu().numpy() for t in ts]

    def critic(x_dev):
        eng._critic_forward(eng.pass1, x_dev, critic=False)
        return pat(eng.pass1.a)

    out = eng._forward_out(True, keep)
    sg = {'enc': pat(eng.enc_a), 'gen': pat([eng.ar] + eng.gen_a)}
    l1_sign = np.sign(out.cpu().numpy() - eng.x.cpu().numpy())
    if which in ('gen', 'disc'):
        sg['d_fake'] = critic(out)
    if which == 'disc':
        sg['d_real'] = critic(eng.x)
        abi.call('uad_interpolate', eng.x.data_ptr(), out.data_ptr(), eng.alpha.data_ptr(), eng.x_hat.data_ptr(), eng.B, eng.S * eng.S,
                 torch.cuda.current_stream().cuda_stream)
        sg['d_hat'] = critic(eng.x_hat)
    torch.cuda.synchronize()
    return sg, l1_sign


@pytest.mark.parametrize('mode', [0, 1])
@pytest.mark.parametrize('S,B', [(32, 4), (64, 2)])
@pytest.mark.parametrize('which', ['vae', 'gen', 'disc'])
def test_anovaegan_train_ops_match_oracle(which, S, B, mode):
    rate, lr = 0.2, 1e-3
    eng, P, x, eps, alpha, masks = _setup(S, B, rate, mode)
    tr = AO.Trainer(P, lr=lr, dropout_rate=rate, scale=10.0, kl_weight=1.0, dtype=torch.float64)
    sg, l1_sign = _signs(eng, which, 1.0 / (1.0 - rate))
    out, G = tr.step(which, x, eps, alpha, masks, signs=sg, l1_sign=l1_sign)
    res = {'vae': eng.step_vae, 'gen': eng.step_gen, 'disc': eng.step_disc}[which](lr, dropout_rate=rate, dropout=True, parity_noise=True)
    torch.cuda.synchronize()
    for k, v in res.items():
        if k in out:
            assert abs(v - float(out[k])) <= TOL * max(abs(float(out[k])), 1e-3), (k, v, float(out[k]))
    assert _rel(eng.x_gen.cpu().numpy(), out['out'].numpy()) < TOL
    got = eng.fp.to_numpy(eng.fp.grads)
    gmax = max(float(v.abs().max()) for v in G.values())
    for k, v in G.items():
        ref = v.numpy()
        if k.endswith('/bias') and float(np.abs(ref).max()) < 1e-5 * gmax:
            assert float(np.abs(got[k]).max()) <= 1e-5 * gmax, k
            continue
        assert _rel(got[k], ref) < GTOL, (k, _rel(got[k], ref))


@pytest.mark.parametrize('which', ['vae', 'gen', 'disc'])
def test_anovaegan_graph_replay_equals_eager(which):
    runs = []
    for use_graph in (False, True):
        eng, *_ = _setup(32, 4, 0.2, 1)
        step = {'vae': eng.step_vae, 'gen': eng.step_gen, 'disc': eng.step_disc}[which]
        res = [step(1e-3, dropout_rate=0.2, dropout=True, parity_noise=True, use_graph=use_graph) for _ in range(5)]
        torch.cuda.synchronize()
        assert (len(eng._graphs) == 1) == use_graph
        runs.append((eng.fp.to_numpy(), res, dict(eng.t), eng.m_gen.cpu().numpy().copy()))
    (w0, r0, t0, m0), (w1, r1, t1, m1) = runs
    assert t0 == t1 and r0 == r1 and np.array_equal(m0, m1)
    assert all(np.array_equal(w0[k], w1[k]) for k in w0)


def test_anovaegan_trainer(tmp_path):
    from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import SYNTHETIC
    from unsupervised_anomaly_detection_brain_mri_b200.models.anovaegan import anovaegan
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.AnoVAEGAN import AnoVAEGAN
    config = AnoVAEGAN.Config()
    config.outputHeight = config.outputWidth = 32
    config.batchsize, config.numEpochs, config.zDim, config.numChannels = 4, 1, 128, 1
    config.intermediateResolutions = [8, 8]
    config.dropout_rate, config.learningrate = 0.1, 1e-4
    config.checkpointDir = str(tmp_path / 'ckpt')
    config.description, config.dataset = 'gpu-test', 'SYNTHETIC'
    opts = SYNTHETIC.Options()
    opts.sliceResolution = (32, 32)
    opts.numPatients = 1
    opts.sliceStart, opts.sliceEnd = 20, 60
    ds = SYNTHETIC(opts)
    model = AnoVAEGAN(None, config, network=anovaegan)
    w0 = model.engine.fp.to_numpy()
    model.train(ds)
    w1 = model.engine.fp.to_numpy()
    assert all(np.isfinite(v).all() for v in w1.values()) and any(not np.array_equal(w0[k], w1[k]) for k in w0)
    t = model.engine.t
    assert t['vae'] == t['gen'] > 0 and t['disc'] == 5 * t['gen']
    rec = model.reconstruct(ds.next_batch(4, set='VAL')[0][0])
    assert rec['reconstruction'].shape == (1, 32, 32, 1) and np.isfinite(rec['l1err'])
    rec2 = model.reconstruct(ds.next_batch(4, set='VAL')[0], dropout=True)
    assert rec2['reconstruction'].shape == (4, 32, 32, 1)
