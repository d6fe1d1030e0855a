"""GPU (first green hardware run: round 2, gpurun call r2d): first hardware check of the adversarial-autoencoder engine / trainer, written after round 1's GPU
budget was spent.  Its call sequences already match the oracle on CPU through the ABI emulator (tests/test_engine_emulated.py);
what remains is the same comparison with the real kernels, CUDA-graph replay and the device RNG streams."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import aae_cpu as AA  # noqa: E402
from oracle import tf_graph_cpu as O  # noqa: E402

TOL, GTOL = 1e-4, 5e-4


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))


def _setup(S, B, rate, mode, zDim=128, constrained=False, rho=0.7):
    from unsupervised_anomaly_detection_brain_mri_b200.aae_engine import AdversarialAEEngine
    P = AA.perturb(AA.init_params(S, zDim=zDim, seed=1, constrained=constrained))
    eng = AdversarialAEEngine(S, zDim=zDim, batch=B, math_mode=mode, scale=10.0, constrained=constrained, rho=rho)
    eng.fp.load(P)
    rng = np.random.default_rng(5)
    x = O.synthetic_slices(B, S, seed=31)
    z = rng.standard_normal((B, zDim)).astype(np.float32)
    epsilon = rng.random((B, 1), dtype=np.float32)
    masks = {'z': (rng.uniform(size=(B, zDim)) >= rate).astype(np.float32), 'dec': (rng.uniform(size=(B, eng.flat)) >= rate).astype(np.float32)}
    eng.set_inputs(x)
    eng.set_latent(z)
    eng.set_epsilon(epsilon)
    eng.set_noise(None, {'mu': masks['z']} if constrained else {'mu': masks['z'], 'dec': masks['dec']})
    return eng, P, x, z, epsilon, masks


def _signs(eng, which, rate):
    from unsupervised_anomaly_detection_brain_mri_b200 import abi

    def pat():
        return [(t > 0).cpu().numpy() for t in eng.cp.pre]

    eng._keep = 1.0 / (1.0 - rate) if rate > 0 else 1.0
    z_ = eng.encode_latent()
    sg = {}
    if which in ('gen', 'disc'):
        eng.critic_forward(z_)
        sg['d_fake'] = pat()
    if which == 'disc':
        eng.critic_forward(eng.z_real)
        sg['d_real'] = pat()
        abi.call('uad_interpolate', eng.z_real.data_ptr(), z_.data_ptr(), eng.epsilon.data_ptr(), eng.z_hat.data_ptr(), eng.B, eng.zDim,
                 torch.cuda.current_stream().cuda_stream)
        eng.critic_forward(eng.z_hat)
        sg['d_hat'] = pat()
    torch.cuda.synchronize()
    return sg


@pytest.mark.parametrize('constrained', [False, True])
@pytest.mark.parametrize('mode', [0, 1])
@pytest.mark.parametrize('S,B', [(32, 4), (64, 2)])
@pytest.mark.parametrize('which', ['ae', 'disc', 'gen'])
def test_aae_train_ops_match_oracle(which, S, B, mode, constrained):
    rate, lr = 0.2, 1e-3
    eng, P, x, z, epsilon, masks = _setup(S, B, rate, mode, constrained=constrained)
    tr = AA.Trainer(P, lr=lr, dropout_rate=rate, scale=10.0, dtype=torch.float64, constrained=constrained, rho=0.7)
    out, G = tr.step(which, x, z, epsilon, masks, signs=_signs(eng, which, rate))
    res = {'ae': eng.step_ae, 'disc': eng.step_disc, 'gen': eng.step_gen}[which](lr, dropout_rate=rate, dropout=True, parity_noise=True)
    torch.cuda.synchronize()
    for k, v in res.items():
        if k in out and out[k].ndim == 0:
            assert abs(v - float(out[k])) <= TOL * max(abs(float(out[k])), 1e-3), (k, v, float(out[k]))
    assert _rel(eng.br[0].mu.cpu().numpy(), out['z_'].numpy()) < TOL
    got = eng.fp.to_numpy(eng.fp.grads)
    for k, v in G.items():
        assert _rel(got[k], v.numpy()) < GTOL, (k, _rel(got[k], v.numpy()))


@pytest.mark.parametrize('which', ['ae', 'disc', 'gen'])
def test_aae_graph_replay_equals_eager(which):
    runs = []
    for use_graph in (False, True):
        eng, *_ = _setup(32, 4, 0.2, 1)
        step = {'ae': eng.step_ae, 'disc': eng.step_disc, 'gen': eng.step_gen}[which]
        res = [step(1e-3, dropout_rate=0.2, dropout=True, parity_noise=True, use_graph=use_graph) for _ in range(5)]
        torch.cuda.synchronize()
        assert (len(eng._graphs) == 1) == use_graph
        runs.append((eng.fp.to_numpy(), res, dict(eng.op_t), eng.m_gen.cpu().numpy().copy()))
    (w0, r0, t0, m0), (w1, r1, t1, m1) = runs
    assert t0 == t1 and r0 == r1 and np.array_equal(m0, m1)
    assert all(np.array_equal(w0[k], w1[k]) for k in w0)


def test_aae_trainer(tmp_path):
    from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import SYNTHETIC
    from unsupervised_anomaly_detection_brain_mri_b200.models.adversarial_autoencoder import adversarial_autoencoder
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.AAE import AAE
    config = AAE.Config()
    config.outputHeight = config.outputWidth = 32
    config.batchsize, config.numEpochs, config.zDim, config.numChannels = 4, 1, 128, 1
    config.intermediateResolutions = [8, 8]
    config.dropout_rate, config.learningrate, config.d_iters = 0.1, 1e-4, 4
    config.checkpointDir = str(tmp_path / 'ckpt')
    config.description, config.dataset = 'gpu-test', 'SYNTHETIC'
    opts = SYNTHETIC.Options()
    opts.sliceResolution = (32, 32)
    opts.numPatients = 1
    opts.sliceStart, opts.sliceEnd = 20, 60
    ds = SYNTHETIC(opts)
    model = AAE(None, config, network=adversarial_autoencoder)
    w0 = model.engine.fp.to_numpy()
    model.train(ds)
    w1 = model.engine.fp.to_numpy()
    assert all(np.isfinite(v).all() for v in w1.values()) and any(not np.array_equal(w0[k], w1[k]) for k in w0)
    rec = model.reconstruct(ds.next_batch(4, set='VAL')[0][0])
    assert rec['reconstruction'].shape == (1, 32, 32, 1) and np.isfinite(rec['l1err'])
    rec2 = model.reconstruct(ds.next_batch(4, set='VAL')[0], dropout=True)
    assert rec2['reconstruction'].shape == (4, 32, 32, 1)
