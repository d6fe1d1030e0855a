"""Host-side composition checks WITHOUT a GPU (tests/abi_emulator.py replaces every ABI entry point by a float64 torch-CPU
function with the header's semantics).

1. Calibration: the f-AnoGAN engine's three train ops, whose real-kernel runs match the oracle on the B200
   (tests/test_gpu_fanogan.py), must match the same oracle through the emulator - this pins the emulator's reading of the ABI.
2. New compositions that have not run on hardware yet (AnoVAEGAN) are then checked the same way against their own oracle:
   every gradient, every loss scalar, the Adam slot ownership of the three optimisers.
What this does NOT cover: the kernels themselves (GPU parity tests), CUDA-graph capture, device RNG streams."""
import os

import numpy as np
import pytest
import torch

import abi_emulator as E
from oracle import anovaegan_cpu as AO
from oracle import fanogan_cpu as FO
from oracle import tf_graph_cpu as O
from unsupervised_anomaly_detection_brain_mri_b200 import anovaegan_engine, fanogan_engine

TOL = 1e-5


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))


def _pat(ts):
    return [(t > 0).numpy() for t in ts]


def _compare_grads(eng, G, tol=TOL):
    got = eng.fp.to_numpy(eng.fp.grads)
    gmax = max(float(v.abs().max()) for v in G.values())
    for k, v in G.items():
        ref = v.numpy()
        if k.endswith('/bias') and float(np.abs(ref).max()) < 1e-9 * gmax:
            assert float(np.abs(got[k]).max()) <= 1e-9 * gmax, k       # bias feeding a LayerNorm over (H,W): exactly 0
            continue
        assert _rel(got[k], ref) < tol, (k, _rel(got[k], ref))


def _check_update(after, before, ref_P, G, scopes, lr):
    """The Adam update touched exactly the op's variables and equals tf.train.AdamOptimizer's.  The first step is ~lr*sign(g), so
    it is compared where the sign is determined (the engine's gradient buffer is float32, the oracle's float64)."""
    gmax = max(float(v.abs().max()) for v in G.values())
    tot = bad = 0
    for k in after:
        if k.split('/')[0] not in scopes:
            assert np.array_equal(after[k], before[k]), k
            continue
        ref, g = ref_P[k].numpy(), G[k].numpy()
        sel = np.abs(g) > 1e-3 * gmax
        tot += int(sel.sum())
        bad += int((np.abs(after[k].reshape(ref.shape) - ref)[sel] > 0.02 * lr).sum())
    assert tot > 0 and bad == 0, (bad, tot)


def _feed(S, B, rate, flat, zDim=128, seed=21):
    rng = np.random.default_rng(seed)
    x = O.synthetic_slices(B, S, seed=seed)
    z = rng.standard_normal((B, zDim)).astype(np.float32)
    alpha = rng.random((B, 1), dtype=np.float32)
    m_z = [(rng.uniform(size=(B, zDim)) >= rate).astype(np.float32) for _ in range(2)]
    m_gen = (rng.uniform(size=(B, flat)) >= rate).astype(np.float32)
    return x, z, alpha, m_z, m_gen


# ------------------------------------------------------------------------------------------------ 1. calibration on f-AnoGAN
def _fanogan_signs(eng, which, keep):
    def critic(x_dev):
        eng._critic_forward(eng.pass1, x_dev, critic=False)
        return _pat(eng.pass1.a)
    sg = {}
    if which in ('gen', 'disc'):
        eng.generate(eng.z_in, eng.mask_gen, keep, out=eng.x_gen)
        sg['gen_z'] = _pat([eng.ar] + eng.gen_a)
        sg['d_fake'] = critic(eng.x_gen)
    if which == 'disc':
        sg['d_real'] = critic(eng.x)
        E.call('uad_interpolate', eng.x, eng.x_gen, eng.alpha, eng.x_hat, eng.B, eng.S * eng.S, 0)
        sg['d_hat'] = critic(eng.x_hat)
    if which == 'enc':
        z_enc = eng.encode(eng.mask_enc, keep)
        sg['enc'] = _pat(eng.enc_a)
        x_enc = eng.generate(z_enc, eng.mask_gen, keep, out=eng.x_enc)
        sg['gen_enc'] = _pat([eng.ar] + eng.gen_a)
        sg['d_enc'] = critic(x_enc)
        sg['d_real'] = critic(eng.x)
    return sg


@pytest.mark.parametrize('which', ['gen', 'disc', 'enc'])
def test_emulator_reproduces_the_gpu_verified_fanogan_ops(which, monkeypatch):
    E.install(monkeypatch, fanogan_engine)
    S, B, rate, lr = 32, 2, 0.2, 1e-3
    P = FO.perturb(FO.init_params(S, seed=1))
    eng = fanogan_engine.FanoganEngine(S, batch=B, device='cpu', math_mode=0, kappa=1.0, scale=10.0)
    x, z, alpha, m_z, m_gen = _feed(S, B, rate, eng.flat)
    eng.fp.load(P)
    eng.enable_training()
    E.adopt(eng)
    assert E.poison(eng) > 40           # NaN in every work buffer: a read of a never-written buffer would surface in the results
    eng.set_inputs(x)
    eng.set_latent(z)
    eng.alpha.copy_(torch.from_numpy(alpha.reshape(-1)))
    eng.mask_enc.copy_(torch.from_numpy(m_z[0]))
    eng.mask_gen.copy_(torch.from_numpy(m_gen))
    tr = FO.WganTrainer(P, lr=lr, dropout_rate=rate, scale=10.0, kappa=1.0, dtype=torch.float64)
    out, G = tr.step(which, x, z, alpha, mask_enc=m_z[0], mask_gen=m_gen, signs=_fanogan_signs(eng, which, 1.0 / (1.0 - rate)))
    step = {'gen': eng.step_gen, 'disc': eng.step_disc, 'enc': eng.step_enc}[which]
    before = eng.fp.to_numpy()
    res = step(lr, dropout_rate=rate, dropout=True, parity_noise=True)
    for k, v in res.items():
        if k in out:
            assert abs(v - float(out[k])) <= 1e-5 * max(abs(float(out[k])), 1e-3), (k, v, float(out[k]))
    if which == 'disc':
        assert _rel(eng.ddx.numpy(), out['ddx'].numpy()) < TOL
    _compare_grads(eng, G)
    scope = {'gen': 'Generator', 'disc': 'Discriminator', 'enc': 'Encoder'}[which]
    _check_update(eng.fp.to_numpy(), before, tr.P, G, (scope,), lr)


# ------------------------------------------------------------------------------------------------ 2. AnoVAEGAN
def _anovaegan_signs(eng, which, on, keep):
    def critic(x_dev):
        eng._critic_forward(eng.pass1, x_dev, critic=False)
        return _pat(eng.pass1.a)
    out = eng._forward_out(on, keep)
    sg = {'enc': _pat(eng.enc_a), 'gen': _pat([eng.ar] + eng.gen_a)}
    l1_sign = np.sign(out.numpy() - eng.x.numpy())
    if which in ('gen', 'disc'):
        sg['d_fake'] = critic(out)
    if which == 'disc':
        sg['d_real'] = critic(eng.x)
        E.call('uad_interpolate', eng.x, out, eng.alpha, eng.x_hat, eng.B, eng.S * eng.S, 0)
        sg['d_hat'] = critic(eng.x_hat)
    return sg, l1_sign


def _anovaegan(monkeypatch, S=32, B=2, rate=0.2, zDim=128, kl_weight=1.0):
    E.install(monkeypatch, fanogan_engine, anovaegan_engine)
    P = FO.perturb(AO.init_params(S, zDim=zDim, seed=1))
    eng = anovaegan_engine.AnoVaeGanEngine(S, zDim=zDim, batch=B, device='cpu', math_mode=0, kl_weight=kl_weight, scale=10.0)
    assert list(eng.specs) == list(P) and all(tuple(eng.specs[k]) == P[k].shape for k in P)
    x, eps, alpha, m_z, m_gen = _feed(S, B, rate, eng.flat, zDim)
    eng.fp.load(P)
    eng.enable_training()
    E.adopt(eng)
    E.poison(eng)
    eng.set_inputs(x)
    eng.set_noise(eps)
    eng.alpha.copy_(torch.from_numpy(alpha.reshape(-1)))
    eng.mask_mu.copy_(torch.from_numpy(m_z[0]))
    eng.mask_ls.copy_(torch.from_numpy(m_z[1]))
    eng.mask_gen.copy_(torch.from_numpy(m_gen))
    masks = {'mu': m_z[0], 'ls': m_z[1], 'dec': m_gen}
    return eng, P, x, eps, alpha, masks


@pytest.mark.parametrize('kl_weight', [1.0, 0.25])
@pytest.mark.parametrize('which', ['vae', 'gen', 'disc'])
def test_anovaegan_train_ops_match_oracle(which, kl_weight, monkeypatch):
    rate, lr = 0.2, 1e-3
    eng, P, x, eps, alpha, masks = _anovaegan(monkeypatch, rate=rate, kl_weight=kl_weight)
    tr = AO.Trainer(P, lr=lr, dropout_rate=rate, scale=10.0, kl_weight=kl_weight, dtype=torch.float64)
    sg, l1_sign = _anovaegan_signs(eng, which, True, 1.0 / (1.0 - rate))
    out, G = tr.step(which, x, eps, alpha, masks, signs=sg, l1_sign=l1_sign)
    step = {'vae': eng.step_vae, 'gen': eng.step_gen, 'disc': eng.step_disc}[which]
    before = eng.fp.to_numpy()
    res = step(lr, dropout_rate=rate, dropout=True, parity_noise=True)
    for k, v in res.items():
        if k in out:
            assert abs(v - float(out[k])) <= 1e-5 * max(abs(float(out[k])), 1e-3), (k, v, float(out[k]))
    assert _rel(eng.x_gen.numpy(), out['out'].numpy()) < TOL
    assert _rel(eng.mu.numpy(), out['z_mu'].numpy()) < TOL and _rel(eng.sigma.numpy(), out['z_sigma'].numpy()) < TOL
    if which == 'vae':
        assert _rel(eng.l1.numpy(), out['L1'].numpy()) < TOL
    if which == 'disc':
        assert _rel(eng.ddx.numpy(), out['ddx'].numpy()) < TOL
    _compare_grads(eng, G)
    _check_update(eng.fp.to_numpy(), before, tr.P, G, anovaegan_engine.OPS[which], lr)


def test_anovaegan_optimisers_own_their_adam_slots(monkeypatch):
    """One mini-batch of AnoVAEGAN.train (optim_vae, optim_gen, 2 x optim_dis), twice: weights track the oracle trainer, whose
    optim_gen keeps Adam moments for the Generator separate from optim_vae's (TensorFlow creates slots per optimizer)."""
    rate, lr = 0.0, 1e-3
    eng, P, x, eps, alpha, masks = _anovaegan(monkeypatch, rate=rate, zDim=32)
    tr = AO.Trainer(P, lr=lr, dropout_rate=rate, scale=10.0, dtype=torch.float64)
    for it in range(2):
        for which in ('vae', 'gen', 'disc', 'disc'):
            sg, l1_sign = _anovaegan_signs(eng, which, False, 1.0)
            tr.step(which, x, eps, alpha, None, signs=sg, l1_sign=l1_sign)
            {'vae': eng.step_vae, 'gen': eng.step_gen, 'disc': eng.step_disc}[which](lr, dropout_rate=rate, dropout=True,
                                                                                     parity_noise=True)
    assert eng.t == {'vae': 2, 'gen': 2, 'disc': 4}
    assert [int(eng.steps[k][0]) for k in ('vae', 'gen', 'disc')] == [2, 2, 4]
    after = eng.fp.to_numpy()
    off = tot = 0
    for k in after:                     # Adam steps are ~lr*sign(g) early on: elements whose float32 gradient is round-off may differ by O(lr)
        d = np.abs(after[k] - tr.P[k].numpy().reshape(after[k].shape))
        assert float(d.max()) < 4.5 * lr, (k, float(d.max()))
        off += int((d > 0.05 * lr).sum())
        tot += d.size
    assert off < 2e-3 * tot, (off, tot)
    lo, hi = eng.op_range('gen')
    m_vae = eng.fp.m[lo:hi]
    assert float(eng.m_gen.abs().max()) > 0 and not torch.equal(eng.m_gen, m_vae)
    m_ref = tr.slots['gen']['m']
    got = eng.fp.to_numpy(torch.cat([torch.zeros(lo), eng.m_gen, torch.zeros(eng.fp.numel - hi)]))
    mmax = max(float(v.abs().max()) for v in m_ref.values())
    for k, v in m_ref.items():
        if float(v.abs().max()) < 1e-9 * mmax:          # biases feeding a LayerNorm: exactly 0 here, round-off in the oracle
            assert float(np.abs(got[k]).max()) == 0.0, k
            continue
        assert _rel(got[k], v.numpy()) < 5e-2, k    # (trajectories drift by O(lr) elements; shared slots would be off by orders of magnitude)


def test_anovaegan_validation_fetch_and_noise(monkeypatch):
    """train=False evaluates the validation fetches without touching weights; perf-mode noise draws eps / masks / alpha."""
    eng, P, x, eps, alpha, masks = _anovaegan(monkeypatch, rate=0.2)
    before = eng.fp.to_numpy()
    res = eng.step_vae(1e-3, dropout_rate=0.2, dropout=False, train=False, parity_noise=True)
    o = {k: v.detach() for k, v in AO.graph(AO.as_leaves(P, torch.float64), x, eps, dropout_rate=0.2, training=False,
                                            dtype=torch.float64, want=("vae",)).items()}
    assert abs(res['reconstructionLoss'] - float(o['reconstructionLoss'])) < 1e-5 * float(o['reconstructionLoss'])
    assert abs(res['kl'] - float(o['kl'])) < 1e-5 * float(o['kl'])
    assert all(np.array_equal(before[k], v) for k, v in eng.fp.to_numpy().items())
    e0, m0, a0 = eng.eps.clone(), eng.mask_gen.clone(), eng.alpha.clone()
    eng.step_disc(1e-3, dropout_rate=0.2, dropout=True)
    assert not torch.equal(e0, eng.eps) and not torch.equal(m0, eng.mask_gen) and not torch.equal(a0, eng.alpha)
    assert int(eng.rng_ctr[0]) == 1 << 20
    with pytest.raises(NotImplementedError):
        eng.step_enc(1e-3)


def test_anovaegan_trainer_loop(monkeypatch, tmp_path):
    """trainers/AnoVAEGAN.train on the synthetic dataset, engine on the emulator: per mini-batch 1 optim_vae + 1 optim_gen +
    5 optim_dis, validation pass, checkpoint written; the model / trainer keep the reference's protocol."""
    from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import SYNTHETIC
    from unsupervised_anomaly_detection_brain_mri_b200.models.anovaegan import anovaegan
    from unsupervised_anomaly_detection_brain_mri_b200.models.customlayers import Placeholder
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.AnoVAEGAN import AnoVAEGAN
    E.install(monkeypatch, fanogan_engine, anovaegan_engine)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda d: None)
    config = AnoVAEGAN.Config()
    assert (config.modelname, config.scale, config.kappa, config.kl_weight) == ('AnoVAEGAN', 10.0, 1.0, 1.0)
    config.outputHeight = config.outputWidth = 32
    config.batchsize, config.numEpochs, config.zDim, config.numChannels = 2, 1, 16, 1
    config.intermediateResolutions = [8, 8]
    config.dropout_rate, config.learningrate = 0.1, 1e-4
    config.checkpointDir = str(tmp_path / 'ckpt')
    config.description, config.dataset = 'emulated', 'SYNTHETIC'
    config.device, config.math_mode, config.useCudaGraph, config.useTensorboard, config.verbose = 'cpu', 0, False, False, False
    outs = anovaegan(Placeholder([None, 32, 32, 1]), 0.1, False, config)
    assert set(outs) == {'z_mu', 'z_log_sigma', 'z_sigma', 'out', 'd_fake_features', 'd_', 'd_features', 'd', 'x_hat', 'd_hat_features',
                         'd_hat'}
    opts = SYNTHETIC.Options()
    opts.sliceResolution = (32, 32)
    opts.numPatients = 1
    opts.sliceStart, opts.sliceEnd = 20, 28
    ds = SYNTHETIC(opts)
    model = AnoVAEGAN(None, config, network=anovaegan)
    assert model.network.__name__ == 'anovaegan' and 'AnoVAEGAN' in model.model_dir
    model.engine.enable_training()
    E.adopt(model.engine)
    w0 = model.engine.fp.to_numpy()
    model.train(ds)
    w1 = model.engine.fp.to_numpy()
    for scope in ('Encoder', 'Generator', 'Discriminator'):
        assert any(not np.array_equal(w0[k], w1[k]) for k in w0 if k.startswith(scope + '/')), scope
    assert all(np.isfinite(v).all() for v in w1.values())
    t = model.engine.t
    assert t['vae'] == t['gen'] > 0 and t['disc'] == 5 * t['gen']
    assert E.calls.count('uad_randn') == t['vae'] + t['gen'] + t['disc'] + ds.num_batches(2, set='VAL')
    ok, step = model.load(model.checkpointDir)
    assert ok and step == 1
    # reconstruct(): the trainer issues raw ABI calls on data_ptr() integers there - route them through the emulator too
    import types

    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    monkeypatch.setattr(abi, 'call', E.call)
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a, **k: types.SimpleNamespace(cuda_stream=0))
    orig = model._engine_for

    def engine_for(n):
        e = orig(n)
        E.register(e.eps)
        return e
    monkeypatch.setattr(model, '_engine_for', engine_for)
    x = ds.next_batch(2, set='VAL')[0]
    rec = model.reconstruct(x)
    assert rec['reconstruction'].shape == x.shape and np.isfinite(rec['reconstruction']).all() and np.isfinite(rec['l1err'])
    e2 = model._engine_for(2)
    assert float(e2.eps.abs().max()) > 0 and e2.kl_weight == model.engine.kl_weight and e2.fp is model.engine.fp
    one = model.reconstruct(x[0])
    assert one['reconstruction'].shape == (1, 32, 32, 1)


# ------------------------------------------------------------------------------------------------ 3. the AE-family engine
@pytest.mark.parametrize('keep_preact', [False, True])
@pytest.mark.parametrize('arch', [O.VAE, O.CAE])
def test_emulator_reproduces_the_gpu_verified_autoencoder_steps(arch, keep_preact, monkeypatch):
    """Calibration on engine.ConvAutoencoderEngine (VAE: fused final-1x1 backward, reparameterisation; constrained AE: the split
    backward_constrained -> backward_from_gxhat), both backward variants (from z, from the block output)."""
    from unsupervised_anomaly_detection_brain_mri_b200 import engine as eng_mod
    E.install(monkeypatch, eng_mod)
    S, B, rate, lr = 32, 2, 0.2, 1e-3
    P = O.perturb_params(O.init_params(arch, S, seed=1))
    x = O.synthetic_slices(B, S, seed=31)
    eng = eng_mod.ConvAutoencoderEngine(arch, S, batch=B, device='cpu', math_mode=0, keep_preact=keep_preact)
    E.adopt(eng)
    eng.fp.load(P)
    E.poison(eng)
    rng = np.random.default_rng(8)
    mk = lambda n: (rng.uniform(size=(B, n)) >= rate).astype(np.float32)   # noqa: E731
    eng.set_inputs(x)
    if arch == O.CAE:
        om = {'z': mk(128), 'dec': mk(eng.flat), 'z_rec': mk(128)}
        eng.set_noise(None, {'mu': om['z'], 'dec': om['dec']}, {'mu': om['z_rec']})
        kw = dict(masks=om, rho=1.0)
    else:
        om = {'mu': mk(128), 'ls': mk(128), 'dec': mk(eng.flat)}
        eps = rng.standard_normal((B, 128)).astype(np.float32)
        eng.set_noise(eps, om)
        eng._keep = 1.0 / (1.0 - rate)
        eng.forward(training=True, dropout_rate=rate)
        kw = dict(masks={'mu': om['mu'], 'log_sigma': om['ls'], 'dec': om['dec']}, eps=eps, l1_sign=np.sign(eng.br[0].xhat.numpy() - x))
    eng.train_step(lr, beta1=0.5, dropout_rate=rate, dropout=True, parity_noise=True)
    out, L, G = O.loss_and_grads(arch, P, x, dropout_rate=rate, training=True, dtype=torch.float64, **kw)
    assert _rel(eng.br[0].xhat.numpy(), out['x_hat'].numpy()) < TOL
    got = eng.losses()
    for k in got:
        assert abs(got[k] - float(L[k])) <= 1e-5 * abs(float(L[k])), (k, got[k], float(L[k]))
    grads = eng.fp.to_numpy(eng.fp.grads)
    for k in P:
        assert _rel(grads[k], G[k].numpy()) < 2e-5, (k, _rel(grads[k], G[k].numpy()))


# ------------------------------------------------------------------------------------------------ 4. adversarial autoencoder
def _aae(monkeypatch, S=32, B=2, rate=0.2, zDim=32, constrained=False, rho=1.0):
    from oracle import aae_cpu as AA
    from unsupervised_anomaly_detection_brain_mri_b200 import aae_engine
    from unsupervised_anomaly_detection_brain_mri_b200 import engine as eng_mod
    E.install(monkeypatch, eng_mod, aae_engine)
    P = AA.perturb(AA.init_params(S, zDim=zDim, seed=1, constrained=constrained))
    eng = aae_engine.AdversarialAEEngine(S, zDim=zDim, batch=B, device='cpu', math_mode=0, scale=10.0, constrained=constrained, rho=rho)
    E.adopt(eng)
    assert list(eng.specs) == list(P) and all(tuple(eng.specs[k]) == P[k].shape for k in P)
    eng.fp.load(P)
    E.poison(eng)
    rng = np.random.default_rng(5)
    x = O.synthetic_slices(B, S, seed=31)
    z = rng.standard_normal((B, zDim)).astype(np.float32)
    epsilon = rng.random((B, 1), dtype=np.float32)
    masks = {'z': (rng.uniform(size=(B, zDim)) >= rate).astype(np.float32), 'dec': (rng.uniform(size=(B, eng.flat)) >= rate).astype(np.float32)}
    eng.set_inputs(x)
    eng.set_latent(z)
    eng.set_epsilon(epsilon)
    eng.set_noise(None, {'mu': masks['z']} if constrained else {'mu': masks['z'], 'dec': masks['dec']})
    return AA, eng, P, x, z, epsilon, masks


def _aae_signs(eng, which, rate):
    eng._keep = 1.0 / (1.0 - rate) if rate > 0 else 1.0
    z_ = eng.encode_latent()
    sg = {}
    if which in ('gen', 'disc'):
        eng.critic_forward(z_)
        sg['d_fake'] = _pat(eng.cp.pre)
    if which == 'disc':
        eng.critic_forward(eng.z_real)
        sg['d_real'] = _pat(eng.cp.pre)
        E.call('uad_interpolate', eng.z_real, z_, eng.epsilon, eng.z_hat, eng.B, eng.zDim, 0)
        eng.critic_forward(eng.z_hat)
        sg['d_hat'] = _pat(eng.cp.pre)
    return sg


@pytest.mark.parametrize('constrained', [False, True])
@pytest.mark.parametrize('which', ['ae', 'disc', 'gen'])
def test_aae_train_ops_match_oracle(which, constrained, monkeypatch):
    """constrained=True: models/constrained_adversarial_autoencoder.py + trainers/ConstrainedAAE.py (two-pass graph, rho * Rec_z,
    100-50-1 critic, optim_gen over Encoder/* + the 1x1 bottleneck conv + the latent Dense)."""
    rate, lr, rho = 0.2, 1e-3, 0.7
    AA, eng, P, x, z, epsilon, masks = _aae(monkeypatch, rate=rate, constrained=constrained, rho=rho)
    tr = AA.Trainer(P, lr=lr, dropout_rate=rate, scale=10.0, dtype=torch.float64, constrained=constrained, rho=rho)
    out, G = tr.step(which, x, z, epsilon, masks, signs=_aae_signs(eng, which, rate))
    before = eng.fp.to_numpy()
    res = {'ae': eng.step_ae, 'disc': eng.step_disc, 'gen': eng.step_gen}[which](lr, dropout_rate=rate, dropout=True, parity_noise=True)
    for k, v in res.items():
        if k in out and out[k].ndim == 0:
            assert abs(v - float(out[k])) <= 1e-5 * max(abs(float(out[k])), 1e-3), (k, v, float(out[k]))
    assert _rel(eng.br[0].mu.numpy(), out['z_'].numpy()) < TOL
    if which == 'ae':
        assert _rel(eng.br[0].xhat.numpy(), out['x_hat'].numpy()) < TOL
        if constrained:
            assert _rel(eng.br[1].mu.numpy(), out['z_rec'].numpy()) < TOL and 'Rec_z' in res
    if which == 'disc':
        assert _rel(eng.z_hat.numpy(), out['z_hat'].detach().numpy()) < TOL
        assert _rel(eng.ddz.numpy(), out['ddz'].numpy()) < TOL
    got = eng.fp.to_numpy(eng.fp.grads)
    for k, v in G.items():
        assert _rel(got[k], v.numpy()) < 2e-5, (k, _rel(got[k], v.numpy()))
    if constrained and which == 'gen':            # updated: Encoder/* + Bottleneck/conv2d + Bottleneck/dense, nothing else
        assert set(G) == {k for k in P if k.startswith('Encoder/') or k.rsplit('/', 1)[0] in ('Bottleneck/conv2d', 'Bottleneck/dense')}
        after = eng.fp.to_numpy()
        gmax = max(float(v.abs().max()) for v in G.values())
        for k in after:
            if k not in G:
                assert np.array_equal(after[k], before[k]), k
            else:
                sel = np.abs(G[k].numpy()) > 1e-3 * gmax
                assert (np.abs(after[k] - tr.P[k].numpy().reshape(after[k].shape))[sel] <= 0.02 * lr).all(), k
                assert not np.array_equal(after[k], before[k]), k
    else:
        _check_update(eng.fp.to_numpy(), before, tr.P, G, {'ae': ('Encoder', 'Bottleneck', 'Decoder'), 'disc': ('Discriminator',),
                                                           'gen': ('Encoder',)}[which], lr)


def test_aae_optimisers_and_validation(monkeypatch):
    """A few rounds of (optim_ae, 2 x optim_dis, optim_gen) track the oracle trainer: optim_gen's Adam moments for the Encoder are
    its own; the validation fetch changes nothing; perf-mode noise refreshes masks and epsilon."""
    lr = 1e-3
    AA, eng, P, x, z, epsilon, masks = _aae(monkeypatch, rate=0.0)
    tr = AA.Trainer(P, lr=lr, dropout_rate=0.0, scale=10.0, dtype=torch.float64)
    for it in range(2):
        for which in ('ae', 'disc', 'disc', 'gen'):
            tr.step(which, x, z, epsilon, None, signs=_aae_signs(eng, which, 0.0))
            {'ae': eng.step_ae, 'disc': eng.step_disc, 'gen': eng.step_gen}[which](lr, dropout_rate=0.0, dropout=True, parity_noise=True)
    assert eng.op_t == {'ae': 2, 'disc': 4, 'gen': 2}
    after = eng.fp.to_numpy()
    off = tot = 0
    for k in after:
        d = np.abs(after[k] - tr.P[k].numpy().reshape(after[k].shape))
        assert float(d.max()) < 4.5 * lr, (k, float(d.max()))
        off += int((d > 0.05 * lr).sum())
        tot += d.size
    assert off < 2e-3 * tot, (off, tot)
    lo, hi = eng.rng['gen']
    assert float(eng.m_gen.abs().max()) > 0 and not torch.equal(eng.m_gen, eng.fp.m[lo:hi])
    got = eng.fp.to_numpy(torch.cat([torch.zeros(lo), eng.m_gen, torch.zeros(eng.fp.numel - hi)]))
    for k, v in tr.slots['gen']['m'].items():
        assert _rel(got[k], v.numpy()) < 5e-2, k
    before = eng.fp.to_numpy()
    res = eng.step_ae(lr, dropout_rate=0.2, dropout=False, train=False)
    assert all(np.array_equal(before[k], v) for k, v in eng.fp.to_numpy().items()) and res['loss'] > 0
    e0 = eng.epsilon.clone()
    eng.step_disc(lr, dropout_rate=0.2, dropout=True)
    assert not torch.equal(e0, eng.epsilon) and float(eng.epsilon.max()) <= 0.0 and eng.br[0].masks['mu'] is not None


def test_aae_trainer_loop(monkeypatch, tmp_path):
    """trainers/AAE.train on the synthetic dataset through the emulator: d_iters optim_ae + d_iters optim_dis + 1 optim_gen per
    mini-batch (epoch <= 5), validation with all losses, checkpoint; model / trainer protocol of the reference."""
    from unsupervised_anomaly_detection_brain_mri_b200 import aae_engine
    from unsupervised_anomaly_detection_brain_mri_b200 import engine as eng_mod
    from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import SYNTHETIC
    from unsupervised_anomaly_detection_brain_mri_b200.models.adversarial_autoencoder import adversarial_autoencoder
    from unsupervised_anomaly_detection_brain_mri_b200.models.customlayers import Placeholder
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.AAE import AAE
    E.install(monkeypatch, eng_mod, aae_engine)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda d: None)
    config = AAE.Config()
    assert (config.modelname, config.scale) == ('AAE', 10.0)
    config.outputHeight = config.outputWidth = 32
    config.batchsize, config.numEpochs, config.zDim, config.numChannels = 2, 1, 16, 1
    config.intermediateResolutions = [8, 8]
    config.dropout_rate, config.learningrate, config.d_iters = 0.1, 1e-4, 3
    config.checkpointDir = str(tmp_path / 'ckpt')
    config.description, config.dataset = 'emulated', 'SYNTHETIC'
    config.device, config.math_mode, config.useCudaGraph, config.useTensorboard, config.verbose = 'cpu', 0, False, False, False
    outs = adversarial_autoencoder(Placeholder([None, 16]), Placeholder([None, 32, 32, 1]), 0.1, False, config)
    assert set(outs) == {'z_', 'x_hat', 'd_', 'd', 'z_hat', 'd_hat'}
    opts = SYNTHETIC.Options()
    opts.sliceResolution = (32, 32)
    opts.numPatients = 1
    opts.sliceStart, opts.sliceEnd = 20, 28
    ds = SYNTHETIC(opts)
    np.random.seed(0)
    model = AAE(None, config, network=adversarial_autoencoder)
    E.adopt(model.engine)
    w0 = model.engine.fp.to_numpy()
    model.train(ds)
    w1 = model.engine.fp.to_numpy()
    for scope in ('Encoder', 'Bottleneck', 'Decoder', 'Discriminator'):
        assert any(not np.array_equal(w0[k], w1[k]) for k in w0 if k.startswith(scope + '/')), scope
    assert all(np.isfinite(v).all() for v in w1.values())
    t = model.engine.op_t
    assert t['gen'] > 0 and t['ae'] == 3 * t['gen'] and t['disc'] == 3 * t['gen']
    ok, step = model.load(model.checkpointDir)
    assert ok and step == 1


# ------------------------------------------------------------------------------------------------ 5. context encoder (CE)
def test_context_encoder_step_scores_against_the_plain_batch(monkeypatch):
    """trainers/CE.py:21,34: the AE graph runs on the masked batch, the L1 loss is taken against the plain batch - the engine's
    reconstruction target decoupled from its input.  Loss and every gradient vs the oracle; set_target(None) restores target == input."""
    from collections import OrderedDict

    from unsupervised_anomaly_detection_brain_mri_b200 import engine as eng_mod
    E.install(monkeypatch, eng_mod)
    S, B, rate, lr = 32, 2, 0.2, 1e-3
    P = O.perturb_params(O.init_params(O.AE, S, seed=1))
    x = O.synthetic_slices(B, S, seed=31)
    x_ce = x.copy()
    x_ce[:, 8:20, 10:22] = 0
    eng = eng_mod.ConvAutoencoderEngine(O.AE, S, batch=B, device='cpu', math_mode=0)
    E.adopt(eng)
    eng.fp.load(P)
    E.poison(eng)
    mz = (np.random.default_rng(8).uniform(size=(B, 128)) >= rate).astype(np.float32)
    eng.set_inputs(x_ce)
    eng.set_target(x)
    eng.set_noise(None, {'mu': mz})
    eng._keep = 1.0 / (1.0 - rate)
    eng.forward(training=True, dropout_rate=rate)
    l1_sign = np.sign(eng.br[0].xhat.numpy() - x)
    eng.train_step(lr, beta1=0.5, dropout_rate=rate, dropout=True, parity_noise=True)
    Pt = OrderedDict((k, torch.from_numpy(v).double().requires_grad_(True)) for k, v in P.items())
    out = O.forward(O.AE, Pt, x_ce, masks={'z': mz}, dropout_rate=rate, training=True, dtype=torch.float64)
    L = O.losses(O.AE, out, x, dtype=torch.float64, l1_sign=l1_sign)
    G = torch.autograd.grad(L['loss'], list(Pt.values()))
    assert abs(eng.losses()['loss'] - float(L['loss'].detach())) < 1e-5 * float(L['loss'].detach())
    assert _rel(eng.br[0].l1.numpy(), (out['x_hat'].detach() - torch.from_numpy(x).double()).abs().numpy()) < TOL
    grads = eng.fp.to_numpy(eng.fp.grads)
    for k, g in zip(Pt, G):
        assert _rel(grads[k], g.numpy()) < 2e-5, k
    eng.set_target(None)
    eng.forward(training=False)
    assert _rel(eng.br[0].l1.numpy(), np.abs(eng.br[0].xhat.numpy() - x_ce)) < 1e-6


def test_context_encoder_trainer_loop(monkeypatch, tmp_path):
    from unsupervised_anomaly_detection_brain_mri_b200 import engine as eng_mod
    from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import SYNTHETIC
    from unsupervised_anomaly_detection_brain_mri_b200.models.autoencoder import autoencoder
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.AEMODEL import AEMODEL
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.CE import CE
    E.install(monkeypatch, eng_mod)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda d: None)
    monkeypatch.setattr(AEMODEL, '_stage', lambda self, key, arr: torch.from_numpy(np.ascontiguousarray(arr, np.float32)))
    config = CE.Config()
    assert config.modelname == 'CE'
    config.outputHeight = config.outputWidth = 64
    config.batchsize, config.numEpochs, config.zDim, config.numChannels = 2, 1, 16, 1
    config.intermediateResolutions = [8, 8]
    config.dropout_rate, config.learningrate, config.optimizer = 0.1, 1e-4, 'ADAM'
    config.checkpointDir = str(tmp_path / 'ckpt')
    config.description, config.dataset = 'emulated', 'SYNTHETIC'
    config.device, config.math_mode, config.use_cuda_graph, config.useTensorboard, config.verbose = 'cpu', 0, False, False, False
    opts = SYNTHETIC.Options()
    opts.sliceResolution = (64, 64)
    opts.numPatients = 1
    opts.sliceStart, opts.sliceEnd = 40, 48
    ds = SYNTHETIC(opts)
    model = CE(None, config, network=autoencoder)
    E.adopt(model.engine)
    seen = []
    orig = model.engine.set_inputs
    monkeypatch.setattr(model.engine, 'set_inputs', lambda x, x_ce=None: (seen.append(x.clone()), orig(x, x_ce))[1])
    w0 = model.engine.fp.to_numpy()
    import random
    random.seed(0)
    model.train(ds)
    assert any(not np.array_equal(w0[k], v) for k, v in model.engine.fp.to_numpy().items())
    assert model.engine.br[0].target is not None and model.engine.t == ds.num_batches(2, set='TRAIN')
    # TRAIN batches went in masked (exact zeros inside the brain), the target stayed the plain batch
    n_train = ds.num_batches(2, set='TRAIN')
    assert any(float((s == 0).float().mean()) > float((model.engine.br[0].target == 0).float().mean()) for s in seen[:n_train])


def test_constrained_aae_trainer_loop(monkeypatch, tmp_path):
    from unsupervised_anomaly_detection_brain_mri_b200 import aae_engine
    from unsupervised_anomaly_detection_brain_mri_b200 import engine as eng_mod
    from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import SYNTHETIC
    from unsupervised_anomaly_detection_brain_mri_b200.models.constrained_adversarial_autoencoder import constrained_adversarial_autoencoder
    from unsupervised_anomaly_detection_brain_mri_b200.models.customlayers import Placeholder
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.ConstrainedAAE import ConstrainedAAE
    E.install(monkeypatch, eng_mod, aae_engine)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda d: None)
    config = ConstrainedAAE.Config()
    assert (config.modelname, config.rho) == ('ConstrainedAAE', 1)
    config.outputHeight = config.outputWidth = 32
    config.batchsize, config.numEpochs, config.zDim, config.numChannels = 2, 1, 16, 1
    config.intermediateResolutions = [8, 8]
    config.dropout_rate, config.learningrate, config.d_iters, config.rho, config.scale = 0.1, 1e-4, 2, 0.5, 1
    config.checkpointDir = str(tmp_path / 'ckpt')
    config.description, config.dataset = 'emulated', 'SYNTHETIC'
    config.device, config.math_mode, config.useCudaGraph, config.useTensorboard, config.verbose = 'cpu', 0, False, False, False
    outs = constrained_adversarial_autoencoder(Placeholder([None, 16]), Placeholder([None, 32, 32, 1]), 0.1, False, config)
    assert set(outs) == {'z_', 'x_hat', 'z_rec', 'd_', 'd', 'z_hat', 'd_hat'}
    opts = SYNTHETIC.Options()
    opts.sliceResolution = (32, 32)
    opts.numPatients = 1
    opts.sliceStart, opts.sliceEnd = 20, 28
    ds = SYNTHETIC(opts)
    np.random.seed(0)
    model = ConstrainedAAE(None, config, network=constrained_adversarial_autoencoder)
    eng = model.engine
    assert eng.constrained and eng.rho == 0.5 and eng.scale == 1.0 and eng.widths == (100, 50, 1)
    assert eng.specs['Discriminator/dense_2/kernel'] == (16, 100)
    E.adopt(eng)
    w0 = eng.fp.to_numpy()
    model.train(ds)
    w1 = eng.fp.to_numpy()
    for scope in ('Encoder', 'Bottleneck', 'Decoder', 'Discriminator'):
        assert any(not np.array_equal(w0[k], w1[k]) for k in w0 if k.startswith(scope + '/')), scope
    assert all(np.isfinite(v).all() for v in w1.values())
    assert eng.op_t['gen'] > 0 and eng.op_t['ae'] == 2 * eng.op_t['gen'] == eng.op_t['disc']
    assert eng.br[1].masks['mu'] is None and eng.br[0].masks['dec'] is None          # the two Dropout calls without the flag


@pytest.mark.parametrize('arch', [O.AE, O.AES, O.CEVAE])
def test_emulator_reproduces_the_other_gpu_verified_steps(arch, monkeypatch):
    """Calibration / regression guard for the remaining AE-family graphs (dense AE, spatial AE, ceVAE with the input-gradient
    anomaly map): engine.py was edited after its last GPU run (new archs, backward split, decoupled target) - the call sequences
    the GPU suite verified must still reproduce the oracle."""
    from unsupervised_anomaly_detection_brain_mri_b200 import engine as eng_mod
    E.install(monkeypatch, eng_mod)
    S, B, rate, lr = 32, 2, 0.2, 1e-3
    P = O.perturb_params(O.init_params(arch, S, seed=1))
    x = O.synthetic_slices(B, S, seed=1234)
    x_ce = x.copy()
    x_ce[:, 8:20, 10:22] = 0
    eng = eng_mod.ConvAutoencoderEngine(arch, S, batch=B, device='cpu', math_mode=0)
    E.adopt(eng)
    assert list(eng.specs.keys()) == list(P.keys())
    eng.fp.load(P)
    E.poison(eng)
    rng = np.random.default_rng(3)
    eps = rng.standard_normal((B, 128)).astype(np.float32)
    mk = lambda *n: (rng.uniform(size=(B,) + n) >= rate).astype(np.float32)   # noqa: E731
    emc = None
    if arch == O.AE:
        om = {'z': mk(128)}
        em = {'mu': om['z']}
    elif arch == O.AES:
        om = {'z': mk(8, 8, eng.enc_ch[-1])}
        em = {'sp': om['z']}
    else:
        om = {'mu': mk(128), 'log_sigma': mk(128), 'dec': mk(eng.flat), 'mu_ce': mk(128), 'dec_ce': mk(eng.flat)}
        em, emc = {'mu': om['mu'], 'ls': om['log_sigma'], 'dec': om['dec']}, {'mu': om['mu_ce'], 'dec': om['dec_ce']}
    ce = arch == O.CEVAE
    eng.set_inputs(x, x_ce if ce else None)
    eng.set_noise(eps, em, emc)
    eng._keep = 1.0 / (1.0 - rate)
    eng.forward(training=True, dropout_rate=rate)
    sgn = np.sign(eng.br[0].xhat.numpy().astype(np.float64) - x)
    sgn_ce = np.sign(eng.br[1].xhat.numpy().astype(np.float64) - x_ce) if ce else None
    eng.train_step(lr, beta1=0.5, dropout_rate=rate, dropout=True, parity_noise=True, want_anomaly=ce)
    out, L, G = O.loss_and_grads(arch, P, x, x_ce=x_ce, eps=eps, masks=om, dropout_rate=rate, training=True, dtype=torch.float64,
                                 want_anomaly=ce, l1_sign=sgn, l1_sign_ce=sgn_ce)
    assert _rel(eng.br[0].xhat.numpy(), out['x_hat'].numpy()) < TOL
    got = eng.losses()
    for k in got:
        assert abs(got[k] - float(L[k])) <= 1e-5 * abs(float(L[k])), (k, got[k], float(L[k]))
    if ce:
        assert _rel(eng.br[1].xhat.numpy(), out['x_hat_ce'].numpy()) < TOL
        assert _rel(eng.anomaly.numpy(), L['anomaly'].numpy()) < 2e-5
    grads = eng.fp.to_numpy(eng.fp.grads)
    for k in P:
        assert _rel(grads[k], G[k].numpy()) < 2e-5, (k, _rel(grads[k], G[k].numpy()))
    # inference forward, dropout off (what reconstruct() runs)
    eng.fp.load(P)
    for br in eng.br:
        br.masks = {k: None for k in br.masks}
    eng.forward(training=False, dropout_rate=0.0, branches=[0], need_l1=False)
    out2 = O.forward(arch, P, x, x_ce=x_ce, eps=eps, training=False, dtype=torch.float64)
    assert _rel(eng.br[0].xhat.numpy(), out2['x_hat'].numpy()) < TOL


# ------------------------------------------------------------------------------------------------ 6. Gaussian-mixture VAE
def _gmvae(monkeypatch, S=32, B=2, rate=0.2, dz=16, dw=2, dc=5, c_lambda=0.01):
    from oracle import gmvae_cpu as GO
    from unsupervised_anomaly_detection_brain_mri_b200 import engine as eng_mod
    E.install(monkeypatch, eng_mod)
    P = GO.perturb(GO.init_params(S, dim_z=dz, dim_w=dw, dim_c=dc, seed=1))
    eng = eng_mod.ConvAutoencoderEngine(eng_mod.GMVAE, S, zDim=dz, batch=B, device='cpu', math_mode=0, dim_w=dw, dim_c=dc, c_lambda=c_lambda)
    E.adopt(eng)
    assert list(eng.specs) == list(P) and all(tuple(eng.specs[k]) == P[k].shape for k in P)
    eng.fp.load(P)
    E.poison(eng)
    rng = np.random.default_rng(9)
    x = O.synthetic_slices(B, S, seed=31)
    eps_w, eps_z = rng.standard_normal((B, dw)).astype(np.float32), rng.standard_normal((B, dz)).astype(np.float32)
    mk = lambda n: (rng.uniform(size=(B, n)) >= rate).astype(np.float32)   # noqa: E731
    masks = {'w_mu': mk(dw), 'w_ls': mk(dw), 'z_mu': mk(dz), 'dec': mk(eng.flat)}
    eng.set_inputs(x)
    eng.br[0].eps_w.copy_(torch.from_numpy(eps_w))
    eng.set_noise(eps_z, {'wmu': masks['w_mu'], 'wls': masks['w_ls'], 'mu': masks['z_mu'], 'dec': masks['dec']})
    return GO, eng, P, x, eps_w, eps_z, masks


@pytest.mark.parametrize('c_lambda', [0.01, 100.0])
def test_gmvae_train_step_matches_oracle(c_lambda, monkeypatch):
    """models/gaussian_mixture_variational_autoencoder.py + trainers/GMVAE.py:58-88: forward tensors, the four loss terms and every
    gradient (both sides of the tf.maximum gate of the cluster prior)."""
    rate, lr, dc = 0.2, 1e-3, 5
    GO, eng, P, x, eps_w, eps_z, masks = _gmvae(monkeypatch, rate=rate, dc=dc, c_lambda=c_lambda)
    eng._keep = 1.0 / (1.0 - rate)
    eng.forward(training=True, dropout_rate=rate)
    sgn = np.sign(eng.br[0].xhat.numpy().astype(np.float64) - x)
    eng.train_step(lr, beta1=0.5, dropout_rate=rate, dropout=True, parity_noise=True)
    o, L, G = GO.loss_and_grads(P, x, eps_w, eps_z, masks, rate, True, dc, c_lambda, torch.float64, l1_sign=sgn)
    br = eng.br[0]
    for got, key in ((br.xhat, 'xz_mu'), (br.mu, 'z_mu'), (br.ls, 'z_log_sigma'), (br.zv, 'z_sampled'), (br.w_s, 'w_sampled'), (br.pc, 'pc')):
        assert _rel(got.numpy().reshape(o[key].shape), o[key].numpy()) < TOL, key
    assert _rel(br.Mz.numpy().reshape(o['z_wc_mus'].shape), o['z_wc_mus'].numpy()) < TOL
    assert _rel(br.Sz.numpy().reshape(o['z_wc_log_sigma_invs'].shape), o['z_wc_log_sigma_invs'].numpy()) < TOL
    got = eng.losses()
    for k in got:
        assert abs(got[k] - float(L[k])) <= 1e-5 * max(abs(float(L[k])), 1e-6), (k, got[k], float(L[k]))
    assert (float(L['c_prior_loss']) == c_lambda) == (c_lambda == 100.0)
    grads = eng.fp.to_numpy(eng.fp.grads)
    gmax = max(float(v.abs().max()) for v in G.values())
    for k in P:
        ref = G[k].numpy()
        if float(np.abs(ref).max()) < 1e-12 * gmax:
            assert float(np.abs(grads[k]).max()) < 1e-9 * gmax, k
            continue
        assert _rel(grads[k], ref) < 2e-5, (k, _rel(grads[k], ref))
    assert np.array_equal(grads['Variable'], grads['dense_6/bias'])


def test_gmvae_restoration_step_matches_oracle(monkeypatch):
    """GMVAE.reconstruct (trainers/GMVAE.py:166-197): one MAP-restoration iteration, grads = d/dx sum_b [loss + tv_lambda TV(x - xz_mu)]
    with the batch mean of `loss` multiplied back by B (a scalar broadcast over the per-image TV vector)."""
    tv_lambda, lr, dc, c_lambda = 1.3, 1e-3, 5, 0.01
    GO, eng, P, x, eps_w, eps_z, masks = _gmvae(monkeypatch, rate=0.0, dc=dc, c_lambda=c_lambda)
    for br in eng.br:
        br.masks = {k: None for k in br.masks}
    eng.forward(training=False, dropout_rate=0.0, branches=[0], need_l1=False)
    xh = eng.br[0].xhat.numpy().astype(np.float64)
    d = x.astype(np.float64) - xh
    tv_sign = (np.sign(d[:, 1:] - d[:, :-1]), np.sign(d[:, :, 1:] - d[:, :, :-1]))
    # (oracle.restore_gradient sums `loss + tv_lambda * tv` over the batch exactly as tf.gradients does: the scalar batch-mean loss is
    #  broadcast over the per-image TV vector, i.e. multiplied back by B)
    want, o = GO.restore_gradient(P, x, eps_w, eps_z, tv_lambda, dc, c_lambda, torch.float64, l1_sign=np.sign(xh - x), tv_sign=tv_sign)
    x0 = eng.br[0].x.clone()
    eng.restore_step(lr, tv_lambda, parity_noise=True, keep_grads=True)
    assert _rel(eng.restore_grads.numpy(), want.numpy()) < 2e-5
    assert _rel((x0 - eng.br[0].x).numpy(), lr * want.numpy()) < 1e-4


def test_gmvae_trainer_loop_and_restoration(monkeypatch, tmp_path):
    from unsupervised_anomaly_detection_brain_mri_b200 import engine as eng_mod
    from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import SYNTHETIC
    from unsupervised_anomaly_detection_brain_mri_b200.models.customlayers import Placeholder
    from unsupervised_anomaly_detection_brain_mri_b200.models.gaussian_mixture_variational_autoencoder import gaussian_mixture_variational_autoencoder
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.AEMODEL import AEMODEL
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.GMVAE import GMVAE
    E.install(monkeypatch, eng_mod)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda d: None)
    monkeypatch.setattr(AEMODEL, '_stage', lambda self, key, arr: torch.from_numpy(np.ascontiguousarray(arr, np.float32)))
    monkeypatch.setattr(AEMODEL, '_prefetch', lambda self, key, arr: None)
    config = GMVAE.Config()
    assert (config.modelname, config.dim_c, config.dim_z, config.dim_w, config.c_lambda, config.restore_steps) == ('GMVAE', 6, 1, 1, 1, 150)
    config.outputHeight = config.outputWidth = 32
    config.batchsize, config.numEpochs, config.zDim, config.numChannels = 2, 1, 128, 1
    config.dim_c, config.dim_z, config.dim_w, config.c_lambda = 4, 8, 1, 0.5
    config.restore_steps, config.restore_lr, config.tv_lambda = 2, 1e-3, 1.2
    config.intermediateResolutions = [8, 8]
    config.dropout_rate, config.learningrate, config.optimizer = 0.1, 1e-4, 'ADAM'
    config.checkpointDir = str(tmp_path / 'ckpt')
    config.description, config.dataset = 'emulated', 'SYNTHETIC'
    config.device, config.math_mode, config.use_cuda_graph, config.useTensorboard, config.verbose = 'cpu', 0, False, False, False
    outs = gaussian_mixture_variational_autoencoder(Placeholder([None, 32, 32, 1]), 0.1, False, config)
    assert {'w_mu', 'z_mu', 'z_wc_mus', 'xz_mu', 'pc'} <= set(outs) and outs['xz_mu'].graph.zDim == 8
    opts = SYNTHETIC.Options()
    opts.sliceResolution = (32, 32)
    opts.numPatients = 1
    opts.sliceStart, opts.sliceEnd = 20, 28
    ds = SYNTHETIC(opts)
    model = GMVAE(None, config, network=gaussian_mixture_variational_autoencoder)
    eng = model.engine
    assert (eng.arch, eng.zDim, eng.dim_w, eng.dim_c, eng.c_lambda) == (eng_mod.GMVAE, 8, 1, 4, 0.5)
    E.adopt(eng)
    w0 = eng.fp.to_numpy()
    model.train(ds)
    w1 = eng.fp.to_numpy()
    assert all(np.isfinite(v).all() for v in w1.values())
    for name in ('Encoder/enc_conv2D_0/kernel', 'Bottleneck/dense/kernel', 'Bottleneck/dense_3/kernel', 'dense_5/kernel', 'dense_6/kernel',
                 'Variable', 'Decoder/dec_Conv2D_final/kernel'):
        assert not np.array_equal(w0[name], w1[name]), name
    assert eng.t == ds.num_batches(2, set='TRAIN')
    x = ds.next_batch(2, set='VAL')[0]
    orig_eval = model._eval_engine
    monkeypatch.setattr(model, '_eval_engine', lambda n: E.adopt(orig_eval(n)))
    rec = model.reconstruct(x)                       # restored input after 2 iterations
    assert rec['reconstruction'].shape == x.shape and np.isfinite(rec['l1err'])
    assert 0 < np.abs(rec['reconstruction'] - x).max() < 0.1
    model.restore_steps = 0                          # GMVAE.py:170-177: plain reconstruction
    rec0 = model.reconstruct(x)
    assert np.abs(rec0['reconstruction'] - x).max() > np.abs(rec['reconstruction'] - x).max()


def test_fanogan_trainer_loop_after_the_trainer_refactor(monkeypatch, tmp_path):
    """trainers/fAnoGAN.train through the emulator (the trainer grew override hooks for AnoVAEGAN / AAE after its last GPU run):
    WGAN phase (1 generator + 5 critic steps per batch), encoder phase with validation, checkpoints."""
    from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import SYNTHETIC
    from unsupervised_anomaly_detection_brain_mri_b200.models.fanogan import fanogan
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.fAnoGAN import fAnoGAN
    E.install(monkeypatch, fanogan_engine)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda d: None)
    config = fAnoGAN.Config()
    config.outputHeight = config.outputWidth = 32
    config.batchsize, config.numEpochs, config.zDim, config.numChannels = 2, 1, 16, 1
    config.intermediateResolutions = [8, 8]
    config.dropout_rate, config.learningrate = 0.1, 1e-4
    config.checkpointDir = str(tmp_path / 'ckpt')
    config.description, config.dataset = 'emulated', 'SYNTHETIC'
    config.device, config.math_mode, config.useCudaGraph, config.useTensorboard, config.verbose = 'cpu', 0, False, False, False
    opts = SYNTHETIC.Options()
    opts.sliceResolution = (32, 32)
    opts.numPatients = 1
    opts.sliceStart, opts.sliceEnd = 20, 28
    ds = SYNTHETIC(opts)
    np.random.seed(0)
    model = fAnoGAN(None, config, network=fanogan)
    assert model.reconstruction.key == 'x_enc' and model.generated.key == 'x_'
    model.engine.enable_training()
    E.adopt(model.engine)
    w0 = model.engine.fp.to_numpy()
    model.train(ds)
    w1 = model.engine.fp.to_numpy()
    for scope in ('Encoder', 'Generator', 'Discriminator'):
        assert any(not np.array_equal(w0[k], w1[k]) for k in w0 if k.startswith(scope + '/')), scope
    t = model.engine.t
    assert t['Generator'] > 0 and t['Discriminator'] == 5 * t['Generator'] and t['Encoder'] > 0
    e2 = model._engine_for(2)
    assert e2.kappa == model.engine.kappa and e2.fp is model.engine.fp


@pytest.mark.parametrize('arch', [O.VAE, O.AE])
def test_emulator_reproduces_the_gpu_verified_restoration_step(arch, monkeypatch):
    """VAE_You restoration iteration (engine.restore_step -> backward_to_input, edited for the GMVAE branch after its last GPU run)."""
    from unsupervised_anomaly_detection_brain_mri_b200 import engine as eng_mod
    E.install(monkeypatch, eng_mod)
    S, B, lam, lr = 32, 2, 1.8, 1e-3
    P = O.perturb_params(O.init_params(arch, S, seed=1))
    x = O.synthetic_slices(B, S, seed=11)
    eps = np.random.default_rng(4).standard_normal((B, 128)).astype(np.float32)
    eng = eng_mod.ConvAutoencoderEngine(arch, S, batch=B, device='cpu', math_mode=0)
    E.adopt(eng)
    eng.fp.load(P)
    E.poison(eng)
    eng.set_inputs(x)
    eng.set_noise(eps)
    eng.restore_step(lr, lam, parity_noise=True, keep_grads=True)
    xh = eng.br[0].xhat.numpy()
    g_ref, out, tv_ref = O.restore_gradient(arch, P, x, eps=eps, tv_lambda=lam, dtype=torch.float64, sign_from=xh)
    assert _rel(xh, out['x_hat'].numpy()) < TOL and _rel(eng.tv.numpy(), tv_ref.numpy()) < TOL
    assert _rel(eng.restore_grads.numpy(), g_ref.numpy()) < 2e-5


# ------------------------------------------------------------------------------------------------ 7. data-parallel plumbing of the new engines
@pytest.mark.parametrize('family', ['anovaegan', 'aae', 'caae'])
def test_new_engines_data_parallel_equivalence(family, monkeypatch):
    """Two identical ranks: the sum all-reduce of an op's gradient slices (here: x2) with grad_scale = 1 / world must leave exactly
    the single-rank update - on every slice the op's Adam touches (three for the constrained AAE's optim_gen) and nowhere else."""
    lr, rate = 1e-3, 0.2
    results = []
    for world in (1, 2):
        if family == 'anovaegan':
            eng, P, x, eps, alpha, masks = _anovaegan(monkeypatch, rate=rate, zDim=32)
            ops = [('vae', eng.step_vae), ('gen', eng.step_gen), ('disc', eng.step_disc)]
        else:
            AA, eng, P, x, z, epsilon, masks = _aae(monkeypatch, rate=rate, constrained=family == 'caae')
            ops = [('ae', eng.step_ae), ('disc', eng.step_disc), ('gen', eng.step_gen)]
        reduced = []

        def allreduce(t, _log=reduced):
            _log.append(int(t.numel()))
            t.mul_(2.0)
        for name, step in ops:
            step(lr, dropout_rate=rate, dropout=True, parity_noise=True, allreduce=allreduce if world == 2 else None, world=world)
        results.append((eng.fp.to_numpy(), reduced, eng))
    (w1, _, _), (w2, reduced, eng) = results
    for k in w1:
        assert np.allclose(w1[k], w2[k], rtol=0, atol=1e-9), k
    if family == 'anovaegan':
        assert reduced == [hi - lo for lo, hi in (eng.op_range('vae'), eng.op_range('gen'), eng.op_range('disc'))]
    else:
        want = [hi - lo for op in ('ae', 'disc', 'gen') for lo, hi in eng.rngs[op]]
        assert reduced == want and len(want) == (5 if family == 'caae' else 3)


# ------------------------------------------------------------------------------------------------ 8. spatial GMVAE
def _gmvaes(monkeypatch, S=32, B=2, dz=2, dw=1, dc=5, c_lambda=0.01):
    from oracle import gmvae_cpu as GO
    from unsupervised_anomaly_detection_brain_mri_b200 import engine as eng_mod
    E.install(monkeypatch, eng_mod)
    P = GO.perturb(GO.init_params_spatial(S, dim_z=dz, dim_w=dw, dim_c=dc, seed=1))
    eng = eng_mod.ConvAutoencoderEngine(eng_mod.GMVAES, S, zDim=dz, batch=B, device='cpu', math_mode=0, dim_w=dw, dim_c=dc, c_lambda=c_lambda)
    E.adopt(eng)
    assert list(eng.specs) == list(P) and all(tuple(eng.specs[k]) == P[k].shape for k in P)
    eng.fp.load(P)
    E.poison(eng)
    rng = np.random.default_rng(9)
    x = O.synthetic_slices(B, S, seed=31)
    eps_w, eps_z = rng.standard_normal((B, 8, 8, dw)).astype(np.float32), rng.standard_normal((B, 8, 8, dz)).astype(np.float32)
    eng.set_inputs(x)
    eng.br[0].eps_w.copy_(torch.from_numpy(eps_w.reshape(-1, dw)))
    eng.br[0].eps.copy_(torch.from_numpy(eps_z.reshape(-1, dz)))
    return GO, eng, P, x, eps_w, eps_z


@pytest.mark.parametrize('c_lambda', [0.01, 100.0])
def test_gmvae_spatial_train_step_matches_oracle(c_lambda, monkeypatch):
    """models/gaussian_mixture_variational_autoencoder_spatial.py + trainers/GMVAE_spatial.py:58-92 (the decoder runs on the encoder
    output; the prior terms act on the encoder through the 1x1 heads only)."""
    lr, dc = 1e-3, 5
    GO, eng, P, x, eps_w, eps_z = _gmvaes(monkeypatch, dc=dc, c_lambda=c_lambda)
    eng._keep = 1.0
    eng.forward(training=True, dropout_rate=0.0)
    sgn = np.sign(eng.br[0].xhat.numpy().astype(np.float64) - x)
    eng.train_step(lr, beta1=0.5, dropout_rate=0.0, dropout=False, parity_noise=True)
    o, L, G = GO.loss_and_grads_spatial(P, x, eps_w, eps_z, dc, c_lambda, torch.float64, l1_sign=sgn)
    br = eng.br[0]
    for got, key in ((br.xhat, 'xz_mu'), (br.mu, 'z_mu'), (br.zv, 'z_sampled'), (br.w_s, 'w_sampled'), (br.pc, 'pc'), (br.Mz, 'z_wc_mus'),
                     (br.Sz, 'z_wc_log_sigma_invs')):
        assert _rel(got.numpy().reshape(o[key].shape), o[key].numpy()) < TOL, key
    got = eng.losses()
    for k in got:
        assert abs(got[k] - float(L[k])) <= 1e-5 * max(abs(float(L[k])), 1e-6), (k, got[k], float(L[k]))
    grads = eng.fp.to_numpy(eng.fp.grads)
    for k in P:
        assert _rel(grads[k], G[k].numpy()) < 2e-5, (k, _rel(grads[k], G[k].numpy()))


def test_gmvae_spatial_restoration_step_matches_oracle(monkeypatch):
    tv_lambda, lr, dc, c_lambda = 1.3, 1e-3, 5, 0.01
    GO, eng, P, x, eps_w, eps_z = _gmvaes(monkeypatch, dc=dc, c_lambda=c_lambda)
    eng.forward(training=False, dropout_rate=0.0, branches=[0], need_l1=False)
    xh = eng.br[0].xhat.numpy().astype(np.float64)
    d = x.astype(np.float64) - xh
    tv_sign = (np.sign(d[:, 1:] - d[:, :-1]), np.sign(d[:, :, 1:] - d[:, :, :-1]))
    want, _ = GO.restore_gradient_spatial(P, x, eps_w, eps_z, tv_lambda, dc, c_lambda, torch.float64, l1_sign=np.sign(xh - x), tv_sign=tv_sign)
    eng.restore_step(lr, tv_lambda, parity_noise=True, keep_grads=True)
    assert _rel(eng.restore_grads.numpy(), want.numpy()) < 2e-5


def test_gmvae_spatial_trainer_loop(monkeypatch, tmp_path):
    from unsupervised_anomaly_detection_brain_mri_b200 import engine as eng_mod
    from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import SYNTHETIC
    from unsupervised_anomaly_detection_brain_mri_b200.models.gaussian_mixture_variational_autoencoder_spatial import \
        gaussian_mixture_variational_autoencoder_spatial as net
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.AEMODEL import AEMODEL
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.GMVAE_spatial import GMVAE_spatial
    E.install(monkeypatch, eng_mod)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda d: None)
    monkeypatch.setattr(AEMODEL, '_stage', lambda self, key, arr: torch.from_numpy(np.ascontiguousarray(arr, np.float32)))
    monkeypatch.setattr(AEMODEL, '_prefetch', lambda self, key, arr: None)
    config = GMVAE_spatial.Config()
    assert config.modelname == 'GMVAE_spatial'
    config.outputHeight = config.outputWidth = 32
    config.batchsize, config.numEpochs, config.zDim, config.numChannels = 2, 1, 128, 1
    config.dim_c, config.dim_z, config.dim_w, config.c_lambda = 4, 1, 1, 0.5
    config.restore_steps, config.restore_lr, config.tv_lambda = 2, 1e-3, 1.2
    config.intermediateResolutions = [8, 8]
    config.dropout_rate, config.learningrate, config.optimizer = 0.1, 1e-4, 'ADAM'
    config.checkpointDir = str(tmp_path / 'ckpt')
    config.description, config.dataset = 'emulated', 'SYNTHETIC'
    config.device, config.math_mode, config.use_cuda_graph, config.useTensorboard, config.verbose = 'cpu', 0, False, False, False
    opts = SYNTHETIC.Options()
    opts.sliceResolution = (32, 32)
    opts.numPatients = 1
    opts.sliceStart, opts.sliceEnd = 20, 28
    ds = SYNTHETIC(opts)
    model = GMVAE_spatial(None, config, network=net)
    eng = model.engine
    assert (eng.arch, eng.zDim, eng.dim_c) == (eng_mod.GMVAES, 1, 4) and model.network.__name__.endswith('_spatial')
    E.adopt(eng)
    w0 = eng.fp.to_numpy()
    model.train(ds)
    w1 = eng.fp.to_numpy()
    assert all(np.isfinite(v).all() for v in w1.values())
    for name in ('Encoder/enc_conv2D_0/kernel', 'q_wz_x/w_mu/kernel', 'q_wz_x/z_log_sigma/kernel', 'p_z_wc/1x1convlayer/kernel',
                 'p_z_wc/z_wc_mu/kernel', 'Variable', 'Decoder/dec_Conv2D_final/kernel'):
        assert not np.array_equal(w0[name], w1[name]), name
    x = ds.next_batch(2, set='VAL')[0]
    orig_eval = model._eval_engine
    monkeypatch.setattr(model, '_eval_engine', lambda n: E.adopt(orig_eval(n)))
    rec = model.reconstruct(x)
    assert rec['reconstruction'].shape == x.shape and 0 < np.abs(rec['reconstruction'] - x).max() < 0.1


# ------------------------------------------------------------------------------------------------ 9. run.py's call sequence end to end
def _everything_on_the_emulator(monkeypatch):
    """Engines, trainers, Evaluation and Metrics all through the emulator: module-level call / ptr of the engines, the raw
    `abi.call(... data_ptr() ...)` sites of trainers / Evaluation / Metrics, and the few torch.cuda entry points the host code touches."""
    import types

    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    from unsupervised_anomaly_detection_brain_mri_b200 import engine as eng_mod
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.AEMODEL import AEMODEL
    E.install(monkeypatch, eng_mod, fanogan_engine, anovaegan_engine)
    monkeypatch.setattr(abi, 'call', E.call)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda d: None)
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a, **k: types.SimpleNamespace(cuda_stream=0))
    monkeypatch.setattr(AEMODEL, '_stage', lambda self, key, arr: torch.from_numpy(np.ascontiguousarray(arr, np.float32)))
    monkeypatch.setattr(AEMODEL, '_prefetch', lambda self, key, arr: None)


@pytest.mark.parametrize('tname,mname', [('AE', 'autoencoder'), ('VAE', 'variational_autoencoder')])
def test_run_py_call_sequence_on_the_emulator(tname, mname, monkeypatch, tmp_path):
    """The sequence run.py drives (options -> datasets -> config -> Trainer -> train -> checkpoint -> resume -> reconstruct ->
    Evaluation.evaluate), on CPU: the GPU suite's test_gpu_golden equivalent with every kernel emulated - in particular the
    evaluation path (batched reconstruction, erosion / residual / median kernels, device threshold counts, best-Dice search,
    the reference's evalPC result set and files)."""
    import importlib

    from unsupervised_anomaly_detection_brain_mri_b200.utils import Evaluation
    from unsupervised_anomaly_detection_brain_mri_b200.utils.default_config_setup import get_config, get_datasets, get_options
    _everything_on_the_emulator(monkeypatch)
    PKG = 'unsupervised_anomaly_detection_brain_mri_b200'
    trainer = getattr(importlib.import_module(f'{PKG}.trainers.{tname}'), tname)
    network = getattr(importlib.import_module(f'{PKG}.models.{mname}'), mname)
    cfgjson = {'CHECKPOINTDIR': str(tmp_path / 'ckpt'), 'SAMPLEDIR': str(tmp_path / 'samples'), 'BRAINWEBDIR': '', 'SYNTHETICDIR': ''}
    options = get_options(batchsize=4, learningrate=1e-3, numEpochs=1, zDim=128, outputWidth=32, outputHeight=32, slices_start=20,
                          slices_end=32, config=cfgjson)
    options['data']['dir'] = ''
    options['data']['numPatients'] = 2
    options['data']['numTestPatients'] = 2
    hc, pc = get_datasets(options)
    config = get_config(trainer, options, 'ADAM', [8, 8], 0.2, hc)
    config.useTensorboard, config.verbose, config.device, config.math_mode, config.use_cuda_graph = False, False, 'cpu', 0, False
    model = trainer(None, config, network=network)
    E.adopt(model.engine)
    orig_eval = model._eval_engine
    monkeypatch.setattr(model, '_eval_engine', lambda n: E.adopt(orig_eval(n)))
    model.train(hc)
    ckdir = os.path.join(model.checkpointDir, model.model_dir)
    assert f'{config.modelname}.model-1.npz' in os.listdir(ckdir)
    model2 = trainer(None, config, network=network)
    assert model2.load_checkpoint() == 1
    w, w2 = model.engine.fp.to_numpy(), model2.engine.fp.to_numpy()
    assert all(np.array_equal(w[k], w2[k]) for k in w)
    x = hc.next_batch(4, set='VAL')[0]
    r = model.reconstruct(x[0])
    assert r['reconstruction'].shape == (1, 32, 32, 1) and np.isfinite(r['l1err'])
    ev = Evaluation.evaluate(pc, model, options, epoch='1', description='test')
    assert ev['diffs'].shape == (24, 32, 32) and 0.0 <= ev['bestThreshold'] < 1.0 and np.isfinite(ev['diff_AUC'])
    for key in ('DiceScore', 'DiceScorePerPatient', 'TPCC', 'FPCC', 'FNCC', 'TP', 'FP', 'TN', 'FN', 'VD', 'DICE', 'AUC', 'thresholdType'):
        assert key in ev, key
    assert len(ev['DiceScorePerPatient']) == 2 and ev['TP'] + ev['FP'] + ev['TN'] + ev['FN'] == ev['diffs'].size
    assert os.path.isfile(os.path.join(ev['evalDir'], 'evalPC.npy')) and os.path.isfile(os.path.join(ev['evalDir'], 'rocPC.npy'))
    # the device threshold mask is `diffs > t` on the float64 volume (bit-exact on the GPU; the emulator states the same compare)
    thr_mask = Evaluation.Metrics.DeviceScorer(ev['diffs'], ev['labelmaps'] > 0, device='cpu').threshold_mask(ev['threshold'])
    assert np.array_equal(thr_mask.numpy().astype(bool).reshape(ev['diffs'].shape), ev['diffs'] > ev['threshold'])
    best, thr = Evaluation.determine_threshold_on_labeled_patients([pc], model, options, description='VAL')
    assert best == ev['bestDiceScore'] and thr == ev['bestThreshold']        # the synthetic lesion set has no VAL patients: TEST split
    options['threshold'] = float(thr)
    ev2 = Evaluation.evaluate(pc, model, options, epoch='1', description='fixed')
    assert ev2['thresholdType'] == float(thr) and ev2['DICE'] == pytest.approx(ev['DICE'])


def test_fanogan_trainer_reconstruct_and_scoring_on_the_emulator(monkeypatch, tmp_path):
    """tests/test_gpu_golden.py::test_fanogan_trainer_reconstruct_and_scoring on CPU (untrained weights: reconstruct incl. the MC-dropout
    path with its raw ABI calls, checkpoint, single-patient evaluation)."""
    from unsupervised_anomaly_detection_brain_mri_b200.models.fanogan import fanogan
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.fAnoGAN import fAnoGAN
    from unsupervised_anomaly_detection_brain_mri_b200.utils import Evaluation
    from unsupervised_anomaly_detection_brain_mri_b200.utils.default_config_setup import get_config, get_datasets, get_options
    _everything_on_the_emulator(monkeypatch)
    cfgjson = {'CHECKPOINTDIR': str(tmp_path / 'ckpt'), 'SAMPLEDIR': str(tmp_path / 'samples'), 'BRAINWEBDIR': ''}
    options = get_options(batchsize=4, learningrate=1e-4, numEpochs=1, zDim=128, outputWidth=32, outputHeight=32, slices_start=20,
                          slices_end=32, config=cfgjson)
    options['data']['numPatients'] = 1
    options['data']['numTestPatients'] = 1
    hc, pc = get_datasets(options)
    config = get_config(fAnoGAN, options, 'ADAM', [8, 8], 0.2, hc)
    config.useTensorboard, config.verbose, config.device, config.math_mode = False, False, 'cpu', 0
    model = fAnoGAN(None, config, network=fanogan)
    x = hc.next_batch(4, set='TRAIN')[0]
    r = model.reconstruct(x)
    assert r['reconstruction'].shape == x.shape and 0.0 <= r['reconstruction'].min() and r['reconstruction'].max() <= 1.0
    np.random.seed(1)
    rd = model.reconstruct(x, dropout=True)                                    # both Dropout sites live, masks drawn by raw ABI calls
    assert rd['reconstruction'].shape == x.shape and not np.array_equal(rd['reconstruction'], r['reconstruction'])
    model.save(model.checkpointDir, 1)
    assert model.load_checkpoint() == 1
    ev = Evaluation.evaluate(pc, model, options, description='fanogan')
    assert ev['diffs'].shape == (12, 32, 32) and 0.0 <= ev['bestThreshold'] < 1.0 and len(ev['DiceScorePerPatient']) == 1


@pytest.mark.parametrize('tname,mname', [('AE', 'autoencoder_spatial'), ('ConstrainedAE', 'constrained_autoencoder'),
                                         ('ceVAE', 'context_encoder_variational_autoencoder'), ('VAE_You', 'variational_autoencoder')])
def test_remaining_trainer_surfaces_on_the_emulator(tname, mname, monkeypatch, tmp_path):
    """Every other GPU-verified trainer (spatial AE, constrained AE, ceVAE with its masked second input and anomaly map, VAE_You with
    the restoration loop as `reconstruct`) through train -> reconstruct on CPU: AEMODEL / the engine were edited after their last GPU run."""
    import importlib

    from unsupervised_anomaly_detection_brain_mri_b200.utils.default_config_setup import get_config, get_datasets, get_options
    _everything_on_the_emulator(monkeypatch)
    PKG = 'unsupervised_anomaly_detection_brain_mri_b200'
    trainer = getattr(importlib.import_module(f'{PKG}.trainers.{tname}'), tname)
    network = getattr(importlib.import_module(f'{PKG}.models.{mname}'), mname)
    cfgjson = {'CHECKPOINTDIR': str(tmp_path / 'ckpt'), 'SAMPLEDIR': str(tmp_path / 'samples'), 'BRAINWEBDIR': ''}
    options = get_options(batchsize=2, learningrate=1e-3, numEpochs=1, zDim=128, outputWidth=64, outputHeight=64, slices_start=40,
                          slices_end=48, config=cfgjson)
    options['data']['numPatients'] = 2
    hc, _ = get_datasets(options)
    config = get_config(trainer, options, 'ADAM', [8, 8], 0.2, hc)
    config.useTensorboard, config.verbose, config.device, config.math_mode, config.use_cuda_graph = False, False, 'cpu', 0, False
    if tname == 'VAE_You':
        config.restore_steps, config.tv_lambda = 2, 1.5
    import random
    random.seed(0)
    model = trainer(None, config, network=network)
    w0 = model.engine.fp.to_numpy()
    model.train(hc)
    w1 = model.engine.fp.to_numpy()
    assert all(np.isfinite(v).all() for v in w1.values()) and any(not np.array_equal(w0[k], w1[k]) for k in w0)
    assert model.engine.t == hc.num_batches(2, set='TRAIN')
    x = hc.next_batch(2, set='VAL')[0]
    r = model.reconstruct(x)
    assert r['reconstruction'].shape == x.shape and np.isfinite(r['reconstruction']).all()
    if tname == 'VAE_You':
        assert 0 < np.abs(r['reconstruction'] - x).max() < 0.1               # the restored INPUT, two small gradient steps away
    if tname == 'ceVAE':
        assert model.engine.anomaly.shape == (2, 64, 64, 1) and np.isfinite(model.engine.anomaly.numpy()).all()


# ------------------------------------------------------------------------------------------------ ceVAE.reconstruct (reference trainers/ceVAE.py:119-144)
@pytest.mark.parametrize('lam', [0.1, True, 0])
def test_cevae_reconstruct_applies_the_gradient_based_restoration(lam, monkeypatch, tmp_path):
    """``reconstruct`` returns ``x - lambda * anomaly`` with the PER-SLICE gradient of loss_vae (one slice per sess.run in the
    reference), the plain x_hat when the factor is falsy; the batched device stack must equal N single-slice oracle calls."""
    from unsupervised_anomaly_detection_brain_mri_b200.models.context_encoder_variational_autoencoder import \
        context_encoder_variational_autoencoder as net
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.ceVAE import ceVAE
    from unsupervised_anomaly_detection_brain_mri_b200.utils.default_config_setup import get_config, get_datasets, get_options
    _everything_on_the_emulator(monkeypatch)
    cfgjson = {'CHECKPOINTDIR': str(tmp_path / 'ckpt'), 'SAMPLEDIR': str(tmp_path / 'samples'), 'BRAINWEBDIR': '', 'SYNTHETICDIR': ''}
    S, N = 32, 3
    options = get_options(batchsize=2, learningrate=1e-3, numEpochs=1, zDim=128, outputWidth=S, outputHeight=S, config=cfgjson)
    options['data']['dir'] = ''
    hc, _ = get_datasets(options)
    config = get_config(ceVAE, options, 'ADAM', [8, 8], 0.1, hc)
    config.useTensorboard, config.verbose, config.device, config.math_mode, config.use_cuda_graph = False, False, 'cpu', 0, False
    assert config.use_gradient_based_restoration is True            # the reference's Config default (trainers/ceVAE.py:16)
    config.use_gradient_based_restoration = lam
    model = ceVAE(None, config, network=net)
    P = O.perturb_params(O.init_params(O.CEVAE, S, seed=1))
    model.engine.fp.load(P)
    eps = np.random.default_rng(5).standard_normal((N, 128)).astype(np.float32)
    orig_eval = model._eval_engine

    def eval_engine(n):
        e = E.adopt(orig_eval(n))
        monkeypatch.setattr(e, 'draw_noise', lambda dropout, rate: e.set_noise(eps[:n]))
        return e
    monkeypatch.setattr(model, '_eval_engine', eval_engine)
    x = O.synthetic_slices(N, S, seed=9)
    r = model.reconstruct(x)
    xh = model._eval_engines[N].br[0].xhat.numpy()
    ref = O.cevae_reconstruct(P, x, eps=eps, use_gradient_based_restoration=lam, dtype=torch.float64, l1_sign=np.sign(xh - x))
    assert _rel(xh, ref['x_hat']) < TOL
    assert _rel(r['anomaly'], ref['anomaly']) < 5e-5
    assert _rel(x - r['reconstruction'], x - ref['reconstruction']) < 5e-5
    if lam:
        assert np.allclose(x - r['reconstruction'], np.float32(lam) * r['anomaly'], rtol=0, atol=1e-7)
    else:
        assert np.array_equal(r['reconstruction'], xh.astype(np.float32))
    assert r['l1err'] == pytest.approx(np.sum(np.abs(x - r['reconstruction'])))
    one = model.reconstruct(x[1])               # [H,W,C] input as Evaluation passes it: same slice, same eps row 0 -> differs only by eps
    assert one['reconstruction'].shape == (1, S, S, 1)
