"""f-AnoGAN training path (trainers/fAnoGAN.py:50-77): the LayerNormalization backward / JVP / joint-backward kernels, the
WGAN-GP element-wise kernels, and the three train ops (optim_gen / optim_dis / optim_enc) of the CUDA engine against the
oracle's torch-autograd restatement (double backward for the gradient penalty).  Tolerance 1e-4 relative on forward
quantities and losses; gradients 1e-4 with the LeakyReLU / ReLU sub-gradient branches pinned to the implementation's sign pattern (FO._act)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import fanogan_cpu as FO  # noqa: E402
from oracle import tf_graph_cpu as O  # noqa: E402

from gpu_util import dptr  # noqa: E402

TOL = 1e-4
GTOL = 1e-4


def _rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()


def _ln_ref(z, gam, bet, act):
    """NHWC float64 torch LayerNormalization([1,2]) + activation."""
    mu = z.mean(dim=(1, 2), keepdim=True)
    var = z.var(dim=(1, 2), unbiased=False, keepdim=True)
    n = (z - mu) / torch.sqrt(var + 1e-3) * gam[None, :, :, None] + bet[None, :, :, None]
    return torch.nn.functional.leaky_relu(n, 0.3) if act == 1 else torch.relu(n)


@pytest.mark.parametrize('B,H,C,act', [(4, 8, 128, 2), (2, 16, 64, 1), (3, 32, 32, 1)])
def test_layernorm_train_kernels(B, H, C, act):
    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    L = abi.lib()
    rng = np.random.default_rng(5)
    HW = H * H
    z = rng.standard_normal((B, H, H, C)).astype(np.float32) * 1.5 + 0.3
    zd = rng.standard_normal((B, H, H, C)).astype(np.float32)
    dy = rng.standard_normal((B, H, H, C)).astype(np.float32)
    dyd = rng.standard_normal((B, H, H, C)).astype(np.float32)
    gam = (1 + 0.3 * rng.standard_normal((H, H))).astype(np.float32)
    bet = (0.2 * rng.standard_normal((H, H))).astype(np.float32)
    wsb = L.uad_layernorm_hw_train_workspace_bytes(B, HW, C)
    ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    dz_, dzd_, ddy, ddyd, dg, db = (_dev(v) for v in (z, zd, dy, dyd, gam, bet))
    y = torch.empty_like(dz_)
    stats = torch.empty(2 * B * C, device='cuda')
    abi.call('uad_layernorm_hw_fwd_train', dz_.data_ptr(), dg.data_ptr(), db.data_ptr(), y.data_ptr(), stats.data_ptr(), B, HW, C,
             1e-3, act, 0.3, ws.data_ptr(), wsb, st)
    # float64 reference with autograd
    tz = torch.from_numpy(z).double().requires_grad_(True)
    tzd = torch.from_numpy(zd).double().requires_grad_(True)
    tg = torch.from_numpy(gam).double().requires_grad_(True)
    tb = torch.from_numpy(bet).double().requires_grad_(True)
    ty = _ln_ref(tz, tg, tb, act)
    assert _rel(y.cpu().numpy(), ty.detach().numpy()) < TOL
    # backward
    gz, gg, gb = torch.autograd.grad((ty * torch.from_numpy(dy).double()).sum(), (tz, tg, tb), retain_graph=True)
    dx = torch.empty_like(dz_)
    dgam = torch.zeros(HW, device='cuda')
    dbet = torch.zeros(HW, device='cuda')
    abi.call('uad_layernorm_hw_bwd', ddy.data_ptr(), dz_.data_ptr(), stats.data_ptr(), dg.data_ptr(), db.data_ptr(), dx.data_ptr(),
             dgam.data_ptr(), dbet.data_ptr(), B, HW, C, act, 0.3, 0, ws.data_ptr(), wsb, st)
    assert _rel(dx.cpu().numpy(), gz.numpy()) < TOL
    assert _rel(dgam.cpu().numpy().reshape(H, H), gg.numpy()) < TOL
    assert _rel(dbet.cpu().numpy().reshape(H, H), gb.numpy()) < TOL
    # JVP
    _, tyd = torch.autograd.functional.jvp(lambda a: _ln_ref(a, tg, tb, act), (tz,), (tzd,), create_graph=True)
    yd = torch.empty_like(dz_)
    js = torch.empty(2 * B * C, device='cuda')
    abi.call('uad_layernorm_hw_jvp', dzd_.data_ptr(), dz_.data_ptr(), stats.data_ptr(), dg.data_ptr(), db.data_ptr(), yd.data_ptr(),
             js.data_ptr(), B, HW, C, act, 0.3, ws.data_ptr(), wsb, st)
    assert _rel(yd.cpu().numpy(), tyd.detach().numpy()) < TOL
    # joint reverse of (y, ydot)
    for with_dy in (True, False):
        Lsum = (tyd * torch.from_numpy(dyd).double()).sum()
        if with_dy:
            Lsum = Lsum + (ty * torch.from_numpy(dy).double()).sum()
        gz2, gzd2, gg2, gb2 = torch.autograd.grad(Lsum, (tz, tzd, tg, tb), retain_graph=True, allow_unused=True)
        dxd = torch.empty_like(dz_)
        dx2 = torch.empty_like(dz_)
        dgam.zero_()
        dbet.zero_()
        abi.call('uad_layernorm_hw_bwd2', ddyd.data_ptr(), ddy.data_ptr() if with_dy else None, dz_.data_ptr(), dzd_.data_ptr(),
                 stats.data_ptr(), js.data_ptr(), dg.data_ptr(), db.data_ptr(), dxd.data_ptr(), dx2.data_ptr(), dgam.data_ptr(),
                 dbet.data_ptr(), B, HW, C, act, 0.3, 1, ws.data_ptr(), wsb, st)
        assert _rel(dxd.cpu().numpy(), gzd2.numpy()) < TOL
        assert _rel(dx2.cpu().numpy(), gz2.numpy()) < 2 * TOL
        assert _rel(dgam.cpu().numpy().reshape(H, H), gg2.numpy()) < TOL
        if with_dy:
            assert _rel(dbet.cpu().numpy().reshape(H, H), gb2.numpy()) < TOL
        else:
            assert float(dbet.abs().max()) == 0.0


def test_wgan_elementwise_kernels():
    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    L = abi.lib()
    rng = np.random.default_rng(9)
    st = torch.cuda.current_stream().cuda_stream
    B, H, W = 3, 64, 64
    ws = torch.empty(max(L.uad_reduce_workspace_bytes(), B * W * 4, 1 << 20), dtype=torch.uint8, device='cuda')
    wsb = ws.numel()
    ddx = (rng.standard_normal((B, H, W, 1)) * 0.2).astype(np.float32)
    d_ddx = _dev(ddx)
    u = torch.empty_like(d_ddx)
    out = torch.zeros(4, device='cuda')
    abi.call('uad_gradient_penalty', d_ddx.data_ptr(), B, H, W, 10.0, u.data_ptr(), out.data_ptr(), ws.data_ptr(), wsb, st)
    t = torch.from_numpy(ddx).double().requires_grad_(True)
    gp = ((torch.sqrt((t * t).sum(dim=1)) - 1.0) ** 2).mean() * 10.0
    gu, = torch.autograd.grad(gp, t)
    assert abs(float(out[0]) - float(gp.detach())) / float(gp.detach()) < 1e-6
    assert _rel(u.cpu().numpy(), gu.numpy()) < 1e-5
    # sums / mse
    a = rng.standard_normal(100003).astype(np.float32)
    b = rng.standard_normal(100003).astype(np.float32)
    da, db = _dev(a), _dev(b)
    g = torch.empty_like(da)
    abi.call('uad_sum_scaled', da.data_ptr(), a.size, 1.0 / a.size, out.data_ptr(), ws.data_ptr(), wsb, st)
    abi.call('uad_mse', da.data_ptr(), db.data_ptr(), a.size, 0.5, g.data_ptr(), 1.0 / a.size, out[1:].data_ptr(), ws.data_ptr(),
             wsb, st)
    assert abs(float(out[0]) - a.astype(np.float64).mean()) < 1e-7
    assert abs(float(out[1]) - ((a.astype(np.float64) - b) ** 2).mean()) < 1e-6
    assert np.array_equal(g.cpu().numpy(), np.float32(0.5) * (a - b))
    # interpolate, l1 map, activation backward, fill, uniform
    x = rng.random((B, H, W, 1), dtype=np.float32)
    xg = rng.random((B, H, W, 1), dtype=np.float32)
    al = rng.random(B, dtype=np.float32)
    o = torch.empty(B, H, W, 1, device='cuda')
    abi.call('uad_interpolate', dptr(x), dptr(xg), dptr(al), o.data_ptr(), B, H * W, st)
    assert np.allclose(o.cpu().numpy(), x + al[:, None, None, None] * (xg - x), rtol=0, atol=1e-7)
    l1 = torch.empty(B, H, W, 1, device='cuda')
    rec = torch.empty(B, device='cuda')
    abi.call('uad_l1_map', dptr(x), dptr(xg), l1.data_ptr(), rec.data_ptr(), B, H * W, st)
    assert np.array_equal(l1.cpu().numpy(), np.abs(xg - x))
    assert _rel(rec.cpu().numpy(), np.abs(xg.astype(np.float64) - x).sum(axis=(1, 2, 3))) < 1e-6
    for act, f in ((3, lambda t: torch.sigmoid(t)), (4, lambda t: torch.tanh(t))):
        tu = torch.from_numpy(a).double().requires_grad_(True)
        gr, = torch.autograd.grad((f(tu) * torch.from_numpy(b).double()).sum(), tu)
        abi.call('uad_activation_bwd', db.data_ptr(), da.data_ptr(), g.data_ptr(), a.size, act, 0.0, st)
        assert _rel(g.cpu().numpy(), gr.numpy()) < 1e-5
    abi.call('uad_fill', g.data_ptr(), 2.5, a.size, st)
    assert float(g.min()) == 2.5 == float(g.max())
    un = torch.empty(1 << 16, device='cuda')
    abi.call('uad_uniform', un.data_ptr(), un.numel(), 7, 0, None, st)
    h = un.cpu().numpy()
    assert 0.0 <= h.min() and h.max() < 1.0 and abs(h.mean() - 0.5) < 0.01 and abs(h.var() - 1 / 12) < 0.005
    # final 1x1 backward for an arbitrary incoming gradient
    Cin, npx = 32, B * H * W
    act_in = rng.standard_normal((npx, Cin)).astype(np.float32)
    w = rng.standard_normal(Cin).astype(np.float32)
    dxh = rng.standard_normal(npx).astype(np.float32)
    dact = torch.empty(npx, Cin, device='cuda')
    dw = torch.empty(Cin, device='cuda')
    dbias = torch.empty(1, device='cuda')
    abi.call('uad_final1x1_bwd', dptr(act_in), dptr(w), dptr(dxh), dact.data_ptr(), dw.data_ptr(), dbias.data_ptr(),
             B, H * W, Cin, 0, ws.data_ptr(), wsb, st)
    assert np.allclose(dact.cpu().numpy(), dxh[:, None] * w[None, :], rtol=1e-6, atol=1e-7)
    assert _rel(dw.cpu().numpy(), act_in.astype(np.float64).T @ dxh) < 1e-5
    assert abs(float(dbias) - dxh.astype(np.float64).sum()) < 1e-3


def _feed(S, B, rate, flat, seed=21):
    rng = np.random.default_rng(seed)
    x = O.synthetic_slices(B, S, seed=seed)
    z = rng.standard_normal((B, 128)).astype(np.float32)
    alpha = rng.random((B, 1), dtype=np.float32)
    m_enc = (rng.uniform(size=(B, 128)) >= rate).astype(np.float32)
    m_gen = (rng.uniform(size=(B, flat)) >= rate).astype(np.float32)
    return x, z, alpha, m_enc, m_gen


def _signs(eng, which, rate):
    """Activation sign patterns of the implementation's own forward passes (deterministic, so identical to the ones the train
    op computes): the oracle differentiates with the same sub-gradient branch at LeakyReLU / ReLU kinks (see FO._act)."""
    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    keep = 1.0 / (1.0 - rate)

    def pat(ts):
        return [(t > 0).cpu().numpy() for t in ts]

    def critic(x_dev):
        eng._critic_forward(eng.pass1, x_dev, critic=False)
        return pat(eng.pass1.a)

    sg = {}
    if which in ('gen', 'disc'):
        eng.generate(eng.z_in, eng.mask_gen, keep, out=eng.x_gen)
        sg['gen_z'] = pat([eng.ar] + eng.gen_a)
        sg['d_fake'] = critic(eng.x_gen)
    if which == 'disc':
        sg['d_real'] = critic(eng.x)
        abi.call('uad_interpolate', eng.x.data_ptr(), eng.x_gen.data_ptr(), eng.alpha.data_ptr(), eng.x_hat.data_ptr(), eng.B,
                 eng.S * eng.S, torch.cuda.current_stream().cuda_stream)
        sg['d_hat'] = critic(eng.x_hat)
    if which == 'enc':
        z_enc = eng.encode(eng.mask_enc, keep)
        sg['enc'] = pat(eng.enc_a)
        x_enc = eng.generate(z_enc, eng.mask_gen, keep, out=eng.x_enc)
        sg['gen_enc'] = pat([eng.ar] + eng.gen_a)
        sg['d_enc'] = critic(x_enc)
        sg['d_real'] = critic(eng.x)
    torch.cuda.synchronize()
    return sg


def _compare_grads(eng, G, scope, skip_ln_bias=True):
    got = eng.fp.to_numpy(eng.fp.grads)
    gmax = max(float(v.abs().max()) for v in G.values())
    worst = 0.0
    for k, v in G.items():
        ref = v.numpy()
        if skip_ln_bias and k.endswith('/bias') and float(np.abs(ref).max()) < 1e-5 * gmax:
            # biases feeding a LayerNorm over (H,W): exact gradient 0 (autograd returns round-off); the engine writes 0
            assert float(np.abs(got[k]).max()) <= 1e-5 * gmax, k
            continue
        err = _rel(got[k], ref)
        worst = max(worst, err)
        assert err < GTOL, (k, err)
    return worst


@pytest.mark.parametrize('mode', [0, 1])
@pytest.mark.parametrize('S,B', [(32, 4), (64, 2)])
@pytest.mark.parametrize('which', ['gen', 'disc', 'enc'])
def test_fanogan_train_ops_match_oracle(which, S, B, mode):
    from unsupervised_anomaly_detection_brain_mri_b200.fanogan_engine import FanoganEngine
    rate, lr = 0.2, 1e-3
    P = FO.perturb(FO.init_params(S, seed=1))
    eng = FanoganEngine(S, batch=B, math_mode=mode, kappa=1.0, scale=10.0)
    x, z, alpha, m_enc, m_gen = _feed(S, B, rate, eng.flat)
    eng.fp.load(P)
    eng.enable_training()
    eng.set_inputs(x)
    eng.set_latent(z)
    eng.alpha.copy_(torch.from_numpy(alpha.reshape(-1)))
    eng.mask_enc.copy_(torch.from_numpy(m_enc))
    eng.mask_gen.copy_(torch.from_numpy(m_gen))
    tr = FO.WganTrainer(P, lr=lr, dropout_rate=rate, scale=10.0, kappa=1.0, dtype=torch.float64)
    out, G = tr.step(which, x, z, alpha, mask_enc=m_enc, mask_gen=m_gen, signs=_signs(eng, which, rate))
    step = {'gen': eng.step_gen, 'disc': eng.step_disc, 'enc': eng.step_enc}[which]
    scope = {'gen': 'Generator', 'disc': 'Discriminator', 'enc': 'Encoder'}[which]
    before = eng.fp.to_numpy()
    res = step(lr, dropout_rate=rate, dropout=True, parity_noise=True)
    torch.cuda.synchronize()
    for k, v in res.items():
        if k in out:
            assert abs(v - float(out[k])) <= TOL * max(abs(float(out[k])), 1e-3), (k, v, float(out[k]))
    if which == 'disc':
        assert _rel(eng.ddx.cpu().numpy(), out['ddx'].numpy()) < GTOL
        assert _rel(eng.x_hat.cpu().numpy(), out['x_hat'].detach().numpy()) < TOL
    if which == 'enc':
        assert _rel(eng.x_enc.cpu().numpy(), out['x_enc'].numpy()) < TOL
        assert _rel(eng.l1.cpu().numpy(), out['L1'].numpy()) < 2e-4
    _compare_grads(eng, G, scope)
    # the Adam update touched exactly the scope's variables and matches tf.train.AdamOptimizer(lr, 0.5, 0.9)
    after = eng.fp.to_numpy()
    gmax = max(float(v.abs().max()) for v in G.values())
    bad = tot = 0
    for k in after:
        if not k.startswith(scope + '/'):
            assert np.array_equal(after[k], before[k]), k
            continue
        ref = tr.P[k].numpy()
        g = G[k].numpy()
        sel = np.abs(g) > 1e-3 * gmax            # the first Adam step is ~lr*sign(g): compare where the sign is determined
        tot += int(sel.sum())
        bad += int((np.abs(after[k].reshape(ref.shape) - ref)[sel] > 0.02 * lr).sum())
    assert tot > 0 and bad <= 5e-3 * tot, (bad, tot)


def test_fanogan_trainer_train_loop(tmp_path):
    """trainers/fAnoGAN.train: WGAN phase (1 G-step + 5 D-steps per batch) then the encoder phase, on the synthetic dataset."""
    from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import SYNTHETIC
    from unsupervised_anomaly_detection_brain_mri_b200.models.fanogan import fanogan
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.fAnoGAN import fAnoGAN
    config = fAnoGAN.Config()
    config.outputHeight = config.outputWidth = 32
    config.batchsize = 4
    config.numEpochs = 1
    config.zDim = 128
    config.numChannels = 1
    config.intermediateResolutions = [8, 8]
    config.dropout_rate = 0.1
    config.learningrate = 1e-4
    config.checkpointDir = str(tmp_path / 'ckpt')
    config.description = 'gpu-test'
    config.dataset = 'SYNTHETIC'
    opts = SYNTHETIC.Options()
    opts.sliceResolution = (32, 32)
    opts.numPatients = 1
    opts.sliceStart, opts.sliceEnd = 20, 60
    ds = SYNTHETIC(opts)
    np.random.seed(0)
    model = fAnoGAN(None, config, network=fanogan)
    w0 = model.engine.fp.to_numpy()
    model.train(ds)
    w1 = model.engine.fp.to_numpy()
    for scope in ('Encoder', 'Generator', 'Discriminator'):
        assert any(not np.array_equal(w0[k], w1[k]) for k in w0 if k.startswith(scope + '/')), scope
    assert all(np.isfinite(v).all() for v in w1.values())
    assert model.engine.t == {'Generator': model.engine.t['Generator'], 'Discriminator': 5 * model.engine.t['Generator'],
                              'Encoder': model.engine.t['Encoder']}
    assert model.engine.t['Generator'] > 0 and model.engine.t['Encoder'] > 0
    rec = model.reconstruct(ds.next_batch(4, set='VAL')[0][0])
    assert rec['reconstruction'].shape == (1, 32, 32, 1) and np.isfinite(rec['l1err'])


@pytest.mark.parametrize('which', ['gen', 'disc', 'enc'])
def test_fanogan_graph_replay_equals_eager(which):
    """A train op replayed from its CUDA graph (eager warm-up, capture, replays) leaves bit-identical weights, Adam moments
    and loss scalars to the same op issued eagerly (all reductions are deterministic)."""
    from unsupervised_anomaly_detection_brain_mri_b200.fanogan_engine import FanoganEngine
    S, B, rate, lr, steps = 32, 4, 0.2, 1e-3, 5
    P = FO.perturb(FO.init_params(S, seed=1))
    runs = []
    for use_graph in (False, True):
        eng = FanoganEngine(S, batch=B, math_mode=1)
        x, z, alpha, m_enc, m_gen = _feed(S, B, rate, eng.flat)
        eng.fp.load(P)
        eng.enable_training()
        eng.set_inputs(x)
        eng.set_latent(z)
        eng.alpha.copy_(torch.from_numpy(alpha.reshape(-1)))
        eng.mask_enc.copy_(torch.from_numpy(m_enc))
        eng.mask_gen.copy_(torch.from_numpy(m_gen))
        step = {'gen': eng.step_gen, 'disc': eng.step_disc, 'enc': eng.step_enc}[which]
        res = [step(lr, dropout_rate=rate, dropout=True, parity_noise=True, use_graph=use_graph) for _ in range(steps)]
        torch.cuda.synchronize()
        assert (len(eng._graphs) == 1) == use_graph
        runs.append((eng.fp.to_numpy(), eng.fp.to_numpy(eng.fp.m), eng.fp.to_numpy(eng.fp.v), res, dict(eng.t)))
    (w0, m0, v0, r0, t0), (w1, m1, v1, r1, t1) = runs
    assert t0 == t1
    assert r0 == r1
    for k in w0:
        assert np.array_equal(w0[k], w1[k]), k
        assert np.array_equal(m0[k], m1[k]) and np.array_equal(v0[k], v1[k]), k


def test_fanogan_graph_noise_advances():
    """Perf-mode noise under graph replay: the Philox offset lives on the device, so every replay draws fresh dropout
    masks / interpolation coefficients."""
    from unsupervised_anomaly_detection_brain_mri_b200.fanogan_engine import FanoganEngine
    S, B = 32, 4
    eng = FanoganEngine(S, batch=B, math_mode=1)
    eng.enable_training()
    eng.set_inputs(O.synthetic_slices(B, S, seed=5))
    eng.set_latent(np.random.default_rng(0).standard_normal((B, 128)).astype(np.float32))
    seen = []
    for _ in range(4):
        eng.step_disc(1e-4, dropout_rate=0.2, dropout=True, use_graph=True)
        seen.append((eng.alpha.cpu().numpy().copy(), eng.mask_gen.cpu().numpy().copy()))
    assert len(eng._graphs) == 1
    for i in range(1, 4):
        assert not np.array_equal(seen[i][0], seen[i - 1][0])
        assert not np.array_equal(seen[i][1], seen[i - 1][1])
    keep = np.mean([m.mean() for _, m in seen])
    assert 0.75 < keep < 0.85
