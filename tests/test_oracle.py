"""CPU: the oracle restatement against float64 naive loops and against the committed golden fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import naive64, scoring
from oracle import tf_graph_cpu as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def t64(a):
    return torch.from_numpy(np.ascontiguousarray(a)).double()


@pytest.mark.parametrize('H,Ci,Co', [(8, 3, 4), (16, 1, 5), (4, 2, 2)])
def test_conv_same_matches_naive_loop(H, Ci, Co):
    rng = np.random.default_rng(H + Ci)
    x = rng.standard_normal((2, H, H, Ci)).astype(np.float32)
    w = rng.standard_normal((5, 5, Ci, Co)).astype(np.float32)
    b = rng.standard_normal(Co).astype(np.float32)
    y = O.conv2d_same_s2(t64(x).permute(0, 3, 1, 2), t64(w), t64(b)).permute(0, 2, 3, 1).numpy()
    assert np.abs(y - naive64.conv2d_same_s2(x, w, b)).max() < 1e-12
    K = rng.standard_normal((5, 5, Co, Ci)).astype(np.float32)
    yT = O.conv2dT_same_s2(t64(x).permute(0, 3, 1, 2), t64(K), t64(b)).permute(0, 2, 3, 1).numpy()
    assert yT.shape == (2, 2 * H, 2 * H, Co)
    assert np.abs(yT - naive64.conv2dT_same_s2(x, K, b)).max() < 1e-12


def test_convT_is_adjoint_of_conv():
    """Conv2DTranspose(SAME, s2) is the input-gradient of Conv2D(SAME, s2): <conv(x), y> == <x, convT(y)>."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, 8, 8, 3))
    y = rng.standard_normal((1, 4, 4, 2))
    w = rng.standard_normal((5, 5, 3, 2))
    lhs = (naive64.conv2d_same_s2(x, w, np.zeros(2)) * y).sum()
    rhs = (x * naive64.conv2dT_same_s2(y, w, np.zeros(3))).sum()      # w as [kh,kw,Cout=3,Cin=2]
    assert abs(lhs - rhs) < 1e-9 * max(1.0, abs(lhs))


def test_bn_kl_adam_micro_oracles():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 4, 4, 3)).astype(np.float32)
    g, b = rng.standard_normal(3).astype(np.float32), rng.standard_normal(3).astype(np.float32)
    y = O.bn_frozen(t64(x).permute(0, 3, 1, 2), t64(g), t64(b)).permute(0, 2, 3, 1).numpy()
    assert np.abs(y - naive64.bn_frozen(x, g, b)).max() < 1e-12
    mu, ls = rng.standard_normal((3, 7)).astype(np.float32), (0.3 * rng.standard_normal((3, 7))).astype(np.float32)
    out = {'x_hat': torch.zeros(3, 2, 2, 1, dtype=torch.float64), 'z_mu': t64(mu), 'z_sigma': torch.exp(t64(ls))}
    L = O.losses(O.VAE, out, np.zeros((3, 2, 2, 1), np.float32), dtype=torch.float64)
    assert abs(float(L['kl']) - naive64.kl_per_sample(mu, ls).mean()) < 1e-12
    p, gr = rng.standard_normal(9), rng.standard_normal(9)
    Pn, mn, vn = O.adam_tf({'a': t64(p)}, {'a': t64(gr)}, {'a': t64(np.zeros(9))}, {'a': t64(np.zeros(9))}, 1, 1e-3, 0.5)
    pr, mr, vr = naive64.adam_tf(p, gr, np.zeros(9), np.zeros(9), 1, 1e-3)
    assert np.abs(Pn['a'].numpy() - pr).max() < 1e-15
    # first TF-Adam step moves every weight by ~lr*sign(g) (epsilon outside the bias correction)
    assert np.allclose(np.abs(p - pr), 1e-3, rtol=1e-3)


def test_param_counts_match_survey():
    assert sum(v.size for v in O.init_params(O.VAE, 256).values()) == 2192337 + 1792      # SURVEY 8d + BN gamma/beta
    n, enc, dec = O.stack_plan(256)
    assert (n, enc, dec) == (5, [32, 64, 128, 128, 128], [128, 64, 32, 32, 32])
    assert O.stack_plan(128)[0] == 4


def test_ae_dropout_flag_quirk():
    """autoencoder.py:30: the dropout on dec_dense(z) has no training flag -> only the 'z' mask has an effect."""
    P = O.init_params(O.AE, 32)
    x = O.synthetic_slices(2, 32)
    m = {'z': (np.random.default_rng(0).uniform(size=(2, 128)) >= 0.5).astype(np.float32)}
    a = O.forward(O.AE, P, x, masks=m, dropout_rate=0.5, training=True)['x_hat']
    b = O.forward(O.AE, P, x, masks=m, dropout_rate=0.5, training=False)['x_hat']
    assert not torch.allclose(a, b)


@pytest.mark.parametrize('name,arch', [('ae_32_b2', O.AE), ('vae_32_b2', O.VAE), ('cevae_32_b2', O.CEVAE)])
def test_oracle_reproduces_golden(name, arch):
    g = np.load(os.path.join(GOLD, name + '.npz'))
    P = O.perturb_params(O.init_params(arch, 32, seed=1))
    names = list(g['param_names'])
    assert names == list(P.keys())
    assert np.allclose([float(P[k].astype(np.float64).sum()) for k in names], g['param_sum'], rtol=0, atol=1e-9)
    masks = {k[5:]: g[k] for k in g.files if k.startswith('mask_')}
    out, L, G = O.loss_and_grads(arch, P, g['x'], x_ce=g['x_ce'], eps=g['eps'], masks=masks, dropout_rate=float(g['rate']),
                                 training=True, dtype=torch.float32, want_anomaly=(arch == O.CEVAE))
    assert np.abs(out['x_hat'].numpy() - g['x_hat']).max() < 1e-5
    assert abs(float(L['loss']) - float(g['loss_loss'])) < 1e-4 * abs(float(g['loss_loss']))
    gl2 = np.array([float(G[k].double().norm()) for k in names])
    assert np.allclose(gl2, g['grad_l2'], rtol=1e-4)


def test_scoring_oracle_reproduces_golden():
    g = np.load(os.path.join(GOLD, 'scoring.npz'))
    sub = scoring.residual(g['x'], g['x_rec'], g['mask'].astype(bool), float(g['prior']), True, False)
    assert np.array_equal(sub.astype(np.float32), g['diff'])
    best, thr, ths, scs = scoring.best_dice_search(sub, g['labels'], granularity=4)
    assert best == float(g['best_dice']) and thr == float(g['best_thr'])
    assert np.array_equal(np.array(ths), g['threshs'])


def test_threshold_compare_is_float64():
    d = np.array([np.float32(0.3)], np.float64)
    assert scoring.threshold_mask(d, 0.3)[0]            # float32(0.3) = 0.30000001192... > 0.3
    assert not scoring.threshold_mask(d, 0.30000002)[0]


def test_restore_gradient_finite_difference():
    """oracle.restore_gradient (trainers/VAE_You.py:47-54): tf.image.total_variation restated + d/dx checked by central
    differences in float64 at pixels away from the |.| kinks."""
    import torch
    from oracle import tf_graph_cpu as O
    S, B, lam = 16, 2, 1.3
    P = O.perturb_params(O.init_params(O.VAE, S, seed=3))
    x = O.synthetic_slices(B, S, seed=2).astype(np.float64)
    eps = np.random.default_rng(1).standard_normal((B, 128))
    d = np.random.default_rng(0).random((2, 5, 7, 1))
    tv = O.total_variation(torch.from_numpy(d)).numpy()
    ref = np.abs(np.diff(d, axis=1)).sum((1, 2, 3)) + np.abs(np.diff(d, axis=2)).sum((1, 2, 3))
    assert np.allclose(tv, ref)

    def objective(xx):
        out = O.forward(O.VAE, P, xx, eps=eps, dtype=torch.float64)
        xt = torch.from_numpy(xx)
        rec = (out['x_hat'] - xt).abs().sum(dim=(1, 2, 3))
        kl = 0.5 * (out['z_mu'] ** 2 + out['z_sigma'] ** 2 - torch.log(out['z_sigma'] ** 2) - 1).sum(dim=1)
        return float((rec + kl + lam * O.total_variation(xt - out['x_hat'])).sum())

    g, out, _ = O.restore_gradient(O.VAE, P, x, eps=eps, tv_lambda=lam, dtype=torch.float64)
    g = g.numpy()
    h = 1e-6
    rng = np.random.default_rng(5)
    checked = 0
    for _ in range(12):
        b, i, j = int(rng.integers(B)), int(rng.integers(1, S - 1)), int(rng.integers(1, S - 1))
        xp, xm = x.copy(), x.copy()
        xp[b, i, j, 0] += h
        xm[b, i, j, 0] -= h
        fd = (objective(xp) - objective(xm)) / (2 * h)
        if abs(fd - g[b, i, j, 0]) < 1e-4 * max(1.0, abs(fd)):
            checked += 1
    assert checked >= 10          # the remaining probes may straddle a kink of |.|


def test_anovaegan_oracle_gradients_by_central_differences():
    """oracle/anovaegan_cpu: autograd gradients of the three op losses (incl. the second-order gradient-penalty term) against
    central differences in float64, on one entry per scope."""
    from oracle import anovaegan_cpu as AO
    from oracle.fanogan_cpu import perturb
    S, B, Z = 32, 2, 16
    P = perturb(AO.init_params(S, zDim=Z, seed=3))
    rng = np.random.default_rng(0)
    x = rng.uniform(size=(B, S, S, 1)).astype(np.float32)
    eps = rng.standard_normal((B, Z)).astype(np.float32)
    alpha = rng.uniform(size=(B, 1)).astype(np.float32)

    def loss(Pv, which):
        o = AO.graph(AO.as_leaves(Pv, torch.float64), x, eps, alpha, None, 0.0, True, 10.0, 0.5, torch.float64, want=(which,))
        return o[{'vae': 'enc_loss', 'gen': 'gen_loss', 'disc': 'disc_loss'}[which]]

    probes = {'vae': ['Encoder/dense_1/kernel', 'Generator/dec_Conv2DT_0/kernel', 'Encoder/batch_normalization/gamma'],
              'gen': ['Generator/dense_2/kernel', 'Generator/layer_normalization_1/gamma'],
              'disc': ['Discriminator/enc_conv2D_1/kernel', 'Discriminator/dense_3/kernel']}
    for which, names in probes.items():
        L = AO.as_leaves(P, torch.float64)
        o = AO.graph(L, x, eps, alpha, None, 0.0, True, 10.0, 0.5, torch.float64, want=(which,))
        g = torch.autograd.grad(o[{'vae': 'enc_loss', 'gen': 'gen_loss', 'disc': 'disc_loss'}[which]], [L[n] for n in names])
        for n, gn in zip(names, g):
            idx = np.unravel_index(int(np.argmax(np.abs(gn.numpy()))), gn.shape)
            h = 1e-5
            Pp, Pm = dict(P), dict(P)
            Pp[n] = P[n].astype(np.float64).copy()
            Pm[n] = P[n].astype(np.float64).copy()
            Pp[n][idx] += h
            Pm[n][idx] -= h
            fd = (float(loss(Pp, which).detach()) - float(loss(Pm, which).detach())) / (2 * h)
            assert abs(fd - float(gn[idx])) <= 1e-5 * max(abs(fd), 1e-6), (which, n, fd, float(gn[idx]))


def test_sibling_oracles_reproduce_golden():
    """AnoVAEGAN / AAE / constrained AAE / GMVAE / spatial GMVAE oracles against the committed scalars and gradient norms
    (tests/golden/siblings_32_b2.json, written by oracle/make_golden.py from the same seeded feeds)."""
    import json
    from oracle.make_golden import sibling_outputs
    with open(os.path.join(GOLD, 'siblings_32_b2.json')) as fh:
        gold = json.load(fh)
    now = sibling_outputs()
    assert sorted(now) == sorted(gold)
    for case, g in gold.items():
        for k, v in g['scalars'].items():
            assert abs(now[case]['scalars'][k] - v) <= 2e-4 * max(1e-3, abs(v)), (case, k)
        assert sorted(now[case]['grad_l2']) == sorted(g['grad_l2'])
        for k, v in g['grad_l2'].items():
            assert abs(now[case]['grad_l2'][k] - v) <= 2e-4 * max(1e-4, abs(v)) + 1e-7, (case, k)
