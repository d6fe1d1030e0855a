"""fp32/fp64 torch-CPU restatement of the reference's TF-1.15 graphs (oracle; PARITY UNPINNED).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.

Each function cites the reference file:line it restates (paths relative to
/root/reference).  TF-1.15 op semantics that the reference never states
(SAME padding split, kernel layouts, frozen BN, LeakyReLU alpha, flatten order,
dropout scaling, Adam form, initialisers) follow SURVEY.md Appendix A and are
cross-checked against ``oracle/naive64.py`` in ``tests/test_oracle.py``.

All tensors at the interface are NHWC numpy/torch arrays, weights are in TF
layouts (HWIO conv kernels, [kh,kw,Cout,Cin] transposed-conv kernels, [in,out]
dense kernels) keyed by TF variable names.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3          # tf.layers.BatchNormalization default epsilon
LRELU_ALPHA = 0.3      # tf.keras.layers.LeakyReLU() default alpha (models/customlayers.py:23,36)

AE = 'autoencoder'
VAE = 'variational_autoencoder'
CEVAE = 'context_encoder_variational_autoencoder'
AES = 'autoencoder_spatial'
CAE = 'constrained_autoencoder'
ARCHS = (AE, VAE, CEVAE, AES, CAE)


# --------------------------------------------------------------------------- layer plan
def stack_plan(S: int, res: int = 8):
    """Channel plan of build_unified_encoder / build_unified_decoder (models/customlayers.py:16-38)."""
    n = int(math.log(S, 2) - math.log(float(res), 2))
    enc = [int(min(128, 32 * (2 ** i))) for i in range(n)]
    dec = [int(max(32, 128 / (2 ** i))) for i in range(n)]
    return n, enc, dec


def _glorot(rng, shape, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def init_params(arch: str, S: int, C: int = 1, zDim: int = 128, res: int = 8, seed: int = 1):
    """Glorot-uniform kernels, zero biases, gamma=1, beta=0, TF variable names (A.10)."""
    assert arch in ARCHS
    rng = np.random.default_rng(seed)
    n, enc, dec = stack_plan(S, res)
    P = OrderedDict()
    cin = C
    bn = 0

    def bn_name(k):
        return 'batch_normalization' if k == 0 else f'batch_normalization_{k}'

    for i, co in enumerate(enc):
        P[f'Encoder/enc_conv2D_{i}/kernel'] = _glorot(rng, (5, 5, cin, co), 25 * cin, 25 * co)
        P[f'Encoder/enc_conv2D_{i}/bias'] = np.zeros(co, np.float32)
        P[f'Encoder/{bn_name(bn)}/gamma'] = np.ones(co, np.float32)
        P[f'Encoder/{bn_name(bn)}/beta'] = np.zeros(co, np.float32)
        bn += 1
        cin = co
    cb = cin // 8
    if arch != AES:                      # autoencoder_spatial.py has no Bottleneck scope at all
        P['Bottleneck/conv2d/kernel'] = _glorot(rng, (1, 1, cin, cb), cin, cb)
        P['Bottleneck/conv2d/bias'] = np.zeros(cb, np.float32)
        P['Bottleneck/conv2d_1/kernel'] = _glorot(rng, (1, 1, cb, cin), cb, cin)
        P['Bottleneck/conv2d_1/bias'] = np.zeros(cin, np.float32)
        flat = res * res * cb
        heads = 1 if arch in (AE, CAE) else 2
        for h in range(heads):
            nm = 'dense' if h == 0 else f'dense_{h}'
            P[f'Bottleneck/{nm}/kernel'] = _glorot(rng, (flat, zDim), flat, zDim)
            P[f'Bottleneck/{nm}/bias'] = np.zeros(zDim, np.float32)
        nm = f'dense_{heads}'
        P[f'Bottleneck/{nm}/kernel'] = _glorot(rng, (zDim, flat), zDim, flat)
        P[f'Bottleneck/{nm}/bias'] = np.zeros(flat, np.float32)
    P[f'Decoder/{bn_name(bn)}/gamma'] = np.ones(cin, np.float32)
    P[f'Decoder/{bn_name(bn)}/beta'] = np.zeros(cin, np.float32)
    bn += 1
    for i, co in enumerate(dec):
        P[f'Decoder/dec_Conv2DT_{i}/kernel'] = _glorot(rng, (5, 5, co, cin), 25 * co, 25 * cin)
        P[f'Decoder/dec_Conv2DT_{i}/bias'] = np.zeros(co, np.float32)
        P[f'Decoder/{bn_name(bn)}/gamma'] = np.ones(co, np.float32)
        P[f'Decoder/{bn_name(bn)}/beta'] = np.zeros(co, np.float32)
        bn += 1
        cin = co
    P['Decoder/dec_Conv2D_final/kernel'] = _glorot(rng, (1, 1, cin, C), cin, C)
    P['Decoder/dec_Conv2D_final/bias'] = np.zeros(C, np.float32)
    return P


def perturb_params(P, seed: int = 7, scale: float = 0.05):
    """Make biases / gamma / beta non-trivial so parity tests exercise every term."""
    rng = np.random.default_rng(seed)
    Q = OrderedDict()
    for k, v in P.items():
        if k.endswith('/kernel'):
            Q[k] = v.copy()
        elif k.endswith('/gamma'):
            Q[k] = (v + scale * rng.standard_normal(v.shape)).astype(np.float32)
        else:
            Q[k] = (v + scale * rng.standard_normal(v.shape)).astype(np.float32)
    return Q


# --------------------------------------------------------------------------- TF ops (NCHW inside)
def conv2d_same_s2(x, w_hwio, b):
    """tf Conv2D(k, strides=2, padding='same') (models/customlayers.py:21).  SAME on even input with k=5:
    total pad 3 -> 1 before, 2 after."""
    k = w_hwio.shape[0]
    H = x.shape[2]
    out = -(-H // 2)
    pad_total = max((out - 1) * 2 + k - H, 0)
    lo = pad_total // 2
    hi = pad_total - lo
    return F.conv2d(F.pad(x, (lo, hi, lo, hi)), w_hwio.permute(3, 2, 0, 1), b, stride=2)


def conv2dT_same_s2(x, k_hwoi, b):
    """tf Conv2DTranspose(k=5, strides=2, padding='same') (models/customlayers.py:34): out[2i+k-1] += K[k]*x[i]
    cropped to 2n.  Kernel layout [kh,kw,Cout,Cin]."""
    n = x.shape[2]
    y = F.conv_transpose2d(x, k_hwoi.permute(3, 2, 0, 1), None, stride=2, padding=1)
    return y[..., :2 * n, :2 * n] + b.view(1, -1, 1, 1)


def conv1x1(x, w_hwio, b):
    return F.conv2d(x, w_hwio.permute(3, 2, 0, 1), b)


def bn_frozen(x, gamma, beta):
    """BatchNormalization called without training=True: moving_mean=0, moving_var=1 forever."""
    s = gamma / math.sqrt(1.0 + BN_EPS)
    return x * s.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)


def lrelu(x):
    return F.leaky_relu(x, LRELU_ALPHA)


def _act_with_sign(h, slope, signs, key):
    """LeakyReLU (slope = alpha) / ReLU (slope = 0) of the NCHW pre-activation ``h``.  ``signs`` (test aid, default None = the
    literal activation): {key: bool array NHWC, True where the IMPLEMENTATION's pre-activation is positive}.  The activations are
    piecewise linear, so a pre-activation within rounding of 0 legitimately falls on either branch; evaluating the oracle on the
    implementation's branch removes that non-differentiability from the gradient comparison (same device as ``l1_sign``).  The
    fraction of elements on which the imposed pattern differs from the oracle's own is appended to signs['_mismatch']."""
    if signs is None or key not in signs:
        return torch.where(h > 0, h, slope * h)
    m = signs[key]
    m = m if isinstance(m, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(m))
    m = m.reshape(h.shape[0], h.shape[2], h.shape[3], h.shape[1]).permute(0, 3, 1, 2).to(torch.bool)
    signs.setdefault('_mismatch', []).append(float(((h.detach() > 0) != m).double().mean()))
    return torch.where(m, h, slope * h)


def dropout(x, mask, rate, training):
    """Keras Dropout: x*mask/(1-rate) when training.  mask is a {0,1} array supplied by the caller."""
    if not training or mask is None:
        return x
    return x * mask / (1.0 - rate)


# --------------------------------------------------------------------------- graphs
def _t(a, dtype):
    return a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def _bn_names(P, scope):
    ks = [k[:-len('/gamma')] for k in P if k.startswith(scope + '/batch_normalization') and k.endswith('/gamma')]
    return ks


def encoder(P, x, signs=None, tag=''):
    """build_unified_encoder applied (models/customlayers.py:16-24; autoencoder.py:12-17).  x is NCHW.
    signs / tag: optional activation branch patterns, keys 'enc<i><tag>' (see _act_with_sign)."""
    bns = _bn_names(P, 'Encoder')
    h = x
    i = 0
    while f'Encoder/enc_conv2D_{i}/kernel' in P:
        h = conv2d_same_s2(h, P[f'Encoder/enc_conv2D_{i}/kernel'], P[f'Encoder/enc_conv2D_{i}/bias'])
        h = bn_frozen(h, P[bns[i] + '/gamma'], P[bns[i] + '/beta'])
        h = _act_with_sign(h, LRELU_ALPHA, signs, f'enc{i}{tag}')
        i += 1
    return h


def decoder(P, h, signs=None, tag=''):
    """build_unified_decoder applied (models/customlayers.py:27-38).  h is NCHW.
    signs / tag: optional activation branch patterns, keys 'dec_entry<tag>', 'dec<i><tag>' (see _act_with_sign)."""
    bns = _bn_names(P, 'Decoder')
    h = _act_with_sign(bn_frozen(h, P[bns[0] + '/gamma'], P[bns[0] + '/beta']), 0.0, signs, f'dec_entry{tag}')
    i = 0
    while f'Decoder/dec_Conv2DT_{i}/kernel' in P:
        h = conv2dT_same_s2(h, P[f'Decoder/dec_Conv2DT_{i}/kernel'], P[f'Decoder/dec_Conv2DT_{i}/bias'])
        h = bn_frozen(h, P[bns[i + 1] + '/gamma'], P[bns[i + 1] + '/beta'])
        h = _act_with_sign(h, LRELU_ALPHA, signs, f'dec{i}{tag}')
        i += 1
    return conv1x1(h, P['Decoder/dec_Conv2D_final/kernel'], P['Decoder/dec_Conv2D_final/bias'])


def _flatten_nhwc(h):
    return h.permute(0, 2, 3, 1).reshape(h.shape[0], -1)   # Keras Flatten on NHWC


def _unflatten_nhwc(v, res, c):
    return v.reshape(v.shape[0], res, res, c).permute(0, 3, 1, 2)


def forward(arch, P, x, *, x_ce=None, eps=None, masks=None, dropout_rate=0.0, training=False, dtype=torch.float32, act_signs=None):
    """Restates models/autoencoder.py:9-40, variational_autoencoder.py:9-47,
    context_encoder_variational_autoencoder.py:9-59.

    x, x_ce: NHWC.  eps: [B,zDim] standard-normal draw replacing tf.random_normal (always live).
    masks: dict of {0,1} dropout masks: 'z' (AE) | 'mu','log_sigma','dec' (VAE) | + 'mu_ce','dec_ce' (ceVAE).
    act_signs: optional activation branch patterns of the x branch (keys 'enc<i>', 'dec_entry', 'dec<i>') and of the ceVAE x_ce branch
    (same keys + '_ce'); VAE / ceVAE graphs only.
    Returns a dict of NHWC / [B,z] torch tensors (autograd-connected to P if P holds leaf tensors).
    """
    masks = masks or {}
    P = {k: _t(v, dtype) for k, v in P.items()}
    xt = _t(x, dtype).permute(0, 3, 1, 2)
    out = {}
    h = encoder(P, xt, act_signs)

    def M(name):
        m = masks.get(name)
        return None if m is None else _t(m, dtype)

    if arch == AES:
        # models/autoencoder_spatial.py:12-25: z = Dropout(encoder(x), training=dropout) on the NHWC code; x_hat = decoder(z).
        # masks['z']: {0,1} array [B, res, res, C] (NHWC, as the tensor the Keras layer sees)
        m = M('z')
        zs = dropout(h.permute(0, 2, 3, 1), m, dropout_rate, training)
        out['z'] = zs
        out['x_hat'] = decoder(P, zs.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        return out
    h = conv1x1(h, P['Bottleneck/conv2d/kernel'], P['Bottleneck/conv2d/bias'])
    res, cb = h.shape[2], h.shape[1]
    flat = _flatten_nhwc(h)

    if arch in (AE, CAE):
        # autoencoder.py:29-30: dropout on z honours the flag; dropout on dec_dense(z) is called WITHOUT the flag -> identity.
        # constrained_autoencoder.py:29-30: both honour it (masks 'z', 'dec'), and x_hat is re-encoded by the SAME layers
        # (:42-46) with a third, independent Dropout draw (mask 'z_rec').
        z = dropout(flat @ P['Bottleneck/dense/kernel'] + P['Bottleneck/dense/bias'], M('z'), dropout_rate, training)
        out['z'] = z
        d = z @ P['Bottleneck/dense_1/kernel'] + P['Bottleneck/dense_1/bias']
        if arch == CAE:
            d = dropout(d, M('dec'), dropout_rate, training)
        h = conv1x1(_unflatten_nhwc(d, res, cb), P['Bottleneck/conv2d_1/kernel'], P['Bottleneck/conv2d_1/bias'])
        xh = decoder(P, h)
        out['x_hat'] = xh.permute(0, 2, 3, 1)
        if arch == CAE:
            h2 = conv1x1(encoder(P, xh), P['Bottleneck/conv2d/kernel'], P['Bottleneck/conv2d/bias'])
            out['z_rec'] = dropout(_flatten_nhwc(h2) @ P['Bottleneck/dense/kernel'] + P['Bottleneck/dense/bias'], M('z_rec'),
                                   dropout_rate, training)
        return out

    mu = dropout(flat @ P['Bottleneck/dense/kernel'] + P['Bottleneck/dense/bias'], M('mu'), dropout_rate, training)
    ls = dropout(flat @ P['Bottleneck/dense_1/kernel'] + P['Bottleneck/dense_1/bias'], M('log_sigma'), dropout_rate, training)
    sigma = torch.exp(ls)
    e = _t(eps, dtype)
    z = mu + e * sigma
    d = dropout(z @ P['Bottleneck/dense_2/kernel'] + P['Bottleneck/dense_2/bias'], M('dec'), dropout_rate, training)
    h = conv1x1(_unflatten_nhwc(d, res, cb), P['Bottleneck/conv2d_1/kernel'], P['Bottleneck/conv2d_1/bias'])
    out.update(z_mu=mu, z_log_sigma=ls, z_sigma=sigma, z=z)
    out['x_hat'] = decoder(P, h, act_signs).permute(0, 2, 3, 1)
    if arch == CEVAE:
        xc = _t(x_ce, dtype).permute(0, 3, 1, 2)
        hc = conv1x1(encoder(P, xc, act_signs, '_ce'), P['Bottleneck/conv2d/kernel'], P['Bottleneck/conv2d/bias'])
        mu_ce = dropout(_flatten_nhwc(hc) @ P['Bottleneck/dense/kernel'] + P['Bottleneck/dense/bias'], M('mu_ce'), dropout_rate, training)
        dc = dropout(mu_ce @ P['Bottleneck/dense_2/kernel'] + P['Bottleneck/dense_2/bias'], M('dec_ce'), dropout_rate, training)
        hc = conv1x1(_unflatten_nhwc(dc, res, cb), P['Bottleneck/conv2d_1/kernel'], P['Bottleneck/conv2d_1/bias'])
        out['z_mu_ce'] = mu_ce
        out['x_hat_ce'] = decoder(P, hc, act_signs, '_ce').permute(0, 2, 3, 1)
    return out


def losses(arch, out, x, x_ce=None, dtype=torch.float32, l1_sign=None, l1_sign_ce=None, rho=1.0):
    """trainers/AE.py:28-29, VAE.py:36-42, ceVAE.py:38-50.

    l1_sign / l1_sign_ce (test aid, default None = the literal |.|): evaluate |u| as u*sign with a CALLER-FIXED sign
    pattern.  |u| is not differentiable at 0 and a 1e-7 perturbation of x_hat flips d|u|/du = sign(u) on pixels where
    x_hat ~ x; fixing the pattern to the one of the implementation under test makes gradient comparisons measure the
    arithmetic instead of that discontinuity (tests assert separately that the patterns differ on almost no pixel)."""
    xt = _t(x, dtype)
    L = {}
    l1 = (out['x_hat'] - xt).abs() if l1_sign is None else (out['x_hat'] - xt) * _t(l1_sign, dtype)
    rec = l1.sum(dim=(1, 2, 3))
    if arch == CAE:
        # trainers/ConstrainedAE.py:37-43 (tf.losses.mean_squared_error with Reduction.NONE = squared difference)
        L['L1'] = l1
        L['reconstructionLoss'] = rec.mean()
        l2 = ((out['x_hat'] - xt) ** 2).mean(dim=(1, 2, 3))
        rec_z = ((out['z'] - out['z_rec']) ** 2).mean(dim=1)
        L['L2'], L['Rec_z'] = l2.mean(), rec_z.mean()
        L['loss'] = (l2 + rho * rec_z).mean()
        return L
    if arch in (AE, AES):
        L['L1'] = l1
        L['reconstructionLoss'] = L['loss'] = rec.mean()
        return L
    mu, sg = out['z_mu'], out['z_sigma']
    kl = 0.5 * (mu ** 2 + sg ** 2 - torch.log(sg ** 2) - 1).sum(dim=1)
    if arch == VAE:
        L['L1'] = l1
        L['reconstructionLoss'] = rec.mean()
        L['kl'] = kl.mean()
        L['loss'] = (rec + kl).mean()
        return L
    xc = _t(x_ce, dtype)
    l1c = (out['x_hat_ce'] - xc).abs() if l1_sign_ce is None else (out['x_hat_ce'] - xc) * _t(l1_sign_ce, dtype)
    recc = l1c.sum(dim=(1, 2, 3))
    L['L1_vae'], L['L1_ce'] = l1, l1c
    L['L1'] = 0.5 * (l1 + l1c)
    L['Rec_ce'], L['Rec_vae'] = recc.mean(), rec.mean()
    L['reconstructionLoss'] = 0.5 * (rec + recc).mean()
    L['kl'] = kl.mean()
    L['loss'] = (rec + kl + recc).mean()
    L['loss_vae'] = (rec + kl).mean()
    return L


def loss_and_grads(arch, P, x, *, x_ce=None, eps=None, masks=None, dropout_rate=0.0, training=True, dtype=torch.float32,
                   want_anomaly=False, l1_sign=None, l1_sign_ce=None, rho=1.0, act_signs=None):
    """tf.gradients of losses['loss'] w.r.t. every trainable variable (DLMODEL.py:112-131); ceVAE 'anomaly' (ceVAE.py:51)."""
    Pt = OrderedDict((k, _t(v, dtype).clone().requires_grad_(True)) for k, v in P.items())
    xt = _t(x, dtype).clone().requires_grad_(want_anomaly)
    out = forward(arch, Pt, xt, x_ce=x_ce, eps=eps, masks=masks, dropout_rate=dropout_rate, training=training, dtype=dtype,
                  act_signs=act_signs)
    L = losses(arch, out, xt, x_ce=x_ce, dtype=dtype, l1_sign=l1_sign, l1_sign_ce=l1_sign_ce, rho=rho)
    names = list(Pt.keys())
    grads = torch.autograd.grad(L['loss'], [Pt[k] for k in names], retain_graph=want_anomaly, allow_unused=True)
    G = OrderedDict((k, (g if g is not None else torch.zeros_like(Pt[k])).detach()) for k, g in zip(names, grads))
    if want_anomaly and arch == CEVAE:
        gx = torch.autograd.grad(L['loss_vae'], xt)[0]
        L['anomaly'] = (L['L1_vae'] * gx.abs()).detach()
    out = {k: v.detach() for k, v in out.items()}
    L = {k: v.detach() for k, v in L.items()}
    return out, L, G


def cevae_reconstruct(P, x, *, eps, use_gradient_based_restoration=True, l1_sign=None, dtype=torch.float32):
    """Restates reference trainers/ceVAE.py:119-144 as utils/Evaluation.py:246-253 drives it: ONE slice per ``sess.run``
    (feed x_ce = x, dropout off), fetches ``reconstruction`` and every loss incl. ``anomaly = L1_vae * |d loss_vae / d x|``
    (:51) where ``loss_vae = reduce_mean(rec_vae + kl)`` is a mean over that ONE sample; with a truthy
    ``use_gradient_based_restoration`` the returned reconstruction is ``x - lambda * anomaly`` (:136-139).
    x: [N,H,W,C]; eps: [N,zDim] (slice i uses eps[i]).  l1_sign: optional caller-fixed sign pattern (see ``losses``).
    Returns {'x_hat', 'anomaly', 'L1_vae', 'reconstruction'} as float64/float32 numpy arrays, slice by slice."""
    outs = {'x_hat': [], 'anomaly': [], 'L1_vae': [], 'reconstruction': []}
    for i in range(x.shape[0]):
        xi = _t(x[i:i + 1], dtype).clone().requires_grad_(True)
        o = forward(CEVAE, P, xi, x_ce=x[i:i + 1], eps=eps[i:i + 1], training=False, dtype=dtype)
        L = losses(CEVAE, o, xi, x_ce=x[i:i + 1], dtype=dtype, l1_sign=None if l1_sign is None else l1_sign[i:i + 1])
        gx = torch.autograd.grad(L['loss_vae'], xi)[0]
        an = (L['L1_vae'] * gx.abs()).detach()
        rec = o['x_hat'].detach()
        if use_gradient_based_restoration:
            rec = xi.detach() - use_gradient_based_restoration * an
        for k, v in (('x_hat', o['x_hat'].detach()), ('anomaly', an), ('L1_vae', L['L1_vae'].detach()), ('reconstruction', rec)):
            outs[k].append(v.numpy())
    return {k: np.concatenate(v, 0) for k, v in outs.items()}


def total_variation(d):
    """tf.image.total_variation on NHWC images (TF 1.15 image_ops_impl.py): sum |d[:,1:]-d[:,:-1]| + sum |d[:,:,1:]-d[:,:,:-1]|
    per image."""
    return (d[:, 1:] - d[:, :-1]).abs().sum(dim=(1, 2, 3)) + (d[:, :, 1:] - d[:, :, :-1]).abs().sum(dim=(1, 2, 3))


def restore_gradient(arch, P, x, *, eps=None, tv_lambda=0.0, masks=None, dropout_rate=0.0, training=False, dtype=torch.float32,
                     sign_from=None):
    """losses['grads'] of reference trainers/VAE_You.py:47-54:
        pixel_loss = sum_hwc |x_hat - x| + kl   (per sample);  restore = tv_lambda * total_variation(x - x_hat)
        grads = tf.gradients(pixel_loss + restore, x)[0]       (gradient of the SUM over samples; x also enters directly)
    sign_from (test aid, default None = literal |.|): an x_hat array (the implementation's) whose sign patterns
    sign(x_hat - x) and sign of the neighbour differences of x - x_hat replace d|u|/du, see ``losses``.
    Returns (grads NHWC, out dict, per-sample tv)."""
    xt = _t(x, dtype).clone().requires_grad_(True)
    out = forward(arch, P, xt, eps=eps, masks=masks, dropout_rate=dropout_rate, training=training, dtype=dtype)
    xh = out['x_hat']
    d = xt - xh
    if sign_from is None:
        rec = (xh - xt).abs().sum(dim=(1, 2, 3))
        tv = total_variation(d)
    else:
        xs = _t(x, dtype)
        ds = xs - _t(sign_from, dtype)
        rec = ((xh - xt) * torch.sign(-ds)).sum(dim=(1, 2, 3))
        tv = ((d[:, 1:] - d[:, :-1]) * torch.sign(ds[:, 1:] - ds[:, :-1])).sum(dim=(1, 2, 3)) + \
             ((d[:, :, 1:] - d[:, :, :-1]) * torch.sign(ds[:, :, 1:] - ds[:, :, :-1])).sum(dim=(1, 2, 3))
    total = rec + tv_lambda * tv
    if arch != AE:
        mu, sg = out['z_mu'], out['z_sigma']
        total = total + 0.5 * (mu ** 2 + sg ** 2 - torch.log(sg ** 2) - 1).sum(dim=1)
    g = torch.autograd.grad(total.sum(), xt)[0]
    return g.detach(), {k: v.detach() for k, v in out.items()}, tv.detach()


def restore(arch, P, x, *, steps, restore_lr, tv_lambda, eps_list, dtype=torch.float32):
    """VAE_You.reconstruct (trainers/VAE_You.py:125-139): restored -= restore_lr * grads, ``steps`` times, a fresh eps each run."""
    restored = np.array(x, dtype=np.float64 if dtype == torch.float64 else np.float32, copy=True)
    for k in range(steps):
        g, _, _ = restore_gradient(arch, P, restored, eps=eps_list[k], tv_lambda=tv_lambda, dtype=dtype)
        restored = restored - restore_lr * g.numpy()
    return restored


def adam_tf(P, G, m, v, t, lr, beta1=0.5, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer update: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps).
    t is the 1-based step count.  Works on dicts of torch tensors; returns new (P, m, v)."""
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    Pn, mn, vn = OrderedDict(), OrderedDict(), OrderedDict()
    for k in P:
        g = G[k]
        mn[k] = beta1 * m[k] + (1.0 - beta1) * g
        vn[k] = beta2 * v[k] + (1.0 - beta2) * g * g
        Pn[k] = P[k] - lr_t * mn[k] / (vn[k].sqrt() + eps)
    return Pn, mn, vn


class Trainer:
    """Minimal stateful train-step loop equal to AE/VAE/ceVAE.process(TRAIN) minus logging (trainers/AE.py:63-90)."""

    def __init__(self, arch, P, lr=1e-4, beta1=0.5, dropout_rate=0.2, dtype=torch.float32):
        self.arch, self.lr, self.beta1, self.rate, self.dtype = arch, lr, beta1, dropout_rate, dtype
        self.P = OrderedDict((k, _t(v, dtype).clone()) for k, v in P.items())
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in self.P.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in self.P.items())
        self.t = 0

    def step(self, x, x_ce=None, eps=None, masks=None, training=True):
        out, L, G = loss_and_grads(self.arch, self.P, x, x_ce=x_ce, eps=eps, masks=masks, dropout_rate=self.rate,
                                   training=training, dtype=self.dtype)
        self.t += 1
        self.P, self.m, self.v = adam_tf(self.P, G, self.m, self.v, self.t, self.lr, self.beta1)
        return out, L, G


# --------------------------------------------------------------------------- synthetic inputs
def synthetic_slices(B, S, C=1, seed=1234):
    """Synthetic BrainWeb-like slices: float32 NHWC in [0,1], ~50% exact zeros outside an elliptical 'brain'."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:S, 0:S].astype(np.float32)
    cy = cx = (S - 1) / 2.0
    out = np.zeros((B, S, S, C), np.float32)
    for b in range(B):
        r = 0.75 + 0.25 * math.sin(math.pi * (b % 110 + 0.5) / 110.0)
        mask = ((yy - cy) / (0.42 * S * r)) ** 2 + ((xx - cx) / (0.36 * S * r)) ** 2 <= 1.0
        for c in range(C):
            g = gaussian_filter(rng.uniform(size=(S, S)).astype(np.float32), sigma=S / 16.0)
            g = (g - g.min()) / max(float(g.max() - g.min()), 1e-12)
            n = rng.standard_normal((S, S)).astype(np.float32)
            img = np.clip(0.15 + 0.55 * g + 0.05 * n, 0.0, 1.0).astype(np.float32)
            out[b, :, :, c] = np.where(mask, img, 0.0)
    return out
