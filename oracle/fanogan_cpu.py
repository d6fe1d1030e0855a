"""fp32/fp64 torch-CPU restatement of the f-AnoGAN graph and its three train ops (oracle; PARITY UNPINNED; TEST INFRASTRUCTURE ONLY).

Restates models/fanogan.py:11-84: Encoder (unified encoder with BatchNorm + 1x1 conv + Dense + dropout + tanh), Generator
(Dense + dropout + 1x1 conv + unified decoder with LayerNormalization([1,2]) + sigmoid) and the Discriminator feature
stack (unified encoder with LayerNormalization) + Dense(1) on the channel axis, and the WGAN-GP / izi_f
losses and optimiser steps of trainers/fAnoGAN.py:50-77 (torch autograd, create_graph=True for the gradient penalty)."""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from .tf_graph_cpu import LRELU_ALPHA, _glorot, _t, bn_frozen, conv1x1, conv2d_same_s2, conv2dT_same_s2, dropout, stack_plan

LN_EPS = 1e-3


def _ln(k):
    return 'layer_normalization' if k == 0 else f'layer_normalization_{k}'


def _bn(k):
    return 'batch_normalization' if k == 0 else f'batch_normalization_{k}'


def init_params(S, C=1, zDim=128, res=8, seed=1):
    rng = np.random.default_rng(seed)
    n, enc, dec = stack_plan(S, res)
    P = OrderedDict()
    cin = C
    for i, co in enumerate(enc):
        P[f'Encoder/enc_conv2D_{i}/kernel'] = _glorot(rng, (5, 5, cin, co), 25 * cin, 25 * co)
        P[f'Encoder/enc_conv2D_{i}/bias'] = np.zeros(co, np.float32)
        P[f'Encoder/{_bn(i)}/gamma'] = np.ones(co, np.float32)
        P[f'Encoder/{_bn(i)}/beta'] = np.zeros(co, np.float32)
        cin = co
    cb = cin // 8
    flat = res * res * cb
    P['Encoder/conv2d/kernel'] = _glorot(rng, (1, 1, cin, cb), cin, cb)
    P['Encoder/conv2d/bias'] = np.zeros(cb, np.float32)
    P['Encoder/dense/kernel'] = _glorot(rng, (flat, zDim), flat, zDim)
    P['Encoder/dense/bias'] = np.zeros(zDim, np.float32)
    P['Generator/conv2d_1/kernel'] = _glorot(rng, (1, 1, cb, cin), cb, cin)
    P['Generator/conv2d_1/bias'] = np.zeros(cin, np.float32)
    P['Generator/dense_1/kernel'] = _glorot(rng, (zDim, flat), zDim, flat)
    P['Generator/dense_1/bias'] = np.zeros(flat, np.float32)
    ln = 0
    s = res
    P[f'Generator/{_ln(ln)}/gamma'] = np.ones((s, s), np.float32)
    P[f'Generator/{_ln(ln)}/beta'] = np.zeros((s, s), np.float32)
    ln += 1
    for i, co in enumerate(dec):
        P[f'Generator/dec_Conv2DT_{i}/kernel'] = _glorot(rng, (5, 5, co, cin), 25 * co, 25 * cin)
        P[f'Generator/dec_Conv2DT_{i}/bias'] = np.zeros(co, np.float32)
        s *= 2
        P[f'Generator/{_ln(ln)}/gamma'] = np.ones((s, s), np.float32)
        P[f'Generator/{_ln(ln)}/beta'] = np.zeros((s, s), np.float32)
        ln += 1
        cin = co
    P['Generator/dec_Conv2D_final/kernel'] = _glorot(rng, (1, 1, cin, C), cin, C)
    P['Generator/dec_Conv2D_final/bias'] = np.zeros(C, np.float32)
    cin, s = C, S
    for i, co in enumerate(enc):
        P[f'Discriminator/enc_conv2D_{i}/kernel'] = _glorot(rng, (5, 5, cin, co), 25 * cin, 25 * co)
        P[f'Discriminator/enc_conv2D_{i}/bias'] = np.zeros(co, np.float32)
        s //= 2
        P[f'Discriminator/{_ln(ln)}/gamma'] = np.ones((s, s), np.float32)
        P[f'Discriminator/{_ln(ln)}/beta'] = np.zeros((s, s), np.float32)
        ln += 1
        cin = co
    P['Discriminator/dense_2/kernel'] = _glorot(rng, (cin, 1), cin, 1)
    P['Discriminator/dense_2/bias'] = np.zeros(1, np.float32)
    return P


def perturb(P, seed=7, scale=0.05):
    rng = np.random.default_rng(seed)
    return OrderedDict((k, v.copy() if k.endswith('/kernel') else (v + scale * rng.standard_normal(v.shape)).astype(np.float32))
                       for k, v in P.items())


def layernorm_hw(x, gamma, beta):
    """tf.keras LayerNormalization(axis=[1,2]) on NCHW input: stats over (H,W) per (b,c), gamma/beta [H,W]."""
    mu = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * gamma[None, None] + beta[None, None]


def _act(n, sign, alpha):
    """LeakyReLU(alpha) / ReLU (alpha = 0).  ``sign`` (optional, NHWC {0,1} array = "pre-activation > 0") pins the sub-gradient
    branch per element: the derivative jumps at 0, so two fp32 evaluations whose pre-activations differ by round-off near 0
    legitimately pick different branches; parity tests pass the implementation's pattern (same device as l1_sign in
    tf_graph_cpu.losses)."""
    if sign is None:
        return F.leaky_relu(n, alpha) if alpha else F.relu(n)
    m = torch.as_tensor(np.asarray(sign)).to(torch.bool).permute(0, 3, 1, 2)
    return torch.where(m, n, alpha * n)


def _names(P, scope, stem):
    return [k[:-len('/gamma')] for k in P if k.startswith(scope + '/' + stem) and k.endswith('/gamma')]


def encode(P, x, mask=None, dropout_rate=0.0, training=False, dtype=torch.float32, signs=None):
    """fanogan.py:15-29 -> z_enc [B, zDim]."""
    P = {k: _t(v, dtype) for k, v in P.items()}
    h = _t(x, dtype).permute(0, 3, 1, 2)
    bns = _names(P, 'Encoder', 'batch_normalization')
    i = 0
    while f'Encoder/enc_conv2D_{i}/kernel' in P:
        h = _act(bn_frozen(conv2d_same_s2(h, P[f'Encoder/enc_conv2D_{i}/kernel'], P[f'Encoder/enc_conv2D_{i}/bias']),
                           P[bns[i] + '/gamma'], P[bns[i] + '/beta']), None if signs is None else signs[i], LRELU_ALPHA)
        i += 1
    h = conv1x1(h, P['Encoder/conv2d/kernel'], P['Encoder/conv2d/bias'])
    flat = h.permute(0, 2, 3, 1).reshape(h.shape[0], -1)
    z = flat @ P['Encoder/dense/kernel'] + P['Encoder/dense/bias']
    m = None if mask is None else _t(mask, dtype)
    return torch.tanh(dropout(z, m, dropout_rate, training))


def generate(P, z, mask=None, dropout_rate=0.0, training=False, dtype=torch.float32, signs=None):
    """fanogan.py:33-46 -> sigmoid(G(z)) as NHWC."""
    P = {k: _t(v, dtype) for k, v in P.items()}
    z = _t(z, dtype)
    lns = _names(P, 'Generator', 'layer_normalization')
    m = None if mask is None else _t(mask, dtype)
    d = dropout(z @ P['Generator/dense_1/kernel'] + P['Generator/dense_1/bias'], m, dropout_rate, training)
    cb = P['Generator/conv2d_1/kernel'].shape[2]
    res = int(round(math.sqrt(d.shape[1] // cb)))
    h = d.reshape(d.shape[0], res, res, cb).permute(0, 3, 1, 2)
    h = conv1x1(h, P['Generator/conv2d_1/kernel'], P['Generator/conv2d_1/bias'])
    h = _act(layernorm_hw(h, P[lns[0] + '/gamma'], P[lns[0] + '/beta']), None if signs is None else signs[0], 0.0)
    i = 0
    while f'Generator/dec_Conv2DT_{i}/kernel' in P:
        h = conv2dT_same_s2(h, P[f'Generator/dec_Conv2DT_{i}/kernel'], P[f'Generator/dec_Conv2DT_{i}/bias'])
        h = _act(layernorm_hw(h, P[lns[i + 1] + '/gamma'], P[lns[i + 1] + '/beta']), None if signs is None else signs[i + 1],
                 LRELU_ALPHA)
        i += 1
    h = conv1x1(h, P['Generator/dec_Conv2D_final/kernel'], P['Generator/dec_Conv2D_final/bias'])
    return torch.sigmoid(h).permute(0, 2, 3, 1)


def discriminate(P, x, dtype=torch.float32, signs=None):
    """fanogan.py:50-58 -> (features [B,r,r,128] NHWC, critic [B,r,r,1]: Dense(1) acts on the channel axis)."""
    P = {k: _t(v, dtype) for k, v in P.items()}
    h = _t(x, dtype).permute(0, 3, 1, 2)
    lns = _names(P, 'Discriminator', 'layer_normalization')
    i = 0
    while f'Discriminator/enc_conv2D_{i}/kernel' in P:
        h = conv2d_same_s2(h, P[f'Discriminator/enc_conv2D_{i}/kernel'], P[f'Discriminator/enc_conv2D_{i}/bias'])
        h = _act(layernorm_hw(h, P[lns[i] + '/gamma'], P[lns[i] + '/beta']), None if signs is None else signs[i], LRELU_ALPHA)
        i += 1
    f = h.permute(0, 2, 3, 1)
    return f, f @ P['Discriminator/dense_2/kernel'] + P['Discriminator/dense_2/bias']


def reconstruct(P, x, dtype=torch.float32):
    """trainers/fAnoGAN.py:220-239: x_enc = sigmoid(G(E(x))) with dropout off."""
    return generate(P, encode(P, x, dtype=dtype), dtype=dtype)


# --------------------------------------------------------------------------- training graph (trainers/fAnoGAN.py:50-77)
SCOPES = ('Encoder', 'Generator', 'Discriminator')


def as_leaves(P, dtype=torch.float32):
    """{name: ndarray} -> {name: leaf tensor requiring grad} (shared by every sub-graph of one step)."""
    return OrderedDict((k, _t(v, dtype).clone().requires_grad_(True)) for k, v in P.items())


def gradient_penalty(P, x_hat, scale, dtype=torch.float32, signs=None):
    """fAnoGAN.py:55-57: ddx = d sum(d_hat) / d x_hat; slopes = sqrt(sum(ddx^2, axis=1)) (axis 1 = H only, as written);
    gp = mean((slopes-1)^2)*scale.  x_hat is NHWC and must require grad."""
    _, d_hat = discriminate(P, x_hat, dtype, signs)
    ddx = torch.autograd.grad(d_hat.sum(), x_hat, create_graph=True)[0]
    slopes = torch.sqrt((ddx * ddx).sum(dim=1))
    return ((slopes - 1.0) ** 2).mean() * scale, ddx


def wgan_graph(P, x, z, alpha, mask_enc=None, mask_gen_z=None, mask_gen_enc=None, dropout_rate=0.0, training=True, scale=10.0,
               kappa=1.0, dtype=torch.float32, want=('gen', 'disc', 'enc'), signs=None):
    """All losses of fAnoGAN.train (fAnoGAN.py:50-66) on one feed.  P: dict of tensors (as_leaves).  alpha [B,1] is the
    tf.random_uniform draw of fanogan.py:67; the three masks are the three Dropout applications (Encoder z, dec_dense(z),
    dec_dense(z_enc)).  signs: optional {'gen_z','d_fake','d_real','d_hat','enc','gen_enc','d_enc'} -> per-level sign patterns."""
    x = _t(x, dtype)
    out = {}
    sg = (signs or {}).get
    if 'gen' in want or 'disc' in want:
        x_ = generate(P, z, mask_gen_z, dropout_rate, training, dtype, sg('gen_z'))
        _, d_ = discriminate(P, x_, dtype, sg('d_fake'))
        out['x_'] = x_
        out['disc_fake'] = d_.mean()
        out['gen_loss'] = -out['disc_fake']
    if 'disc' in want:
        _, d = discriminate(P, x, dtype, sg('d_real'))
        out['disc_real'] = d.mean()
        a = _t(alpha, dtype).reshape(-1, 1, 1, 1)
        x_hat = (x + a * (x_.detach() - x)).requires_grad_(True)        # only the critic weights receive this gradient
        # (the generator path into x_hat carries gradient in TF too, but disc_loss is minimised w.r.t. dis_vars only)
        gp, ddx = gradient_penalty(P, x_hat, scale, dtype, sg('d_hat'))
        out['x_hat'], out['ddx'], out['gp'] = x_hat, ddx, gp
        out['disc_loss'] = out['disc_fake'] - out['disc_real'] + gp
    if 'enc' in want:
        z_enc = encode(P, x, mask_enc, dropout_rate, training, dtype, sg('enc'))
        x_enc = generate(P, z_enc, mask_gen_enc, dropout_rate, training, dtype, sg('gen_enc'))
        f_enc, _ = discriminate(P, x_enc, dtype, sg('d_enc'))
        f_real, _ = discriminate(P, x, dtype, sg('d_real'))
        out['z_enc'], out['x_enc'] = z_enc, x_enc
        out['loss_img'] = ((x - x_enc) ** 2).mean(dim=(1, 2, 3)).mean()
        out['loss_fts'] = ((f_enc - f_real) ** 2).mean(dim=(1, 2, 3)).mean()
        out['enc_loss'] = out['loss_img'] + kappa * out['loss_fts']
        out['L1'] = (x_enc - x).abs()
        out['reconstructionLoss'] = out['L1'].sum(dim=(1, 2, 3)).mean()
    return out


def scope_grads(P, loss, scope):
    """tf.train.Optimizer.minimize(loss, var_list=[v for v in t_vars if scope in v.name]) gradients (fAnoGAN.py:71-77)."""
    names = [k for k in P if k.startswith(scope + '/')]
    gs = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True, retain_graph=True)
    return OrderedDict((k, torch.zeros_like(P[k]) if g is None else g.detach()) for k, g in zip(names, gs))


class WganTrainer:
    """The three AdamOptimizer(lr, beta1=0.5, beta2=0.9) train ops of fAnoGAN.py:75-77 with their own step counters."""

    def __init__(self, P, lr=1e-4, dropout_rate=0.0, scale=10.0, kappa=1.0, dtype=torch.float32):
        self.dtype, self.lr, self.rate, self.scale, self.kappa = dtype, lr, dropout_rate, scale, kappa
        self.P = OrderedDict((k, _t(v, dtype).clone()) for k, v in P.items())
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in self.P.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in self.P.items())
        self.t = {s: 0 for s in SCOPES}

    def _apply(self, scope, G):
        from .tf_graph_cpu import adam_tf
        self.t[scope] += 1
        names = list(G)
        Pn, mn, vn = adam_tf(OrderedDict((k, self.P[k]) for k in names), G, OrderedDict((k, self.m[k]) for k in names),
                             OrderedDict((k, self.v[k]) for k in names), self.t[scope], self.lr, 0.5, 0.9, 1e-8)
        for k in names:
            self.P[k], self.m[k], self.v[k] = Pn[k].detach(), mn[k], vn[k]

    def step(self, which, x, z, alpha=None, mask_enc=None, mask_gen=None, training=True, signs=None):
        """which in {'gen','disc','enc'}: one sess.run of optim_gen / optim_dis / optim_enc.  Returns (losses, grads)."""
        L = as_leaves(self.P, self.dtype)
        kw = dict(dropout_rate=self.rate, training=training, scale=self.scale, kappa=self.kappa, dtype=self.dtype, want=(which,),
                  signs=signs)
        if which == 'enc':
            out = wgan_graph(L, x, z, alpha, mask_enc=mask_enc, mask_gen_enc=mask_gen, **kw)
            loss, scope = out['enc_loss'], 'Encoder'
        else:
            out = wgan_graph(L, x, z, alpha, mask_gen_z=mask_gen, **kw)
            loss, scope = (out['gen_loss'], 'Generator') if which == 'gen' else (out['disc_loss'], 'Discriminator')
        G = scope_grads(L, loss, scope)
        self._apply(scope, G)
        return {k: v.detach() for k, v in out.items()}, G
