"""fp32/fp64 torch-CPU restatement of the f-AnoGAN graph's FORWARD paths (oracle; PARITY UNPINNED; TEST INFRASTRUCTURE ONLY).

Restates models/fanogan.py:11-84: Encoder (unified encoder with BatchNorm + 1x1 conv + Dense + dropout + tanh), Generator
(Dense + dropout + 1x1 conv + unified decoder with LayerNormalization([1,2]) + sigmoid) and the Discriminator feature
stack (unified encoder with LayerNormalization) + Dense(1) on the channel axis.  The WGAN-GP training
graph (trainers/fAnoGAN.py:50-77) is NOT restated yet."""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from .tf_graph_cpu import (LRELU_ALPHA, _glorot, _t, bn_frozen, conv1x1, conv2d_same_s2, conv2dT_same_s2, dropout, lrelu,
                           stack_plan)

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


def _names(P, scope, stem):
    return [k[:-len('/gamma')] for k in P if k.startswith(scope + '/' + stem) and k.endswith('/gamma')]


def encode(P, x, mask=None, dropout_rate=0.0, training=False, dtype=torch.float32):
    """fanogan.py:15-29 -> z_enc [B, zDim]."""
    P = {k: _t(v, dtype) for k, v in P.items()}
    h = _t(x, dtype).permute(0, 3, 1, 2)
    bns = _names(P, 'Encoder', 'batch_normalization')
    i = 0
    while f'Encoder/enc_conv2D_{i}/kernel' in P:
        h = lrelu(bn_frozen(conv2d_same_s2(h, P[f'Encoder/enc_conv2D_{i}/kernel'], P[f'Encoder/enc_conv2D_{i}/bias']),
                            P[bns[i] + '/gamma'], P[bns[i] + '/beta']))
        i += 1
    h = conv1x1(h, P['Encoder/conv2d/kernel'], P['Encoder/conv2d/bias'])
    flat = h.permute(0, 2, 3, 1).reshape(h.shape[0], -1)
    z = flat @ P['Encoder/dense/kernel'] + P['Encoder/dense/bias']
    m = None if mask is None else _t(mask, dtype)
    return torch.tanh(dropout(z, m, dropout_rate, training))


def generate(P, z, mask=None, dropout_rate=0.0, training=False, dtype=torch.float32):
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
    h = F.relu(layernorm_hw(h, P[lns[0] + '/gamma'], P[lns[0] + '/beta']))
    i = 0
    while f'Generator/dec_Conv2DT_{i}/kernel' in P:
        h = conv2dT_same_s2(h, P[f'Generator/dec_Conv2DT_{i}/kernel'], P[f'Generator/dec_Conv2DT_{i}/bias'])
        h = lrelu(layernorm_hw(h, P[lns[i + 1] + '/gamma'], P[lns[i + 1] + '/beta']))
        i += 1
    h = conv1x1(h, P['Generator/dec_Conv2D_final/kernel'], P['Generator/dec_Conv2D_final/bias'])
    return torch.sigmoid(h).permute(0, 2, 3, 1)


def discriminate(P, x, dtype=torch.float32):
    """fanogan.py:50-58 -> (features [B,r,r,128] NHWC, critic [B,r,r,1]: Dense(1) acts on the channel axis)."""
    P = {k: _t(v, dtype) for k, v in P.items()}
    h = _t(x, dtype).permute(0, 3, 1, 2)
    lns = _names(P, 'Discriminator', 'layer_normalization')
    i = 0
    while f'Discriminator/enc_conv2D_{i}/kernel' in P:
        h = conv2d_same_s2(h, P[f'Discriminator/enc_conv2D_{i}/kernel'], P[f'Discriminator/enc_conv2D_{i}/bias'])
        h = lrelu(layernorm_hw(h, P[lns[i] + '/gamma'], P[lns[i] + '/beta']))
        i += 1
    f = h.permute(0, 2, 3, 1)
    return f, f @ P['Discriminator/dense_2/kernel'] + P['Discriminator/dense_2/bias']


def reconstruct(P, x, dtype=torch.float32):
    """trainers/fAnoGAN.py:220-239: x_enc = sigmoid(G(E(x))) with dropout off."""
    return generate(P, encode(P, x, dtype=dtype), dtype=dtype)
