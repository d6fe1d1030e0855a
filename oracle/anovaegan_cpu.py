"""fp32/fp64 torch-CPU restatement of the AnoVAEGAN graph and its three train ops (oracle; PARITY UNPINNED; TEST INFRASTRUCTURE ONLY).

Restates models/anovaegan.py:10-83: a VAE whose decoder is the f-AnoGAN Generator stack (LayerNormalization([1,2]), NO final
sigmoid: outputs['out'] is the 1x1 conv's output, :50-54) with the f-AnoGAN critic on top, and the losses / optimisers of
trainers/AnoVAEGAN.py:50-83:
    optim_vae  minimises  enc_loss = mean_b sum|x - out| + kl_weight * mean_b kl      over Encoder + Generator variables
    optim_gen  minimises  gen_loss = -mean(D(out))                                    over Generator variables
    optim_dis  minimises  mean(D(out)) - mean(D(x)) + scale*mean((|dD(x_hat)/dx_hat|_{axis 1} - 1)^2)   over Discriminator variables
each a separate tf.train.AdamOptimizer(lr, beta1=0.5, beta2=0.9) - the Generator variables therefore own TWO sets of Adam
slots (one in optim_vae, one in optim_gen), as TensorFlow creates slots per optimizer instance.
TF variable names: the tf.layers / keras name counters run over the whole graph, so Encoder/{conv2d, dense (mu), dense_1
(log sigma)}, Generator/{conv2d_1, dense_2}, Discriminator/dense_3; LayerNormalization counters as in fanogan."""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from .fanogan_cpu import _act, _bn, _ln, _names, as_leaves, layernorm_hw
from .tf_graph_cpu import LRELU_ALPHA, _glorot, _t, adam_tf, bn_frozen, conv1x1, conv2d_same_s2, conv2dT_same_s2, dropout, stack_plan

SCOPES = ('Encoder', 'Generator', 'Discriminator')


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
    for name in ('Encoder/dense', 'Encoder/dense_1'):                    # mu_layer, sigma_layer (anovaegan.py:28-29)
        P[name + '/kernel'] = _glorot(rng, (flat, zDim), flat, zDim)
        P[name + '/bias'] = np.zeros(zDim, np.float32)
    P['Generator/conv2d_1/kernel'] = _glorot(rng, (1, 1, cb, cin), cb, cin)
    P['Generator/conv2d_1/bias'] = np.zeros(cin, np.float32)
    P['Generator/dense_2/kernel'] = _glorot(rng, (zDim, flat), zDim, flat)
    P['Generator/dense_2/bias'] = np.zeros(flat, np.float32)
    ln, s = 0, res
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
    P['Discriminator/dense_3/kernel'] = _glorot(rng, (cin, 1), cin, 1)
    P['Discriminator/dense_3/bias'] = np.zeros(1, np.float32)
    return P


def encode(P, x, eps, mask_mu=None, mask_ls=None, dropout_rate=0.0, training=False, dtype=torch.float32, signs=None):
    """anovaegan.py:14-35 -> (z_mu, z_log_sigma, z_sigma, z_vae).  P: dict of tensors."""
    h = _t(x, dtype).permute(0, 3, 1, 2)
    bns = _names(P, 'Encoder', 'batch_normalization')
    i = 0
    while f'Encoder/enc_conv2D_{i}/kernel' in P:
        h = _act(bn_frozen(conv2d_same_s2(h, P[f'Encoder/enc_conv2D_{i}/kernel'], P[f'Encoder/enc_conv2D_{i}/bias']),
                           P[bns[i] + '/gamma'], P[bns[i] + '/beta']), None if signs is None else signs[i], LRELU_ALPHA)
        i += 1
    h = conv1x1(h, P['Encoder/conv2d/kernel'], P['Encoder/conv2d/bias'])
    flat = h.permute(0, 2, 3, 1).reshape(h.shape[0], -1)
    mm = None if mask_mu is None else _t(mask_mu, dtype)
    ml = None if mask_ls is None else _t(mask_ls, dtype)
    z_mu = dropout(flat @ P['Encoder/dense/kernel'] + P['Encoder/dense/bias'], mm, dropout_rate, training)
    z_ls = dropout(flat @ P['Encoder/dense_1/kernel'] + P['Encoder/dense_1/bias'], ml, dropout_rate, training)
    z_sigma = torch.exp(z_ls)
    return z_mu, z_ls, z_sigma, z_mu + _t(eps, dtype) * z_sigma


def generate(P, z, mask=None, dropout_rate=0.0, training=False, dtype=torch.float32, signs=None):
    """anovaegan.py:37-54 -> out NHWC (no output non-linearity)."""
    lns = _names(P, 'Generator', 'layer_normalization')
    m = None if mask is None else _t(mask, dtype)
    d = dropout(z @ P['Generator/dense_2/kernel'] + P['Generator/dense_2/bias'], m, dropout_rate, training)
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
    return conv1x1(h, P['Generator/dec_Conv2D_final/kernel'], P['Generator/dec_Conv2D_final/bias']).permute(0, 2, 3, 1)


def discriminate(P, x, dtype=torch.float32, signs=None):
    """anovaegan.py:57-72 -> (features NHWC, critic [B,r,r,1]: Dense(1) acts on the channel axis)."""
    h = _t(x, dtype).permute(0, 3, 1, 2)
    lns = _names(P, 'Discriminator', 'layer_normalization')
    i = 0
    while f'Discriminator/enc_conv2D_{i}/kernel' in P:
        h = conv2d_same_s2(h, P[f'Discriminator/enc_conv2D_{i}/kernel'], P[f'Discriminator/enc_conv2D_{i}/bias'])
        h = _act(layernorm_hw(h, P[lns[i] + '/gamma'], P[lns[i] + '/beta']), None if signs is None else signs[i], LRELU_ALPHA)
        i += 1
    f = h.permute(0, 2, 3, 1)
    return f, f @ P['Discriminator/dense_3/kernel'] + P['Discriminator/dense_3/bias']


def graph(P, x, eps, alpha=None, masks=None, dropout_rate=0.0, training=True, scale=10.0, kl_weight=1.0, dtype=torch.float32,
          want=('vae', 'gen', 'disc'), signs=None, l1_sign=None):
    """All losses of AnoVAEGAN.train (AnoVAEGAN.py:50-71) on one feed.  P: as_leaves(...).  masks: {'mu','ls','dec'} Dropout
    masks (one Dropout layer object, three applications); alpha [B,1]: the tf.random_uniform draw of anovaegan.py:75.
    l1_sign (optional, NHWC in {-1,0,1}): pins the sub-gradient of |x - out| (see tf_graph_cpu.losses)."""
    x = _t(x, dtype)
    mk = (masks or {}).get
    sg = (signs or {}).get
    o = {}
    z_mu, z_ls, z_sigma, z = encode(P, x, eps, mk('mu'), mk('ls'), dropout_rate, training, dtype, sg('enc'))
    out = generate(P, z, mk('dec'), dropout_rate, training, dtype, sg('gen'))
    o.update(z_mu=z_mu, z_log_sigma=z_ls, z_sigma=z_sigma, out=out)
    if 'vae' in want:
        kl = 0.5 * (z_mu ** 2 + z_sigma ** 2 - torch.log(z_sigma ** 2) - 1).sum(dim=1)
        o['kl'] = kl.mean()
        diff = out - x
        l1 = diff.abs() if l1_sign is None else diff * _t(l1_sign, dtype)      # caller-fixed sign: see tf_graph_cpu.losses
        o['L1'] = l1
        o['reconstructionLoss'] = o['loss'] = l1.sum(dim=(1, 2, 3)).mean()
        o['enc_loss'] = o['reconstructionLoss'] + kl_weight * o['kl']
        o['loss_img'] = ((x - out) ** 2).mean(dim=(1, 2, 3)).mean()
    if 'gen' in want or 'disc' in want:
        f_fake, d_ = discriminate(P, out, dtype, sg('d_fake'))
        o['disc_fake'] = d_.mean()
        o['gen_loss'] = -o['disc_fake']
    if 'disc' in want:
        f_real, d = discriminate(P, x, dtype, sg('d_real'))
        o['disc_real'] = d.mean()
        o['loss_fts'] = ((f_fake - f_real) ** 2).mean(dim=(1, 2, 3)).mean()
        a = _t(alpha, dtype).reshape(-1, 1, 1, 1)
        x_hat = (x + a * (out.detach() - x)).requires_grad_(True)       # disc_loss is minimised over the critic's variables only
        _, d_hat = discriminate(P, x_hat, dtype, sg('d_hat'))
        ddx = torch.autograd.grad(d_hat.sum(), x_hat, create_graph=True)[0]
        slopes = torch.sqrt((ddx * ddx).sum(dim=1))                     # axis 1 only, as the reference writes it (:56)
        o['gp'] = ((slopes - 1.0) ** 2).mean() * scale
        o['x_hat'], o['ddx'] = x_hat, ddx
        o['disc_loss'] = o['disc_fake'] - o['disc_real'] + o['gp']
    return o


class Trainer:
    """The three train ops with TensorFlow's slot ownership: optim_vae holds (m, v, t) for Encoder + Generator, optim_gen a
    second (m, v, t) for Generator, optim_dis one for Discriminator."""
    OPS = {'vae': ('Encoder', 'Generator'), 'gen': ('Generator',), 'disc': ('Discriminator',)}

    def __init__(self, P, lr=1e-4, dropout_rate=0.0, scale=10.0, kl_weight=1.0, dtype=torch.float32):
        self.dtype, self.lr, self.rate, self.scale, self.kl_weight = dtype, lr, dropout_rate, scale, kl_weight
        self.P = OrderedDict((k, _t(v, dtype).clone()) for k, v in P.items())
        self.slots = {}
        for op, scopes in self.OPS.items():
            names = [k for k in self.P if k.split('/')[0] in scopes]
            self.slots[op] = dict(names=names, m=OrderedDict((k, torch.zeros_like(self.P[k])) for k in names),
                                  v=OrderedDict((k, torch.zeros_like(self.P[k])) for k in names), t=0)

    def step(self, which, x, eps, alpha=None, masks=None, training=True, signs=None, l1_sign=None):
        L = as_leaves(self.P, self.dtype)
        o = graph(L, x, eps, alpha, masks, self.rate, training, self.scale, self.kl_weight, self.dtype, want=(which,), signs=signs,
                  l1_sign=l1_sign)
        loss = {'vae': 'enc_loss', 'gen': 'gen_loss', 'disc': 'disc_loss'}[which]
        sl = self.slots[which]
        gs = torch.autograd.grad(o[loss], [L[k] for k in sl['names']], allow_unused=True)
        G = OrderedDict((k, torch.zeros_like(L[k]) if g is None else g.detach()) for k, g in zip(sl['names'], gs))
        sl['t'] += 1
        Pn, mn, vn = adam_tf(OrderedDict((k, self.P[k]) for k in sl['names']), G, sl['m'], sl['v'], sl['t'], self.lr, 0.5, 0.9, 1e-8)
        for k in sl['names']:
            self.P[k], sl['m'][k], sl['v'][k] = Pn[k].detach(), mn[k], vn[k]
        return {k: v.detach() for k, v in o.items()}, G
