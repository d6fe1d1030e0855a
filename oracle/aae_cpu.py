"""fp32/fp64 torch-CPU restatement of the adversarial autoencoder and its three train ops (oracle; PARITY UNPINNED; TEST
INFRASTRUCTURE ONLY).

Restates models/adversarial_autoencoder.py:10-73 - the dense autoencoder (Encoder | Bottleneck | Decoder scopes; both bottleneck
Dropout calls honour the flag, :30-31) plus a latent critic Dense(50, leaky_relu) -> Dense(50, leaky_relu) -> Dense(1)
(tf.nn.leaky_relu: alpha = 0.2) run on z_ (fake = the code of x), z (real = the fed N(0,1) sample) and
z_hat = z + epsilon * (z - z_) (:64-65, as written) - and trainers/AAE.py:41-69:
    optim_ae   minimises  loss = mean_b mean_hwc (x - x_hat)^2                 over ALL trainable variables (the critic's get no
               gradient from it: TensorFlow skips them, their Adam slots in this optimizer are never touched)
    optim_dis  minimises  mean(D(z_)) - mean(D(z)) + mean((|dD(z_hat)/dz_hat|_2 - 1)^2 * scale)     over Discriminator variables
    optim_gen  minimises  -mean(D(z_))                                          over ENCODER-scope variables only (:64; the
               Bottleneck layers between the encoder and z_ carry the gradient but are not updated)
each its own tf.train.AdamOptimizer(lr, beta1=0.5, beta2=0.9).

``constrained=True`` restates models/constrained_adversarial_autoencoder.py:10-79 + trainers/ConstrainedAAE.py:44-72: x_hat is
re-encoded by the same layers (z_rec), loss = mean_b(L2 + rho * mean_j (z_rec - z_)^2), the Dropout calls on dec_dense(z_) and on
z_rec carry no flag (identity), the critic is 100-50-1, and the 'Encoder' scope that optim_gen selects by substring also holds the
1x1 bottleneck conv, the latent Dense and dec_dense (which receives no gradient from gen_loss) - variable names here stay the
canonical Encoder/ | Bottleneck/ | Decoder/ ones of the dense AE."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from .tf_graph_cpu import AE, _glorot, _t, adam_tf, conv1x1, decoder, dropout, encoder, init_params as ae_init_params
from .tf_graph_cpu import _flatten_nhwc, _unflatten_nhwc

CRITIC = (50, 50, 1)
CRITIC_CONSTRAINED = (100, 50, 1)
CRITIC_ALPHA = 0.2


def init_params(S, C=1, zDim=128, res=8, seed=1, constrained=False):
    P = ae_init_params(AE, S, C, zDim, res, seed)                  # same variables / names as the dense AE
    rng = np.random.default_rng(seed + 1000)
    k = zDim
    for j, width in enumerate(CRITIC_CONSTRAINED if constrained else CRITIC):
        P[f'Discriminator/dense_{2 + j}/kernel'] = _glorot(rng, (k, width), k, width)
        P[f'Discriminator/dense_{2 + j}/bias'] = np.zeros(width, np.float32)
        k = width
    return P


def perturb(P, seed=7, scale=0.05):
    rng = np.random.default_rng(seed)
    return OrderedDict((k, v.copy() if k.endswith('/kernel') else (v + scale * rng.standard_normal(v.shape)).astype(np.float32))
                       for k, v in P.items())


def as_leaves(P, dtype=torch.float32):
    return OrderedDict((k, _t(v, dtype).clone().requires_grad_(True)) for k, v in P.items())


def encode(P, x, mask_z=None, dropout_rate=0.0, training=False, dtype=torch.float32):
    """x -> z_ (adversarial_autoencoder.py:13-30); also returns the bottleneck geometry for decode()."""
    h = conv1x1(encoder(P, _t(x, dtype).permute(0, 3, 1, 2)), P['Bottleneck/conv2d/kernel'], P['Bottleneck/conv2d/bias'])
    m = None if mask_z is None else _t(mask_z, dtype)
    z_ = dropout(_flatten_nhwc(h) @ P['Bottleneck/dense/kernel'] + P['Bottleneck/dense/bias'], m, dropout_rate, training)
    return z_, (h.shape[2], h.shape[1])


def decode(P, z_, geom, mask_dec=None, dropout_rate=0.0, training=False, dtype=torch.float32):
    m = None if mask_dec is None else _t(mask_dec, dtype)
    d = dropout(z_ @ P['Bottleneck/dense_1/kernel'] + P['Bottleneck/dense_1/bias'], m, dropout_rate, training)
    h = conv1x1(_unflatten_nhwc(d, *geom), P['Bottleneck/conv2d_1/kernel'], P['Bottleneck/conv2d_1/bias'])
    return decoder(P, h).permute(0, 2, 3, 1)


def critic(P, z, signs=None):
    """[B, zDim] -> [B, 1].  signs (optional): per hidden layer the {0,1} pattern "pre-activation > 0" that pins the
    leaky-ReLU branch (see fanogan_cpu._act)."""
    h = z
    n_layers = sum(1 for k in P if k.startswith('Discriminator/') and k.endswith('/kernel'))
    for j in range(n_layers):
        h = h @ P[f'Discriminator/dense_{2 + j}/kernel'] + P[f'Discriminator/dense_{2 + j}/bias']
        if j < n_layers - 1:
            if signs is None:
                h = F.leaky_relu(h, CRITIC_ALPHA)
            else:
                h = torch.where(torch.as_tensor(np.asarray(signs[j])).to(torch.bool), h, CRITIC_ALPHA * h)
    return h


def graph(P, x, z, epsilon=None, masks=None, dropout_rate=0.0, training=True, scale=10.0, dtype=torch.float32,
          want=('ae', 'gen', 'disc'), signs=None, constrained=False, rho=1.0):
    """All losses of AAE.train (AAE.py:41-57) on one feed.  P: as_leaves(...).  masks: {'z','dec'}; epsilon [B,1]: the
    tf.random_uniform draw of adversarial_autoencoder.py:64; signs: {'d_fake','d_real','d_hat'} -> critic sign patterns."""
    x, z = _t(x, dtype), _t(z, dtype)
    mk = (masks or {}).get
    sg = (signs or {}).get
    o = {}
    z_, geom = encode(P, x, mk('z'), dropout_rate, training, dtype)
    o['z_'] = z_
    if 'ae' in want:
        x_hat = decode(P, z_, geom, None if constrained else mk('dec'), dropout_rate, training, dtype)
        o['x_hat'] = x_hat
        o['L1'] = (x_hat - x).abs()
        o['reconstructionLoss'] = o['L1'].sum(dim=(1, 2, 3)).mean()
        o['L2'] = ((x - x_hat) ** 2).mean(dim=(1, 2, 3))
        o['loss'] = o['L2'].mean()
        if constrained:
            o['z_rec'], _ = encode(P, x_hat, None, dropout_rate, training, dtype)
            o['Rec_z'] = ((o['z_rec'] - z_) ** 2).mean(dim=1)
            o['loss'] = (o['L2'] + rho * o['Rec_z']).mean()
    if 'gen' in want or 'disc' in want:
        o['disc_fake'] = critic(P, z_, sg('d_fake')).mean()
        o['gen_loss'] = -o['disc_fake']
    if 'disc' in want:
        o['disc_real'] = critic(P, z, sg('d_real')).mean()
        o['disc_loss_without_grad'] = o['disc_fake'] - o['disc_real']
        e = _t(epsilon, dtype).reshape(-1, 1)
        z_hat = (z + e * (z - z_.detach())).requires_grad_(True)          # minimised over the critic's variables only
        d_hat = critic(P, z_hat, sg('d_hat'))
        ddz = torch.autograd.grad(d_hat.sum(), z_hat, create_graph=True)[0]
        slopes = torch.sqrt((ddz * ddz).sum(dim=1))
        o['gp'] = ((slopes - 1.0) ** 2 * scale).mean()
        o['z_hat'], o['ddz'] = z_hat, ddz
        o['disc_loss'] = o['disc_loss_without_grad'] + o['gp']
    return o


class Trainer:
    OPS = {'ae': ('Encoder', 'Bottleneck', 'Decoder', 'Discriminator'), 'disc': ('Discriminator',), 'gen': ('Encoder',)}
    LOSS = {'ae': 'loss', 'disc': 'disc_loss', 'gen': 'gen_loss'}

    def __init__(self, P, lr=1e-4, dropout_rate=0.0, scale=10.0, dtype=torch.float32, constrained=False, rho=1.0):
        self.dtype, self.lr, self.rate, self.scale, self.constrained, self.rho = dtype, lr, dropout_rate, scale, constrained, rho
        self.P = OrderedDict((k, _t(v, dtype).clone()) for k, v in P.items())
        self.slots = {}
        for op, scopes in self.OPS.items():
            names = [k for k in self.P if k.split('/')[0] in scopes]
            if constrained and op == 'gen':      # the reference's 'Encoder' scope: + conv2d, dense (z_layer), dense_1 (dec_dense)
                names = [k for k in self.P if k.split('/')[0] == 'Encoder' or k.rsplit('/', 1)[0] in
                         ('Bottleneck/conv2d', 'Bottleneck/dense', 'Bottleneck/dense_1')]
            self.slots[op] = dict(names=names, m=OrderedDict((k, torch.zeros_like(self.P[k])) for k in names),
                                  v=OrderedDict((k, torch.zeros_like(self.P[k])) for k in names), t=0)

    def step(self, which, x, z, epsilon=None, masks=None, training=True, signs=None):
        L = as_leaves(self.P, self.dtype)
        o = graph(L, x, z, epsilon, masks, self.rate, training, self.scale, self.dtype, want=(which,), signs=signs,
                  constrained=self.constrained, rho=self.rho)
        sl = self.slots[which]
        gs = torch.autograd.grad(o[self.LOSS[which]], [L[k] for k in sl['names']], allow_unused=True)
        used = [k for k, g in zip(sl['names'], gs) if g is not None]      # compute_gradients drops (None, var) pairs
        G = OrderedDict((k, g.detach()) for k, g in zip(sl['names'], gs) if g is not None)
        sl['t'] += 1
        Pn, mn, vn = adam_tf(OrderedDict((k, self.P[k]) for k in used), G, OrderedDict((k, sl['m'][k]) for k in used),
                             OrderedDict((k, sl['v'][k]) for k in used), sl['t'], self.lr, 0.5, 0.9, 1e-8)
        for k in used:
            self.P[k], sl['m'][k], sl['v'][k] = Pn[k].detach(), mn[k], vn[k]
        return {k: v.detach() for k, v in o.items()}, G
