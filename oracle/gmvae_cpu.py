"""fp32/fp64 torch-CPU restatement of the Gaussian-mixture VAE (oracle; PARITY UNPINNED; TEST INFRASTRUCTURE ONLY).

Restates models/gaussian_mixture_variational_autoencoder.py:11-73 and the loss / restoration graph of trainers/GMVAE.py:58-92:
    Encoder -> 1x1 conv -> flatten -> four Dense heads  w_mu, w_log_sigma (dropout with the flag), z_mu (dropout with the flag),
    z_log_sigma (Dropout called WITHOUT the flag: identity, :41);  w = w_mu + eps_w * exp(0.5 w_log_sigma), z likewise;
    dec_dense(z) (dropout with the flag) -> 1x1 conv -> Decoder -> xz_mu;
    p(z|w,c): z_wc_mu = Dense(dim_z*dim_c)(w), z_wc_log_sigma_inv = Dense(dim_z*dim_c)(w) + a trainable bias initialised to 0.1,
    both reshaped [B, dim_z, dim_c];  pc = softmax_c(sum_j loglh).
    loss = mean_b sum|x - xz_mu| + mean_b con + mean_b w_loss + mean_b max(closs1, c_lambda)        (GMVAE.py:60-88)
    grads = d/dx sum_b [ loss + tv_lambda * TV(x - xz_mu)_b ]                                       (:89-90; `loss` is a scalar
    broadcast over the per-image TV vector, so the batch mean is multiplied back by B: per-sample sums)
Variable names (a convention): Bottleneck/{conv2d, conv2d_1, dense (w_mu), dense_1 (w_log_sigma), dense_2 (z_mu),
dense_3 (z_log_sigma), dense_4 (dec_dense)}, then the un-scoped dense_5 (z_wc_mu), dense_6 (z_wc_log_sigma) and `Variable` (the 0.1 bias)."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from .tf_graph_cpu import VAE, _flatten_nhwc, _glorot, _t, _unflatten_nhwc, conv1x1, decoder, dropout, encoder, total_variation
from .tf_graph_cpu import init_params as ae_init_params


def init_params(S, C=1, dim_z=128, dim_w=1, dim_c=9, res=8, seed=1):
    base = ae_init_params(VAE, S, C, dim_z, res, seed)
    rng = np.random.default_rng(seed + 2000)
    flat = base['Bottleneck/dense/kernel'].shape[0]
    P = OrderedDict()
    for k, v in base.items():
        if k.startswith('Bottleneck/dense'):
            continue
        if k.startswith('Decoder/') and 'Bottleneck/dense/kernel' not in P:      # the Dense heads follow the two 1x1 convs
            for name, shape in (('dense', (flat, dim_w)), ('dense_1', (flat, dim_w)), ('dense_2', (flat, dim_z)),
                                ('dense_3', (flat, dim_z)), ('dense_4', (dim_z, flat))):
                P[f'Bottleneck/{name}/kernel'] = _glorot(rng, shape, shape[0], shape[1])
                P[f'Bottleneck/{name}/bias'] = np.zeros(shape[1], np.float32)
        P[k] = v
    n = dim_z * dim_c
    for name in ('dense_5', 'dense_6'):
        P[name + '/kernel'] = _glorot(rng, (dim_w, n), dim_w, n)
        P[name + '/bias'] = np.zeros(n, np.float32)
    P['Variable'] = np.full(n, 0.1, np.float32)
    return P


def perturb(P, seed=7, scale=0.05):
    rng = np.random.default_rng(seed)
    return OrderedDict((k, v.copy() if k.endswith('/kernel') else (v + scale * rng.standard_normal(v.shape)).astype(np.float32))
                       for k, v in P.items())


def forward(P, x, eps_w, eps_z, masks=None, dropout_rate=0.0, training=False, dim_c=9, dtype=torch.float32):
    """masks: {'w_mu','w_ls','z_mu','dec'} {0,1} arrays (the four Dropout applications that honour the flag)."""
    P = {k: _t(v, dtype) for k, v in P.items()}
    mk = (masks or {}).get
    M_ = lambda n: None if mk(n) is None else _t(mk(n), dtype)      # noqa: E731
    h = conv1x1(encoder(P, _t(x, dtype).permute(0, 3, 1, 2)), P['Bottleneck/conv2d/kernel'], P['Bottleneck/conv2d/bias'])
    res, cb = h.shape[2], h.shape[1]
    flat = _flatten_nhwc(h)
    o = {}
    o['w_mu'] = dropout(flat @ P['Bottleneck/dense/kernel'] + P['Bottleneck/dense/bias'], M_('w_mu'), dropout_rate, training)
    o['w_log_sigma'] = dropout(flat @ P['Bottleneck/dense_1/kernel'] + P['Bottleneck/dense_1/bias'], M_('w_ls'), dropout_rate, training)
    o['w_sampled'] = o['w_mu'] + _t(eps_w, dtype) * torch.exp(0.5 * o['w_log_sigma'])
    o['z_mu'] = dropout(flat @ P['Bottleneck/dense_2/kernel'] + P['Bottleneck/dense_2/bias'], M_('z_mu'), dropout_rate, training)
    o['z_log_sigma'] = flat @ P['Bottleneck/dense_3/kernel'] + P['Bottleneck/dense_3/bias']
    o['z_sampled'] = o['z_mu'] + _t(eps_z, dtype) * torch.exp(0.5 * o['z_log_sigma'])
    d = dropout(o['z_sampled'] @ P['Bottleneck/dense_4/kernel'] + P['Bottleneck/dense_4/bias'], M_('dec'), dropout_rate, training)
    h = conv1x1(_unflatten_nhwc(d, res, cb), P['Bottleneck/conv2d_1/kernel'], P['Bottleneck/conv2d_1/bias'])
    dz = o['z_mu'].shape[1]
    o['z_wc_mus'] = (o['w_sampled'] @ P['dense_5/kernel'] + P['dense_5/bias']).reshape(-1, dz, dim_c)
    o['z_wc_log_sigma_invs'] = (o['w_sampled'] @ P['dense_6/kernel'] + P['dense_6/bias'] + P['Variable']).reshape(-1, dz, dim_c)
    o['xz_mu'] = decoder(P, h).permute(0, 2, 3, 1)
    z_sample = o['z_sampled'].unsqueeze(-1).expand(-1, -1, dim_c)
    loglh = -0.5 * ((z_sample - o['z_wc_mus']) ** 2 * torch.exp(o['z_wc_log_sigma_invs'])) - o['z_wc_log_sigma_invs'] + np.log(np.pi)
    o['pc_logit'] = loglh.sum(1)
    o['pc'] = torch.softmax(o['pc_logit'], dim=-1)
    return o


def losses(o, x, dim_c=9, c_lambda=1.0, dtype=torch.float32, l1_sign=None):
    xt = _t(x, dtype)
    L = {}
    diff = o['xz_mu'] - xt
    L['L1'] = diff.abs() if l1_sign is None else diff * _t(l1_sign, dtype)
    L['reconstructionLoss'] = L['mean_p_loss'] = L['L1'].sum(dim=(1, 2, 3)).mean()
    zmu = o['z_mu'].unsqueeze(-1).expand(-1, -1, dim_c)
    zlv = o['z_log_sigma'].unsqueeze(-1).expand(-1, -1, dim_c)
    d_var = (torch.exp(zlv) + (zmu - o['z_wc_mus']) ** 2) * (torch.exp(o['z_wc_log_sigma_invs']) + 1e-6)
    kl = (d_var - (o['z_wc_log_sigma_invs'] + zlv) - 1) * 0.5
    L['conditional_prior_loss'] = torch.matmul(kl, o['pc'].unsqueeze(-1)).squeeze(-1).sum(1).mean()
    L['w_prior_loss'] = (0.5 * (o['w_mu'] ** 2 + torch.exp(o['w_log_sigma']) - o['w_log_sigma'] - 1).sum(1)).mean()
    closs1 = (o['pc'] * torch.log(o['pc'] * dim_c + 1e-8)).sum(1)
    L['c_prior_loss'] = torch.maximum(closs1, torch.full_like(closs1, c_lambda)).mean()
    L['loss'] = L['mean_p_loss'] + L['conditional_prior_loss'] + L['w_prior_loss'] + L['c_prior_loss']
    return L


def loss_and_grads(P, x, eps_w, eps_z, masks=None, dropout_rate=0.0, training=True, dim_c=9, c_lambda=1.0, dtype=torch.float32,
                   l1_sign=None):
    Pt = OrderedDict((k, _t(v, dtype).clone().requires_grad_(True)) for k, v in P.items())
    o = forward(Pt, x, eps_w, eps_z, masks, dropout_rate, training, dim_c, dtype)
    L = losses(o, x, dim_c, c_lambda, dtype, l1_sign)
    gs = torch.autograd.grad(L['loss'], list(Pt.values()))
    return ({k: v.detach() for k, v in o.items()}, {k: v.detach() for k, v in L.items()},
            OrderedDict((k, g.detach()) for k, g in zip(Pt, gs)))


def restore_gradient(P, x, eps_w, eps_z, tv_lambda, dim_c=9, c_lambda=1.0, dtype=torch.float64, l1_sign=None, tv_sign=None):
    """losses['grads'] of GMVAE.py:89-90 for a batch: d/dx sum_b (loss + tv_lambda * TV(x - xz_mu)_b).  tv_sign (optional): the
    pair of caller-fixed sign patterns of the vertical / horizontal differences (as oracle.tf_graph_cpu.restore_gradient)."""
    xt = _t(x, dtype).clone().requires_grad_(True)
    o = forward(P, xt, eps_w, eps_z, None, 0.0, False, dim_c, dtype)
    L = losses(o, xt, dim_c, c_lambda, dtype, l1_sign)
    d = xt - o['xz_mu']
    if tv_sign is None:
        tv = total_variation(d)
    else:
        sv, sh = (_t(s, dtype) for s in tv_sign)
        tv = ((d[:, 1:] - d[:, :-1]) * sv).sum(dim=(1, 2, 3)) + ((d[:, :, 1:] - d[:, :, :-1]) * sh).sum(dim=(1, 2, 3))
    total = (L['loss'] + tv_lambda * tv).sum()
    return torch.autograd.grad(total, xt)[0].detach(), {k: v.detach() for k, v in o.items()}


# --------------------------------------------------------------------------- spatial variant
def init_params_spatial(S, C=1, dim_z=1, dim_w=1, dim_c=9, res=8, seed=1):
    """models/gaussian_mixture_variational_autoencoder_spatial.py:10-67.  The model opens no variable scope; the conv stacks keep this
    repo's canonical Encoder/ | Decoder/ names, the explicitly named 1x1 heads keep the reference's (q_wz_x/..., p_z_wc/...)."""
    from .tf_graph_cpu import AES
    base = ae_init_params(AES, S, C, 128, res, seed)
    rng = np.random.default_rng(seed + 3000)
    ctop = [v.shape[3] for k, v in base.items() if k.startswith('Encoder/enc_conv2D_') and k.endswith('/kernel')][-1]
    P = OrderedDict(base)
    n = dim_z * dim_c
    for name, cin, cout in (('q_wz_x/w_mu', ctop, dim_w), ('q_wz_x/w_log_sigma', ctop, dim_w), ('q_wz_x/z_mu', ctop, dim_z),
                            ('q_wz_x/z_log_sigma', ctop, dim_z), ('p_z_wc/1x1convlayer', dim_w, 64), ('p_z_wc/z_wc_mu', 64, n),
                            ('p_z_wc/z_wc_log_sigma', 64, n)):
        P[name + '/kernel'] = _glorot(rng, (1, 1, cin, cout), cin, cout)
        P[name + '/bias'] = np.zeros(cout, np.float32)
    P['Variable'] = np.full(n, 0.1, np.float32)
    return P


def forward_spatial(P, x, eps_w, eps_z, dim_c=9, dtype=torch.float32):
    """eps_w [B,r,r,dim_w], eps_z [B,r,r,dim_z] (NHWC).  No Dropout in this graph; the DECODER runs on the encoder output itself
    (temp_out is never reassigned between the encoder and the decoder, :22-55): z only enters through the prior terms."""
    P = {k: _t(v, dtype) for k, v in P.items()}
    h = encoder(P, _t(x, dtype).permute(0, 3, 1, 2))
    hn = h.permute(0, 2, 3, 1)                                            # NHWC: a 1x1 conv is a matmul on the channel axis

    def c1(t, name):
        return t @ P[name + '/kernel'][0, 0] + P[name + '/bias']
    o = {}
    o['w_mu'], o['w_log_sigma'] = c1(hn, 'q_wz_x/w_mu'), c1(hn, 'q_wz_x/w_log_sigma')
    o['w_sampled'] = o['w_mu'] + _t(eps_w, dtype) * torch.exp(0.5 * o['w_log_sigma'])
    o['z_mu'], o['z_log_sigma'] = c1(hn, 'q_wz_x/z_mu'), c1(hn, 'q_wz_x/z_log_sigma')
    o['z_sampled'] = o['z_mu'] + _t(eps_z, dtype) * torch.exp(0.5 * o['z_log_sigma'])
    mid = torch.relu(c1(o['w_sampled'], 'p_z_wc/1x1convlayer'))
    B, r, dz = hn.shape[0], hn.shape[1], o['z_mu'].shape[3]
    o['z_wc_mus'] = c1(mid, 'p_z_wc/z_wc_mu').reshape(B, r, r, dz, dim_c)
    o['z_wc_log_sigma_invs'] = (c1(mid, 'p_z_wc/z_wc_log_sigma') + P['Variable']).reshape(B, r, r, dz, dim_c)
    o['xz_mu'] = decoder(P, h).permute(0, 2, 3, 1)
    zs = o['z_sampled'].unsqueeze(-1)
    loglh = -0.5 * ((zs - o['z_wc_mus']) ** 2 * torch.exp(o['z_wc_log_sigma_invs'])) - o['z_wc_log_sigma_invs'] + np.log(np.pi)
    o['pc_logit'] = loglh.sum(3)
    o['pc'] = torch.softmax(o['pc_logit'], dim=-1)
    return o


def losses_spatial(o, x, dim_c=9, c_lambda=1.0, dtype=torch.float32, l1_sign=None):
    """trainers/GMVAE_spatial.py:58-92: the GMVAE terms per spatial position, summed over the positions, batch means."""
    xt = _t(x, dtype)
    L = {}
    diff = o['xz_mu'] - xt
    L['L1'] = diff.abs() if l1_sign is None else diff * _t(l1_sign, dtype)
    L['reconstructionLoss'] = L['mean_p_loss'] = L['L1'].sum(dim=(1, 2, 3)).mean()
    zmu, zlv = o['z_mu'].unsqueeze(-1), o['z_log_sigma'].unsqueeze(-1)
    d_var = (torch.exp(zlv) + (zmu - o['z_wc_mus']) ** 2) * (torch.exp(o['z_wc_log_sigma_invs']) + 1e-6)
    kl = (d_var - (o['z_wc_log_sigma_invs'] + zlv) - 1) * 0.5
    L['conditional_prior_loss'] = torch.matmul(kl, o['pc'].unsqueeze(-1)).squeeze(-1).sum(dim=(1, 2, 3)).mean()
    L['w_prior_loss'] = (0.5 * (o['w_mu'] ** 2 + torch.exp(o['w_log_sigma']) - o['w_log_sigma'] - 1).sum(dim=(1, 2, 3))).mean()
    closs1 = (o['pc'] * torch.log(o['pc'] * dim_c + 1e-8)).sum(3)
    L['c_prior_loss'] = torch.maximum(closs1, torch.full_like(closs1, c_lambda)).sum(dim=(1, 2)).mean()
    L['loss'] = L['mean_p_loss'] + L['conditional_prior_loss'] + L['w_prior_loss'] + L['c_prior_loss']
    return L


def loss_and_grads_spatial(P, x, eps_w, eps_z, dim_c=9, c_lambda=1.0, dtype=torch.float32, l1_sign=None):
    Pt = OrderedDict((k, _t(v, dtype).clone().requires_grad_(True)) for k, v in P.items())
    o = forward_spatial(Pt, x, eps_w, eps_z, dim_c, dtype)
    L = losses_spatial(o, x, dim_c, c_lambda, dtype, l1_sign)
    gs = torch.autograd.grad(L['loss'], list(Pt.values()))
    return ({k: v.detach() for k, v in o.items()}, {k: v.detach() for k, v in L.items()},
            OrderedDict((k, g.detach()) for k, g in zip(Pt, gs)))


def restore_gradient_spatial(P, x, eps_w, eps_z, tv_lambda, dim_c=9, c_lambda=1.0, dtype=torch.float64, l1_sign=None, tv_sign=None):
    xt = _t(x, dtype).clone().requires_grad_(True)
    o = forward_spatial(P, xt, eps_w, eps_z, dim_c, dtype)
    L = losses_spatial(o, xt, dim_c, c_lambda, dtype, l1_sign)
    d = xt - o['xz_mu']
    if tv_sign is None:
        tv = total_variation(d)
    else:
        sv, sh = (_t(s, dtype) for s in tv_sign)
        tv = ((d[:, 1:] - d[:, :-1]) * sv).sum(dim=(1, 2, 3)) + ((d[:, :, 1:] - d[:, :, :-1]) * sh).sum(dim=(1, 2, 3))
    return torch.autograd.grad((L['loss'] + tv_lambda * tv).sum(), xt)[0].detach(), {k: v.detach() for k, v in o.items()}
