"""Generates tests/golden/*.npz from the oracle (TEST INFRASTRUCTURE ONLY; PARITY UNPINNED).

The reference holds no golden vectors and cannot be run here (TensorFlow 1.15), so these fixtures pin the ORACLE
(fp32 torch-CPU restatement, seeds fixed) - they guard the restatement against drift and give the GPU tests a
committed target that does not depend on the torch build of the GPU box.   Run:  python -m oracle.make_golden
"""
import json
import os

import numpy as np
import torch

from . import aae_cpu, anovaegan_cpu, fanogan_cpu, gmvae_cpu, scoring
from . import tf_graph_cpu as O

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def case(arch, S, B, seed=1):
    rate, lr = 0.2, 1e-3
    P = O.perturb_params(O.init_params(arch, S, seed=seed))
    x = O.synthetic_slices(B, S, seed=1234)
    x_ce = x.copy()
    x_ce[:, S // 4:S // 4 + 20, S // 3:S // 3 + 20] = 0
    rng = np.random.default_rng(3)
    eps = rng.standard_normal((B, 128)).astype(np.float32)
    flat = 8 * 8 * (O.stack_plan(S)[1][-1] // 8)
    mk = lambda n: (rng.uniform(size=(B, n)) >= rate).astype(np.float32)  # noqa: E731
    if arch == O.AE:
        masks = {'z': mk(128)}
    else:
        masks = {'mu': mk(128), 'log_sigma': mk(128), 'dec': mk(flat)}
        if arch == O.CEVAE:
            masks.update(mu_ce=mk(128), dec_ce=mk(flat))
    out, L, G = O.loss_and_grads(arch, P, x, x_ce=x_ce, eps=eps, masks=masks, dropout_rate=rate, training=True,
                                 dtype=torch.float32, want_anomaly=(arch == O.CEVAE))
    Pt = {k: torch.from_numpy(v) for k, v in P.items()}
    zeros = {k: torch.zeros_like(v) for k, v in Pt.items()}
    Pn, _, _ = O.adam_tf(Pt, G, zeros, zeros, 1, lr, 0.5)
    d = {'x': x, 'x_ce': x_ce, 'eps': eps, 'x_hat': out['x_hat'].numpy(), 'lr': np.float32(lr), 'rate': np.float32(rate)}
    for k, v in masks.items():
        d['mask_' + k] = v
    for k, v in L.items():
        if v.ndim == 0:
            d['loss_' + k] = np.float64(v)
    if arch == O.CEVAE:
        d['x_hat_ce'] = out['x_hat_ce'].numpy()
        d['anomaly'] = L['anomaly'].numpy()
    names = list(P.keys())
    d['param_names'] = np.array(names)
    d['param_sum'] = np.array([float(P[k].astype(np.float64).sum()) for k in names])
    d['grad_l2'] = np.array([float(G[k].double().norm()) for k in names])
    d['grad_sum'] = np.array([float(G[k].double().sum()) for k in names])
    for k in names:                      # small tensors in full: biases, gammas, betas
        if G[k].numel() <= 1024:
            d['grad|' + k.replace('/', '|')] = G[k].numpy()
            d['new|' + k.replace('/', '|')] = Pn[k].numpy()
    return d


def scoring_case():
    rng = np.random.default_rng(77)
    N, S = 6, 32
    x = O.synthetic_slices(N, S, seed=5)[..., 0]
    xr = np.clip(x + 0.15 * rng.standard_normal(x.shape), 0, 1).astype(np.float32)
    mask = np.stack([scoring.erode_brainmask(x[i] > 0, 2) for i in range(N)])
    prior = float(np.quantile(x, 0.9))
    sub = scoring.residual(x, xr, mask, prior, True, False)
    labels = (rng.uniform(size=x.shape) < 0.08).astype(np.uint8) * (x > 0)
    best, thr, ths, scs = scoring.best_dice_search(sub, labels, granularity=4)
    return {'x': x, 'x_rec': xr, 'mask': mask.astype(np.uint8), 'prior': np.float64(prior), 'diff': sub.astype(np.float32),
            'labels': labels.astype(np.uint8), 'best_dice': np.float64(best), 'best_thr': np.float64(thr),
            'threshs': np.array(ths, np.float64), 'scores': np.array(scs, np.float64)}


def sibling_inputs(S=32, B=2, zDim=128):
    """Seeded feeds shared by make_golden and tests/test_oracle.py (so the fixtures hold outputs only)."""
    rng = np.random.default_rng(11)
    x = O.synthetic_slices(B, S, seed=1234)
    flat = 8 * 8 * (O.stack_plan(S)[1][-1] // 8)
    mk = lambda n: (rng.uniform(size=(B, n)) >= 0.2).astype(np.float32)  # noqa: E731
    return dict(x=x, eps=rng.standard_normal((B, zDim)).astype(np.float32), z=rng.standard_normal((B, zDim)).astype(np.float32),
                alpha=rng.uniform(size=(B, 1)).astype(np.float32), eps_w=rng.standard_normal((B, 1)).astype(np.float32),
                masks_z=dict(mu=mk(zDim), ls=mk(zDim), dec=mk(flat), z=mk(zDim), w_mu=mk(1), w_ls=mk(1), z_mu=mk(zDim)),
                eps_w_sp=rng.standard_normal((B, 8, 8, 1)).astype(np.float32),
                eps_z_sp=rng.standard_normal((B, 8, 8, 1)).astype(np.float32))


def _scalars(o):
    return {k: float(v) for k, v in o.items() if torch.is_tensor(v) and v.ndim == 0}


def sibling_outputs(S=32, B=2):
    """Loss scalars and per-variable gradient L2 norms of every sibling oracle on sibling_inputs() (fp32, dropout 0.2)."""
    f = sibling_inputs(S, B)
    m = f['masks_z']
    out = {}
    P = fanogan_cpu.perturb(anovaegan_cpu.init_params(S, seed=1))
    for op in ('vae', 'gen', 'disc'):
        o, G = anovaegan_cpu.Trainer(P, lr=1e-3, dropout_rate=0.2).step(op, f['x'], f['eps'], f['alpha'],
                                                                          masks=dict(mu=m['mu'], ls=m['ls'], dec=m['dec']))
        out['anovaegan_' + op] = dict(scalars=_scalars(o), grad_l2={k: float(g.double().norm()) for k, g in G.items()})
    for constrained in (False, True):
        P = aae_cpu.perturb(aae_cpu.init_params(S, seed=1, constrained=constrained))
        for op in ('ae', 'disc', 'gen'):
            o, G = aae_cpu.Trainer(P, lr=1e-3, dropout_rate=0.2, constrained=constrained, rho=0.7).step(
                op, f['x'], f['z'], f['alpha'], masks=dict(z=m['z'], dec=m['dec']))
            out[('caae_' if constrained else 'aae_') + op] = dict(scalars=_scalars(o),
                                                                   grad_l2={k: float(g.double().norm()) for k, g in G.items()})
    P = gmvae_cpu.perturb(gmvae_cpu.init_params(S, seed=1))
    o, L, G = gmvae_cpu.loss_and_grads(P, f['x'], f['eps_w'], f['eps'], masks=dict(w_mu=m['w_mu'], w_ls=m['w_ls'], z_mu=m['z_mu'],
                                                                                    dec=m['dec']), dropout_rate=0.2, c_lambda=0.5)
    out['gmvae'] = dict(scalars=_scalars(L), grad_l2={k: float(g.double().norm()) for k, g in G.items()})
    P = gmvae_cpu.perturb(gmvae_cpu.init_params_spatial(S, seed=1))
    o, L, G = gmvae_cpu.loss_and_grads_spatial(P, f['x'], f['eps_w_sp'], f['eps_z_sp'], c_lambda=0.5)
    out['gmvae_spatial'] = dict(scalars=_scalars(L), grad_l2={k: float(g.double().norm()) for k, g in G.items()})
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, 'ae_32_b2.npz'), **case(O.AE, 32, 2))
    np.savez_compressed(os.path.join(OUT, 'vae_32_b2.npz'), **case(O.VAE, 32, 2))
    np.savez_compressed(os.path.join(OUT, 'cevae_32_b2.npz'), **case(O.CEVAE, 32, 2))
    np.savez_compressed(os.path.join(OUT, 'scoring.npz'), **scoring_case())
    with open(os.path.join(OUT, 'siblings_32_b2.json'), 'w') as fh:
        json.dump(sibling_outputs(), fh, indent=1, sort_keys=True)
    print('wrote', sorted(os.listdir(OUT)))


if __name__ == '__main__':
    main()
