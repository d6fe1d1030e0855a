"""GPU (first green hardware run: round 2, gpurun call r2d): the context-encoder trainer's step (reconstruction target decoupled from the input, engine.set_target)
with the real kernels; verified on CPU through the ABI emulator (tests/test_engine_emulated.py), not yet run on hardware."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tf_graph_cpu as O  # noqa: E402


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))


@pytest.mark.parametrize('mode', [0, 1])
@pytest.mark.parametrize('use_graph', [False, True])
def test_context_encoder_step(mode, use_graph):
    from unsupervised_anomaly_detection_brain_mri_b200.engine import AE, ConvAutoencoderEngine
    S, B, rate, lr = 64, 4, 0.2, 1e-3
    P = O.perturb_params(O.init_params(O.AE, S, seed=1))
    x = O.synthetic_slices(B, S, seed=31)
    x_ce = x.copy()
    x_ce[:, 20:40, 22:42] = 0
    eng = ConvAutoencoderEngine(AE, S, batch=B, math_mode=mode)
    eng.fp.load(P)
    mz = (np.random.default_rng(8).uniform(size=(B, 128)) >= rate).astype(np.float32)
    eng.set_inputs(x_ce)
    eng.set_target(x)
    eng.set_noise(None, {'mu': mz})
    eng._keep = 1.0 / (1.0 - rate)
    eng.forward(training=True, dropout_rate=rate)
    l1_sign = np.sign(eng.br[0].xhat.cpu().numpy() - x)
    if use_graph:                                  # perf-mode path: eager warm-up, capture, replay - the target pointer is captured
        for _ in range(3):
            eng.fp.load(P)
            eng.train_step(lr, beta1=0.5, dropout_rate=0.0, dropout=False, use_graph=True)
        rate, mz = 0.0, None
        l1_sign = None
    else:
        eng.train_step(lr, beta1=0.5, dropout_rate=rate, dropout=True, parity_noise=True)
    torch.cuda.synchronize()
    Pt = OrderedDict((k, torch.from_numpy(v).double().requires_grad_(True)) for k, v in P.items())
    out = O.forward(O.AE, Pt, x_ce, masks={'z': mz} if mz is not None else None, dropout_rate=rate, training=True, dtype=torch.float64)
    L = O.losses(O.AE, out, x, dtype=torch.float64, l1_sign=l1_sign)
    assert abs(eng.losses()['loss'] - float(L['loss'].detach())) < 1e-4 * float(L['loss'].detach())
    if not use_graph:
        G = torch.autograd.grad(L['loss'], list(Pt.values()))
        grads = eng.fp.to_numpy(eng.fp.grads)
        for k, g in zip(Pt, G):
            assert _rel(grads[k], g.numpy()) < 5e-4, k
