"""Developer aid: per-tensor gradient error of the engine (SIMT vs tcgen05 paths) against the float64 oracle."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from oracle import tf_graph_cpu as O
from unsupervised_anomaly_detection_brain_mri_b200.engine import ConvAutoencoderEngine

def rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-30))

for arch, S, B in [(O.AE, 128, 16), (O.VAE, 64, 4), (O.VAE, 256, 2)]:
    P = O.perturb_params(O.init_params(arch, S, seed=1))
    x = O.synthetic_slices(B, S, seed=1234)
    eps = np.random.default_rng(3).standard_normal((B, 128)).astype(np.float32)
    out, L, G = O.loss_and_grads(arch, P, x, eps=eps, training=False, dtype=torch.float64)
    out32, L32, G32 = O.loss_and_grads(arch, P, x, eps=eps, training=False, dtype=torch.float32)
    res = {}
    for mode in (0, 1):
        eng = ConvAutoencoderEngine(arch, S, batch=B, math_mode=mode)
        eng.fp.load(P); eng.set_inputs(x); eng.set_noise(eps)
        eng.train_step(1e-3, dropout_rate=0.0, dropout=False, parity_noise=True)
        torch.cuda.synchronize()
        g = eng.fp.to_numpy(eng.fp.grads)
        res[mode] = {k: rel(g[k], G[k].numpy()) for k in P}
        print(arch, S, B, 'mode', mode, 'xhat rel', rel(eng.br[0].xhat.cpu().numpy(), out['x_hat'].numpy()),
              'loss rel', abs(eng.losses()['loss'] - float(L['loss'])) / float(L['loss']))
    cpu32 = {k: rel(G32[k].numpy(), G[k].numpy()) for k in P}
    worst = sorted(P, key=lambda k: -res[1][k])[:8]
    for k in worst:
        print(f'   {k:42s} simt {res[0][k]:.2e}  tc {res[1][k]:.2e}  torch-cpu-fp32 {cpu32[k]:.2e}')
