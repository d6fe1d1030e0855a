"""Developer aid: smoke()'s exact set-up with per-tensor gradient errors (bisect with UAD_HS=0 / UAD_WGRAD_SS=0)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, '.')
from oracle import tf_graph_cpu as O
from unsupervised_anomaly_detection_brain_mri_b200.engine import ConvAutoencoderEngine
arch, S, B, lr = O.VAE, 64, 4, 1e-3
P = O.perturb_params(O.init_params(arch, S, seed=1))
x = O.synthetic_slices(B, S, seed=1234)
for seed in (int(v) for v in os.environ.get('SEEDS', '2').split(',')):
  eps = np.random.default_rng(seed).standard_normal((B, 128)).astype(np.float32)
  for mode in (0, 1):
      eng = ConvAutoencoderEngine(arch, S, batch=B, device='cuda:0', math_mode=mode)
      eng.fp.load(P); eng.set_inputs(x); eng.set_noise(eps)
      eng.train_step(lr, dropout_rate=0.0, dropout=False, parity_noise=True)
      torch.cuda.synchronize()
      sgn = np.sign(eng.br[0].xhat.cpu().numpy().astype(np.float64) - x)
      out, L, G = O.loss_and_grads(arch, P, x, eps=eps, training=True, dtype=torch.float64, l1_sign=sgn)
      g = eng.fp.to_numpy(eng.fp.grads)
      errs = {k: float(np.abs(g[k] - G[k].numpy()).max() / max(np.abs(G[k].numpy()).max(), 1e-30)) for k in P}
      print('seed', seed, 'mode', mode, 'HS', os.environ.get('UAD_HS', '1'), 'WS', os.environ.get('UAD_WGRAD_SS', '1'), 'sign mismatch', float((np.sign(out['x_hat'].numpy() - x) != sgn).mean()))
      for k in [k for k in P if errs[k] > 2e-5]:
          print(f'    {k:42s} {errs[k]:.3e}')
