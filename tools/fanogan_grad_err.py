"""Per-variable gradient error of the f-AnoGAN train ops vs the float64 oracle (developer aid)."""
import sys

import torch

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from oracle import fanogan_cpu as FO
from test_gpu_fanogan import _feed, _rel, _signs
from unsupervised_anomaly_detection_brain_mri_b200.fanogan_engine import FanoganEngine

which, S, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
for mode in (0, 1):
    P = FO.perturb(FO.init_params(S, seed=1))
    eng = FanoganEngine(S, batch=B, math_mode=mode)
    x, z, alpha, m_enc, m_gen = _feed(S, B, 0.2, eng.flat)
    eng.fp.load(P)
    eng.enable_training()
    eng.set_inputs(x)
    eng.set_latent(z)
    eng.alpha.copy_(torch.from_numpy(alpha.reshape(-1)))
    eng.mask_enc.copy_(torch.from_numpy(m_enc))
    eng.mask_gen.copy_(torch.from_numpy(m_gen))
    tr = FO.WganTrainer(P, lr=1e-3, dropout_rate=0.2, dtype=torch.float64)
    out, G = tr.step(which, x, z, alpha, mask_enc=m_enc, mask_gen=m_gen, signs=_signs(eng, which, 0.2) if len(sys.argv) > 4 else None)
    {'gen': eng.step_gen, 'disc': eng.step_disc, 'enc': eng.step_enc}[which](1e-3, dropout_rate=0.2, dropout=True, parity_noise=True,
                                                                            apply=False)
    got = eng.fp.to_numpy(eng.fp.grads)
    print('mode', mode)
    for k, v in G.items():
        print(f'  {k:50s} max|g| {float(v.abs().max()):.3e}  relerr {_rel(got[k], v.numpy()):.2e}')
