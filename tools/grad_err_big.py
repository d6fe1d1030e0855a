"""Developer aid: per-tensor gradient error of one VAE train step against the float64 oracle (in sub-batches) for several
(S, B, math_mode): where does the error of the large-batch step come from?"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from oracle import tf_graph_cpu as O
from test_gpu_step import _noise, _oracle_in_chunks, _relerr
from unsupervised_anomaly_detection_brain_mri_b200.engine import ConvAutoencoderEngine

def run(arch, S, B, mode, rate=0.2):
    P = O.perturb_params(O.init_params(arch, S, seed=1))
    x = O.synthetic_slices(B, S, seed=1234)
    eng = ConvAutoencoderEngine(arch, S, batch=B, math_mode=mode)
    eng.fp.load(P)
    eps, om, em, emc = _noise(arch, B, 128, eng.flat, rate)
    eng.set_inputs(x, None)
    eng.set_noise(eps, em, emc)
    eng.train_step(1e-3, beta1=0.5, dropout_rate=rate, dropout=True, parity_noise=True)
    torch.cuda.synchronize()
    xh_dev = eng.br[0].xhat.cpu().numpy()
    sgn = np.sign(xh_dev.astype(np.float64) - x)
    xh, _, _, L, G = _oracle_in_chunks(arch, P, x, None, eps, om, rate, sgn, None, False)
    grads = eng.fp.to_numpy(eng.fp.grads)
    print(f'--- {arch} S={S} B={B} mode={mode}: xhat err {_relerr(xh_dev, xh):.2e}  sign mismatch {(np.sign(xh - x) != sgn).mean():.2e}')
    for k in P:
        e = _relerr(grads[k], G[k].numpy())
        flag = ' <<<' if e > 1e-4 else ''
        print(f'   {k:45s} {e:.2e}  max|g| {np.abs(G[k].numpy()).max():.3e}{flag}')

for S, B, mode in [(256, 64, 1), (256, 64, 0), (256, 8, 1), (64, 64, 1)]:
    run(O.VAE, S, B, mode)
