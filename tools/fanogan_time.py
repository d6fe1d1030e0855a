"""Device time of the three f-AnoGAN train ops at full size (developer aid; CUDA events, resident inputs).
usage: python tools/fanogan_time.py [S] [B] [math_mode] [iters] [graph 0|1]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from oracle.tf_graph_cpu import synthetic_slices
from unsupervised_anomaly_detection_brain_mri_b200 import abi
from unsupervised_anomaly_detection_brain_mri_b200.fanogan_engine import FanoganEngine

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
graph = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
eng = FanoganEngine(S, batch=B, math_mode=mode)
eng.enable_training()
eng.set_inputs(synthetic_slices(B, S, seed=3))
eng.set_latent(np.random.default_rng(0).standard_normal((B, 128)).astype(np.float32))
res = {'S': S, 'B': B, 'math_mode': mode, 'cuda_graph': graph}
for name, fn in (('gen', eng.step_gen), ('disc', eng.step_disc), ('enc', eng.step_enc)):
    for _ in range(3):
        fn(1e-4, dropout_rate=0.2, use_graph=graph)
    torch.cuda.synchronize()
    l0 = abi.lib().uad_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn(1e-4, dropout_rate=0.2, use_graph=graph)
    e1.record()
    torch.cuda.synchronize()
    res[name + '_ms'] = e0.elapsed_time(e1) / iters
    res[name + '_launches'] = (abi.lib().uad_launch_count() - l0) // iters
    res[name + '_out'] = {k: round(v, 6) for k, v in out.items()}
res['wgan_batch_ms'] = res['gen_ms'] + 5 * res['disc_ms']
res['wgan_slices_per_s'] = B / res['wgan_batch_ms'] * 1e3
res['enc_slices_per_s'] = B / res['enc_ms'] * 1e3
res['mem_GB'] = torch.cuda.max_memory_allocated() / 2 ** 30
print(json.dumps(res))
