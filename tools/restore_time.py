"""Device time of the VAE_You restoration loop at full size (developer aid; CUDA events).
usage: python tools/restore_time.py [S] [N slices] [steps] [graph 0|1]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from unsupervised_anomaly_detection_brain_mri_b200 import abi
from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import make_volume
from unsupervised_anomaly_detection_brain_mri_b200.engine import ConvAutoencoderEngine

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 110
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 150
graph = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
eng = ConvAutoencoderEngine('variational_autoencoder', S, batch=N, math_mode=1)
x = make_volume(S, N, seed=1000, lesions=True)[0][..., None]
eng.set_inputs(x)
eng.restore(3, 1e-3, 1.0, use_graph=graph)            # warm-up (+ capture)
torch.cuda.synchronize()
eng.set_inputs(x)
l0 = abi.lib().uad_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.restore(steps, 1e-3, 1.0, use_graph=graph)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
out = eng.br[0].x.cpu().numpy()
print(json.dumps({'S': S, 'slices': N, 'restore_steps': steps, 'cuda_graph': graph, 'ms_total': ms, 'ms_per_iteration': ms / steps,
                  'slice_iterations_per_s': N * steps / ms * 1e3, 'volumes_per_s': 1e3 / ms, 'finite': bool(np.isfinite(out).all()),
                  'mean_abs_change': float(np.abs(out - x).mean()), 'mem_GB': torch.cuda.max_memory_allocated() / 2 ** 30}))
