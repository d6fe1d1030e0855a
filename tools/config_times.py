"""Device-resident train-step rate of the other BASELINE.json configs that fit one GPU (CUDA events, CUDA-graph replay,
resident inputs; parity for these configs is covered by tests/test_gpu_step.py):
  C1 dense AE 128x128 B=16 | C2 VAE 256x256 B=64 (bench.py's headline) | C3 ceVAE 256x256 B=128 (masked-patch branch + anomaly)
usage: python tools/config_times.py [steps]"""
import json
import sys

import torch

sys.path.insert(0, '.')
from unsupervised_anomaly_detection_brain_mri_b200 import abi
from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import make_volume
from unsupervised_anomaly_detection_brain_mri_b200.engine import ConvAutoencoderEngine

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
out = []
for name, arch, S, B in (('C1 dense AE 128x128 B=16', 'autoencoder', 128, 16), ('C2 VAE 256x256 B=64', 'variational_autoencoder', 256, 64),
                         ('C3 ceVAE 256x256 B=128', 'context_encoder_variational_autoencoder', 256, 128),
                         ('spatial AE 256x256 B=64', 'autoencoder_spatial', 256, 64)):
    eng = ConvAutoencoderEngine(arch, S, batch=B, math_mode=abi.MATH_TC_3XTF32)
    x = make_volume(S, B, seed=1000, lesions=False)[0][..., None]
    x_ce = x.copy()
    x_ce[:, S // 3:S // 3 + 20, S // 3:S // 3 + 20] = 0
    ce = arch.startswith('context')
    eng.set_inputs(x, x_ce if ce else None)
    for _ in range(4):
        eng.train_step(1e-4, dropout_rate=0.2, dropout=True, want_anomaly=ce, use_graph=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.train_step(1e-4, dropout_rate=0.2, dropout=True, want_anomaly=ce, use_graph=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out.append({'config': name, 'ms_per_step': ms, 'slices_per_s': B / ms * 1e3, 'loss': eng.losses()['loss'],
                'mem_GB': torch.cuda.max_memory_allocated() / 2 ** 30})
    del eng
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
