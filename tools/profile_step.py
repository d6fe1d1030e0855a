"""Eager (no CUDA graph) VAE-256 B=64 train steps for ncu: `ncu ... python tools/profile_step.py [steps] [math: simt | tc3 | tc1] [batch]`."""
import sys
import torch
sys.path.insert(0, '.')
from unsupervised_anomaly_detection_brain_mri_b200 import abi
from unsupervised_anomaly_detection_brain_mri_b200.engine import ConvAutoencoderEngine
from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import make_volume

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mode = {'simt': 0, 'tc3': 1, 'tc1': 2}[sys.argv[2]] if len(sys.argv) > 2 else 1
B, S = int(sys.argv[3]) if len(sys.argv) > 3 else 64, 256
eng = ConvAutoencoderEngine('variational_autoencoder', S, batch=B, math_mode=mode)
x = make_volume(S, B, seed=1000, lesions=False)[0][..., None]
eng.set_inputs(x)
for i in range(steps):
    eng.train_step(1e-4, dropout_rate=0.2, dropout=True, use_graph=False)
torch.cuda.synchronize()
print('loss', eng.losses()['loss'], 'launches', abi.lib().uad_launch_count())
