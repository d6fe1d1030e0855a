"""Developer aid: forward conv blocks writing both outputs (z and a) vs the activated output only (VAE-256 shapes, B=64)."""
import sys
import torch
sys.path.insert(0, '.')
from unsupervised_anomaly_detection_brain_mri_b200 import abi
from unsupervised_anomaly_detection_brain_mri_b200.abi import call
L = abi.lib()
DEV = 'cuda:0'
B = 64


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


cases = [('enc0', 'conv', 256, 1, 32), ('enc1', 'conv', 128, 32, 64), ('enc2', 'conv', 64, 64, 128), ('enc3', 'conv', 32, 128, 128),
         ('dec1', 'convT', 16, 128, 64), ('dec2', 'convT', 32, 64, 32), ('dec3', 'convT', 64, 32, 32), ('dec4', 'convT', 128, 32, 32)]
tot = [0.0, 0.0]
for name, kind, H, Cin, Cout in cases:
    opid = 0 if kind == 'conv' else 3
    wsb = L.uad_conv_workspace_bytes(opid, B, H, H, Cin, Cout, 5, 1)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    x = torch.randn(B, H, H, Cin, device=DEV)
    Ho = H // 2 if kind == 'conv' else 2 * H
    a = torch.empty(B, Ho, Ho, Cout, device=DEV)
    z = torch.empty_like(a)
    w = torch.randn(5, 5, Cin, Cout, device=DEV) * 0.05 if kind == 'conv' else torch.randn(5, 5, Cout, Cin, device=DEV) * 0.05
    g = torch.ones(Cout, device=DEV)
    bt = torch.zeros(Cout, device=DEV)
    fn = 'uad_conv2d_fwd' if kind == 'conv' else 'uad_convT2d_fwd'
    st = torch.cuda.current_stream().cuda_stream
    res = []
    for zp in (z.data_ptr(), None):
        t = timeit(lambda: call(fn, x.data_ptr(), w.data_ptr(), bt.data_ptr(), g.data_ptr(), bt.data_ptr(), zp, a.data_ptr(), B, H, H, Cin, Cout,
                                5, 1, 0.3, 0.9995, 1, ws.data_ptr(), wsb, st))
        res.append(t)
    tot[0] += res[0]
    tot[1] += res[1]
    print(f'{name}: z+a {res[0]:.3f} ms   a only {res[1]:.3f} ms', flush=True)
print(f'total: z+a {tot[0]:.3f} ms   a only {tot[1]:.3f} ms')
