"""Developer aid: event-times the filter-gradient launches of the VAE-256 layer shapes (wgrad_ss) under the current UAD_WS_* switches."""
import os, sys
import torch
sys.path.insert(0, '.')
from unsupervised_anomaly_detection_brain_mri_b200 import abi
from unsupervised_anomaly_detection_brain_mri_b200.abi import call
DEV = 'cuda:0'
L = abi.lib()
st = lambda: torch.cuda.current_stream().cuda_stream
B = int(os.environ.get('B', 64)); MATH = int(os.environ.get('MATH', 1))
# (name, op, H of the conv input / convT input, Cin, Cout)
cases = [('enc1', 2, 128, 32, 64), ('enc2', 2, 64, 64, 128), ('enc3', 2, 32, 128, 128), ('enc4', 2, 16, 128, 128),
         ('dec0', 5, 8, 128, 128), ('dec1', 5, 16, 128, 128), ('dec2', 5, 32, 128, 64), ('dec3', 5, 64, 64, 32), ('dec4', 5, 128, 32, 32)]
tot = 0.0
out = []
for name, op, H, Cin, Cout in cases:
    wsb = L.uad_conv_workspace_bytes(op, B, H, H, Cin, Cout, 5, MATH)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    if op == 2:
        x = torch.randn(B, H, H, Cin, device=DEV); dz = torch.randn(B, H // 2, H // 2, Cout, device=DEV); dw = torch.empty(5, 5, Cin, Cout, device=DEV)
        fn = lambda: call('uad_conv2d_wgrad', x.data_ptr(), dz.data_ptr(), dw.data_ptr(), B, H, H, Cin, Cout, 5, 0, MATH, ws.data_ptr(), wsb, st())
    else:
        x = torch.randn(B, H, H, Cin, device=DEV); dz = torch.randn(B, 2 * H, 2 * H, Cout, device=DEV); dw = torch.empty(5, 5, Cout, Cin, device=DEV)
        fn = lambda: call('uad_convT2d_wgrad', x.data_ptr(), dz.data_ptr(), dw.data_ptr(), B, H, H, Cin, Cout, 5, 0, MATH, ws.data_ptr(), wsb, st())
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10
    tot += t
    out.append(f'{name}:{t:.3f}')
print(f"WAVES={os.environ.get('UAD_WS_WAVES','4')} DEPTH={os.environ.get('UAD_WS_DEPTH','4096')}  " + ' '.join(out) + f'  total {tot:.3f}')
