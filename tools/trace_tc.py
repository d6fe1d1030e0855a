"""Developer aid: clock64 trace of one steady-state CTA of the N=32 tcgen05 kernel (T4.fwd / T4.dgrad shapes)."""
import ctypes, os, sys
import torch
sys.path.insert(0, '.')
from unsupervised_anomaly_detection_brain_mri_b200 import abi
from unsupervised_anomaly_detection_brain_mri_b200.abi import call
L = abi.lib(); DEV = 'cuda:0'; B = 64
st = lambda: torch.cuda.current_stream().cuda_stream
os.environ['UAD_TC_DEBUG'] = '16'
for name, op, H, Cin, Cout in [('T4.fwd', 'fwd', 128, 32, 32), ('T4.dgrad', 'dgrad', 128, 32, 32)]:
    wsb = L.uad_conv_workspace_bytes(3 if op == 'fwd' else 4, B, H, H, Cin, Cout, 5, 1)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    x = torch.randn(B, H, H, Cin, device=DEV); y = torch.randn(B, 2 * H, 2 * H, Cout, device=DEV); w = torch.randn(5, 5, Cout, Cin, device=DEV) * 0.05
    z = torch.empty_like(y); dx = torch.empty_like(x)
    for _ in range(3):
        if op == 'fwd': call('uad_convT2d_fwd', x.data_ptr(), w.data_ptr(), None, None, None, z.data_ptr(), y.data_ptr(), B, H, H, Cin, Cout, 5, 1, 0.3, 1.0, 1, ws.data_ptr(), wsb, st())
        else: call('uad_convT2d_dgrad', y.data_ptr(), w.data_ptr(), dx.data_ptr(), B, H, H, Cin, Cout, 5, 1, ws.data_ptr(), wsb, st())
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 64)()
    L.uad_debug_trace(buf)
    n = buf[63]
    t = [buf[i] - buf[0] for i in range(n)]
    print(name, 'stamps (cycles from entry):', t)
    print('   deltas:', [t[i + 1] - t[i] for i in range(n - 1)])
print('stamps: entry, prologue done, then per class: [conversions issued, accumulators complete, TMEM read done, stores issued], final: after teardown sync')
