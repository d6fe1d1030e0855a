"""Developer aid: event-times conv_halo_ss (Form F / Form T at the VAE-256 layer shapes) under the UAD_HS_DEBUG switches."""
import os, sys
import torch
sys.path.insert(0, '.')
from unsupervised_anomaly_detection_brain_mri_b200 import abi
from unsupervised_anomaly_detection_brain_mri_b200.abi import call
DEV = 'cuda:0'
L = abi.lib()
st = lambda: torch.cuda.current_stream().cuda_stream

REPS, WARM = int(os.environ.get('REPS', 10)), int(os.environ.get('WARM', 3))

def timeit(fn, reps=REPS):
    for _ in range(WARM): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

B = int(os.environ.get('B', 64))
MATH = int(os.environ.get('MATH', 1))   # 1 = 3xTF32, 2 = 1xTF32
cases = {
  'T4.fwd   FormT N=32  C=32  128^2->256^2': ('convT_fwd', 128, 32, 32),
  'T4.dgrad FormF N=32  C=32  256^2->128^2': ('convT_dgrad', 128, 32, 32),
  'T3.fwd   FormT N=32  C=32  64^2->128^2': ('convT_fwd', 64, 32, 32),
  'enc1.fwd FormF N=64  C=32  128^2->64^2': ('conv_fwd', 128, 32, 64),
  'enc1.dgr FormT N=32  C=64  64^2->128^2': ('conv_dgrad', 128, 32, 64),
  'enc2.fwd FormF N=128 C=64  64^2->32^2': ('conv_fwd', 64, 64, 128),
  'enc2.dgr FormT N=64  C=128 32^2->64^2': ('conv_dgrad', 64, 64, 128),
  'enc3.fwd FormF N=128 C=128 32^2->16^2': ('conv_fwd', 32, 128, 128),
  'enc3.dgr FormT N=128 C=128 16^2->32^2': ('conv_dgrad', 32, 128, 128),
}
modes = [int(m) for m in os.environ.get('MODES', '0,1,2,4,8,16,24,26,27,31').split(',')]
print('modes (bits): 1 no lo pass, 2 no MMAs, 4 no global stores, 8 no weight loads, 16 no halo loads')
CASES = os.environ.get('CASES')
if CASES:
    keys = list(cases); cases = {keys[int(i)]: cases[keys[int(i)]] for i in CASES.split(',')}
for name, (op, H, Cin, Cout) in cases.items():
    opid = {'conv_fwd': 0, 'conv_dgrad': 1, 'conv_wgrad': 2, 'convT_fwd': 3, 'convT_dgrad': 4, 'convT_wgrad': 5}[op]
    wsb = L.uad_conv_workspace_bytes(opid, B, H, H, Cin, Cout, 5, 1)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    if op.startswith('convT'):
        x = torch.randn(B, H, H, Cin, device=DEV); y = torch.randn(B, 2 * H, 2 * H, Cout, device=DEV); w = torch.randn(5, 5, Cout, Cin, device=DEV) * 0.05
    else:
        x = torch.randn(B, H, H, Cin, device=DEV); y = torch.randn(B, H // 2, H // 2, Cout, device=DEV); w = torch.randn(5, 5, Cin, Cout, device=DEV) * 0.05
    dx = torch.empty_like(x)
    def run():
        if op == 'conv_fwd': call('uad_conv2d_fwd', x.data_ptr(), w.data_ptr(), None, None, None, None, y.data_ptr(), B, H, H, Cin, Cout, 5, 1, 0.3, 1.0, MATH, ws.data_ptr(), wsb, st())
        elif op == 'conv_dgrad': call('uad_conv2d_dgrad', y.data_ptr(), w.data_ptr(), dx.data_ptr(), B, H, H, Cin, Cout, 5, MATH, ws.data_ptr(), wsb, st())
        elif op == 'convT_fwd': call('uad_convT2d_fwd', x.data_ptr(), w.data_ptr(), None, None, None, None, y.data_ptr(), B, H, H, Cin, Cout, 5, 1, 0.3, 1.0, MATH, ws.data_ptr(), wsb, st())
        elif op == 'convT_dgrad': call('uad_convT2d_dgrad', y.data_ptr(), w.data_ptr(), dx.data_ptr(), B, H, H, Cin, Cout, 5, MATH, ws.data_ptr(), wsb, st())
    res = []
    for dbg in modes:
        os.environ['UAD_HS_DEBUG'] = str(dbg)
        res.append(f'{dbg}:{timeit(run):.3f}')
    os.environ['UAD_HS_DEBUG'] = '0'
    print(f'{name:42s} ' + ' '.join(res), flush=True)
