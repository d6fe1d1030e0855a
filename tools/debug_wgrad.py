"""Developer aid for the tcgen05 wgrad kernel: UAD_WGRAD_DEBUG=0/1/2 variants on a tiny case."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from unsupervised_anomaly_detection_brain_mri_b200 import abi
from unsupervised_anomaly_detection_brain_mri_b200.abi import call
DEV = 'cuda:0'
L = abi.lib()

def wgrad(x, dz, B, H, Cin, Cout, mode):
    wsb = L.uad_conv_workspace_bytes(2, B, H, H, Cin, Cout, 5, mode)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=DEV)
    out = torch.full((5, 5, Cin, Cout), float('nan'), device=DEV)
    xd, dd = torch.from_numpy(x).to(DEV), torch.from_numpy(dz).to(DEV)
    call('uad_conv2d_wgrad', xd.data_ptr(), dd.data_ptr(), out.data_ptr(), B, H, H, Cin, Cout, 5, 0, mode, ws.data_ptr(), wsb,
         torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out.cpu().numpy()

rng = np.random.default_rng(0)
for (B, H, Cin, Cout) in [(2, 16, 32, 32), (2, 16, 32, 64), (1, 32, 64, 128)]:
    x = rng.standard_normal((B, H, H, Cin)).astype(np.float32)
    dz = rng.standard_normal((B, H // 2, H // 2, Cout)).astype(np.float32)
    ref = wgrad(x, dz, B, H, Cin, Cout, 0)
    for dbg in ('0', '1', '2'):
        os.environ['UAD_WGRAD_DEBUG'] = dbg
        got = wgrad(x, dz, B, H, Cin, Cout, 1)
        colsum = dz.reshape(-1, Cout).sum(0)
        print(f'B={B} H={H} Cin={Cin} Cout={Cout} dbg={dbg}: rel={np.abs(got-ref).max()/np.abs(ref).max():.3e} '
              f'nan={int(np.isnan(got).sum())} got[0,0,0,:4]={got[0,0,0,:4]} got[2,2,5,:4]={got[2,2,5,:4]} got[4,4,31,-4:]={got[4,4,-1,-4:]}')
        if dbg == '0':
            print('      ref[0,0,0,:4]=', ref[0, 0, 0, :4], 'ref[2,2,5,:4]=', ref[2, 2, 5, :4])
        if dbg == '1':
            print('      colsum(dz)[:4]=', colsum[:4], '(tap (2,2) = centre sees every pixel)')
    os.environ['UAD_WGRAD_DEBUG'] = '0'
