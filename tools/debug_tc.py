"""Developer aid: compares the tcgen05 conv path against the SIMT path and the float64 oracle on one shape, with
error localisation (which rows / channels / k-slices are wrong) to diagnose descriptor / layout mistakes."""
import math
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from unsupervised_anomaly_detection_brain_mri_b200 import abi  # noqa: E402
from unsupervised_anomaly_detection_brain_mri_b200.abi import call  # noqa: E402

DEV = 'cuda:0'


def run(op, B, H, Cin, Cout, mode, x, w, b=None):
    L = abi.lib()
    wsb = L.uad_conv_workspace_bytes({'conv': 0, 'convT': 3, 'conv_dgrad': 1, 'convT_dgrad': 4}[op], B, H, H, Cin, Cout, 5, mode)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    xd, wd = torch.from_numpy(x).to(DEV), torch.from_numpy(w).to(DEV)
    bd = None if b is None else torch.from_numpy(b).to(DEV)
    if op == 'conv':
        z = torch.full((B, H // 2, H // 2, Cout), float('nan'), device=DEV)
        call('uad_conv2d_fwd', xd.data_ptr(), wd.data_ptr(), bd.data_ptr() if bd is not None else None, None, None, z.data_ptr(), None,
             B, H, H, Cin, Cout, 5, 0, 0.0, 1.0, mode, ws.data_ptr(), wsb, st)
    elif op == 'convT':
        z = torch.full((B, 2 * H, 2 * H, Cout), float('nan'), device=DEV)
        call('uad_convT2d_fwd', xd.data_ptr(), wd.data_ptr(), bd.data_ptr() if bd is not None else None, None, None, z.data_ptr(), None,
             B, H, H, Cin, Cout, 5, 0, 0.0, 1.0, mode, ws.data_ptr(), wsb, st)
    elif op == 'conv_dgrad':        # x is dz [B,H/2,H/2,Cout]
        z = torch.full((B, H, H, Cin), float('nan'), device=DEV)
        call('uad_conv2d_dgrad', xd.data_ptr(), wd.data_ptr(), z.data_ptr(), B, H, H, Cin, Cout, 5, mode, ws.data_ptr(), wsb, st)
    else:                           # convT_dgrad: x is dz [B,2H,2H,Cout]
        z = torch.full((B, H, H, Cin), float('nan'), device=DEV)
        call('uad_convT2d_dgrad', xd.data_ptr(), wd.data_ptr(), z.data_ptr(), B, H, H, Cin, Cout, 5, mode, ws.data_ptr(), wsb, st)
    torch.cuda.synchronize()
    return z.cpu().numpy()


def main():
    rng = np.random.default_rng(0)
    cases = [('conv', 2, 32, 32, 32), ('conv', 2, 16, 64, 128), ('conv', 3, 16, 128, 64), ('convT', 2, 16, 32, 32),
             ('convT', 2, 8, 128, 128), ('conv_dgrad', 2, 32, 32, 64), ('convT_dgrad', 2, 16, 64, 32), ('conv', 1, 256, 32, 64),
             ('convT', 1, 128, 32, 32)]
    for op, B, H, Cin, Cout in cases:
        if op == 'conv':
            x = rng.standard_normal((B, H, H, Cin)).astype(np.float32)
            w = (rng.standard_normal((5, 5, Cin, Cout)) / math.sqrt(25 * Cin)).astype(np.float32)
        elif op == 'convT':
            x = rng.standard_normal((B, H, H, Cin)).astype(np.float32)
            w = (rng.standard_normal((5, 5, Cout, Cin)) / math.sqrt(6 * Cin)).astype(np.float32)
        elif op == 'conv_dgrad':
            x = rng.standard_normal((B, H // 2, H // 2, Cout)).astype(np.float32)
            w = (rng.standard_normal((5, 5, Cin, Cout)) / math.sqrt(6 * Cout)).astype(np.float32)
        else:
            x = rng.standard_normal((B, 2 * H, 2 * H, Cout)).astype(np.float32)
            w = (rng.standard_normal((5, 5, Cout, Cin)) / math.sqrt(25 * Cout)).astype(np.float32)
        b = rng.standard_normal(Cout).astype(np.float32) if op in ('conv', 'convT') else None
        ref = run(op, B, H, Cin, Cout, 0, x, w, b).astype(np.float64)
        got = run(op, B, H, Cin, Cout, 1, x, w, b).astype(np.float64)
        err = np.abs(got - ref)
        rel = err.max() / np.abs(ref).max()
        nan = int(np.isnan(got).sum())
        print(f'{op:12s} B={B} H={H} Cin={Cin} Cout={Cout}: rel={rel:.3e} nan={nan} '
              f'tc_supported={abi.lib().uad_conv_tc_supported({"conv": 0, "convT": 3, "conv_dgrad": 1, "convT_dgrad": 4}[op], B, H, H, Cin, Cout, 5)}')
        if not (rel < 1e-4) or nan:
            e = np.nan_to_num(err, nan=1e9)
            print('   worst per batch      ', e.max(axis=(1, 2, 3)))
            print('   worst per row (b=0)  ', np.round(e[0].max(axis=(1, 2)), 4)[:16])
            print('   worst per col (b=0)  ', np.round(e[0].max(axis=(0, 2)), 4)[:16])
            print('   worst per channel    ', np.round(e.max(axis=(0, 1, 2)), 4))
            print('   got[0,0,0,:8]', got[0, 0, 0, :8], '\n   ref[0,0,0,:8]', ref[0, 0, 0, :8])
            print('   got[0,3,5,:8]', got[0, 3, 5, :8], '\n   ref[0,3,5,:8]', ref[0, 3, 5, :8])


if __name__ == '__main__':
    main()
