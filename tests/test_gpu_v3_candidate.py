"""GPU, OPT-IN: first hardware check of the round-2 candidate kernel `gather_gemm_tc3` (N = 32 layers; see csrc/uad_conv_tc.cu).
The kernel was written after round 1's GPU budget was spent and has never run, so this file is skipped unless the process is
started with UAD_TC_V3=1 (the launcher reads the switch once):   UAD_TC_V3=1 python -m pytest tests/test_gpu_v3_candidate.py -m gpu
It compares the tcgen05 path (3xTF32) with the exact-fp32 SIMT path on multi-item shapes of both GEMM forms."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get('UAD_TC_V3') != '1', reason='opt-in: UAD_TC_V3=1')]


def _run(op, B, H, Cin, Cout, mode, seed=0):
    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    from unsupervised_anomaly_detection_brain_mri_b200.abi import call
    L = abi.lib()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device='cuda').manual_seed(seed)
    opid = {'convT_fwd': 3, 'convT_dgrad': 4, 'conv_dgrad': 1}[op]
    wsb = L.uad_conv_workspace_bytes(opid, B, H, H, Cin, Cout, 5, mode)
    ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
    if op.startswith('convT'):
        x = torch.randn(B, H, H, Cin, device='cuda', generator=g)
        y = torch.randn(B, 2 * H, 2 * H, Cout, device='cuda', generator=g)
        w = torch.randn(5, 5, Cout, Cin, device='cuda', generator=g) * 0.05
    else:
        x = torch.randn(B, H, H, Cin, device='cuda', generator=g)
        y = torch.randn(B, H // 2, H // 2, Cout, device='cuda', generator=g)
        w = torch.randn(5, 5, Cin, Cout, device='cuda', generator=g) * 0.05
    if op == 'convT_fwd':
        out = torch.empty_like(y)
        bias = torch.randn(Cout, device='cuda', generator=g) * 0.1
        call('uad_convT2d_fwd', x.data_ptr(), w.data_ptr(), bias.data_ptr(), None, None, None, out.data_ptr(), B, H, H, Cin, Cout, 5, 1,
             0.3, 1.0, mode, ws.data_ptr(), wsb, st)
    elif op == 'convT_dgrad':
        out = torch.empty_like(x)
        call('uad_convT2d_dgrad', y.data_ptr(), w.data_ptr(), out.data_ptr(), B, H, H, Cin, Cout, 5, mode, ws.data_ptr(), wsb, st)
    else:
        out = torch.empty_like(x)
        call('uad_conv2d_dgrad', y.data_ptr(), w.data_ptr(), out.data_ptr(), B, H, H, Cin, Cout, 5, mode, ws.data_ptr(), wsb, st)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize('op,B,H,Cin,Cout', [('convT_fwd', 64, 128, 32, 32), ('convT_fwd', 16, 64, 64, 32), ('convT_dgrad', 64, 128, 32, 32),
                                             ('conv_dgrad', 32, 128, 32, 64), ('convT_fwd', 3, 16, 32, 32)])
def test_v3_matches_fp32_simt(op, B, H, Cin, Cout):
    a = _run(op, B, H, Cin, Cout, 1)
    b = _run(op, B, H, Cin, Cout, 0)
    err = float(np.abs(a - b).max() / np.abs(b).max())
    assert err < 2e-5, err
    a2 = _run(op, B, H, Cin, Cout, 1)
    assert np.array_equal(a, a2)                 # deterministic across runs (no schedule-dependent result)
