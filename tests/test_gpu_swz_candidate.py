"""GPU, OPT-IN: first hardware check of the swizzled epilogue staging of `gather_gemm_tc2` (kernel `gather_gemm_tc2_swz`,
csrc/uad_conv_tc.cu + csrc/uad_staging.h) and, with it, of the N = 128 column-split dual issue on an EVEN four-stage ring.
Written after round 1's GPU budget was spent; never run, so skipped unless the process is started with UAD_TC_V2=21
(1: N = 64 on v2, 4: N = 128 column split, 16: swizzled staging; the launcher reads the switch once):
    UAD_TC_V2=21 python -m pytest tests/test_gpu_swz_candidate.py -m gpu
Compares the tcgen05 path (3xTF32) with the exact-fp32 SIMT path on N = 64 and N = 128 shapes of both GEMM forms."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get('UAD_TC_V2') != '21', reason='opt-in: UAD_TC_V2=21')]


def _run(op, B, H, Cin, Cout, mode, seed=0):
    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    from unsupervised_anomaly_detection_brain_mri_b200.abi import call
    L = abi.lib()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device='cuda').manual_seed(seed)
    opid = {'conv_fwd': 0, 'conv_dgrad': 1, 'convT_fwd': 3, 'convT_dgrad': 4}[op]
    wsb = L.uad_conv_workspace_bytes(opid, B, H, H, Cin, Cout, 5, mode)
    ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
    up = op.startswith('convT')
    x = torch.randn(B, H, H, Cin, device='cuda', generator=g)
    y = torch.randn(B, 2 * H if up else H // 2, 2 * H if up else H // 2, Cout, device='cuda', generator=g)
    w = torch.randn(5, 5, *((Cout, Cin) if up else (Cin, Cout)), device='cuda', generator=g) * 0.05
    bias = torch.randn(Cout, device='cuda', generator=g) * 0.1
    gamma = 1 + 0.1 * torch.randn(Cout, device='cuda', generator=g)
    beta = 0.1 * torch.randn(Cout, device='cuda', generator=g)
    if op in ('conv_fwd', 'convT_fwd'):
        z, a = torch.empty_like(y), torch.empty_like(y)
        call('uad_conv2d_fwd' if op == 'conv_fwd' else 'uad_convT2d_fwd', x.data_ptr(), w.data_ptr(), bias.data_ptr(), gamma.data_ptr(),
             beta.data_ptr(), z.data_ptr(), a.data_ptr(), B, H, H, Cin, Cout, 5, 1, 0.3, 0.9995, mode, ws.data_ptr(), wsb, st)
        out = torch.cat([z, a])
    else:
        out = torch.empty_like(x)
        call('uad_conv2d_dgrad' if op == 'conv_dgrad' else 'uad_convT2d_dgrad', y.data_ptr(), w.data_ptr(), out.data_ptr(), B, H, H, Cin,
             Cout, 5, mode, ws.data_ptr(), wsb, st)
    torch.cuda.synchronize()
    return out.cpu().numpy()


# GEMM N: conv_fwd / convT_fwd -> Cout, conv_dgrad / convT_dgrad -> Cin
@pytest.mark.parametrize('op,B,H,Cin,Cout', [('conv_fwd', 64, 64, 64, 128), ('conv_fwd', 64, 32, 128, 128), ('conv_fwd', 64, 128, 32, 64),
                                             ('conv_dgrad', 64, 64, 128, 128), ('conv_dgrad', 16, 64, 64, 128), ('convT_fwd', 64, 16, 128, 128),
                                             ('convT_fwd', 64, 32, 128, 64), ('convT_dgrad', 64, 32, 128, 64), ('conv_fwd', 3, 16, 32, 128)])
def test_swizzled_staging_matches_fp32_simt(op, B, H, Cin, Cout):
    a = _run(op, B, H, Cin, Cout, 1)
    b = _run(op, B, H, Cin, Cout, 0)
    err = float(np.abs(a - b).max() / np.abs(b).max())
    assert err < 2e-5, err
    assert np.array_equal(a, _run(op, B, H, Cin, Cout, 1))       # deterministic across runs (no schedule-dependent result)
