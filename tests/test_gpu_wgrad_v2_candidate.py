"""GPU, OPT-IN: first hardware check of the plane-resident Form-W candidate kernel `wgrad_tc2` (csrc/uad_conv_tc.cu,
csrc/uad_wgrad_tiles.h).  Written after round 1's GPU budget was spent; never run, and it presumes the answer of
experiment E7 of tools/ubench/operand_probe.cu (tcgen05.mma A operand at an arbitrary tensor-memory column) - run the probe
first.  Skipped unless the process is started with UAD_WGRAD_V2=1 (the launcher reads the switch once):
    UAD_WGRAD_V2=1 python -m pytest tests/test_gpu_wgrad_v2_candidate.py -m gpu
Compares the tcgen05 path (3xTF32) with the exact-fp32 SIMT path for both filter-gradient ops."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get('UAD_WGRAD_V2') != '1', reason='opt-in: UAD_WGRAD_V2=1')]


def _run(op, B, H, Cin, Cout, mode, accumulate=0, seed=0):
    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    from unsupervised_anomaly_detection_brain_mri_b200.abi import call
    L = abi.lib()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device='cuda').manual_seed(seed)
    up = op == 'convT_wgrad'
    wsb = L.uad_conv_workspace_bytes(5 if up else 2, B, H, H, Cin, Cout, 5, mode)
    ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
    x = torch.randn(B, H, H, Cin, device='cuda', generator=g)
    dz = torch.randn(B, 2 * H if up else H // 2, 2 * H if up else H // 2, Cout, device='cuda', generator=g)
    dw = torch.full((5, 5, *((Cout, Cin) if up else (Cin, Cout))), 0.25, device='cuda')
    call('uad_convT2d_wgrad' if up else 'uad_conv2d_wgrad', x.data_ptr(), dz.data_ptr(), dw.data_ptr(), B, H, H, Cin, Cout, 5, accumulate,
         mode, ws.data_ptr(), wsb, st)
    torch.cuda.synchronize()
    return dw.cpu().numpy()


# Form W: conv_wgrad gathers x (Cg = Cin, Co = Cout); convT_wgrad gathers dz (Cg = Cout, Co = Cin).  Window tiles per CTA: 5 / 3 / 2 at Co = 32 / 64 / 128.
@pytest.mark.parametrize('op,B,H,Cin,Cout', [('convT_wgrad', 64, 128, 32, 32), ('convT_wgrad', 16, 64, 64, 32), ('conv_wgrad', 64, 128, 32, 64),
                                             ('conv_wgrad', 16, 64, 64, 64), ('convT_wgrad', 8, 32, 64, 128), ('conv_wgrad', 3, 16, 32, 32),
                                             ('convT_wgrad', 2, 8, 32, 32), ('conv_wgrad', 64, 64, 64, 128), ('conv_wgrad', 16, 32, 128, 128),
                                             ('convT_wgrad', 16, 16, 128, 128)])
@pytest.mark.parametrize('accumulate', [0, 1])
def test_plane_resident_wgrad_matches_fp32_simt(op, B, H, Cin, Cout, accumulate):
    a = _run(op, B, H, Cin, Cout, 1, accumulate)
    b = _run(op, B, H, Cin, Cout, 0, accumulate)
    err = float(np.abs(a - b).max() / np.abs(b).max())
    assert err < 3e-5, err
    assert np.array_equal(a, _run(op, B, H, Cin, Cout, 1, accumulate))       # deterministic across runs
