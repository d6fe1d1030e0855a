"""GPU, OPT-IN: first hardware check of the halo-resident candidate `gather_gemm_tc_np_halo` (csrc/uad_conv_tc.cu): N = 32 layers
of the stride-1 form (transposed-conv forward, conv input gradient) load the 10 x 18 pixel halo of an 8 x 16 tile once per
channel block and convert every tap's window from it - a k-block then moves only its 8 KB weight image.
Written after round 1's GPU budget was spent; never run, so skipped unless the process is started with UAD_TC_HALO=1:
    UAD_TC_HALO=1 python -m pytest tests/test_gpu_halo_candidate.py -m gpu
Compares the tcgen05 path (3xTF32) with the exact-fp32 SIMT path; the last two shapes are too small for the halo tiling and
must take the shipped kernel."""
import os

import numpy as np
import pytest

from test_gpu_swz_candidate import _run

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get('UAD_TC_HALO') != '1', reason='opt-in: UAD_TC_HALO=1')]


# GEMM N: convT_fwd -> Cout, conv_dgrad -> Cin (both must be 32 here); channel blocks of the gathered tensor: Cin / 32, Cout / 32
@pytest.mark.parametrize('op,B,H,Cin,Cout', [('convT_fwd', 64, 128, 32, 32), ('convT_fwd', 16, 64, 64, 32), ('convT_fwd', 8, 32, 128, 32),
                                             ('conv_dgrad', 32, 128, 32, 64), ('conv_dgrad', 8, 64, 32, 32), ('convT_fwd', 3, 16, 32, 32),
                                             ('convT_fwd', 5, 8, 32, 32), ('conv_dgrad', 4, 16, 32, 64)])
def test_halo_form_matches_fp32_simt(op, B, H, Cin, Cout):
    a = _run(op, B, H, Cin, Cout, 1)
    b = _run(op, B, H, Cin, Cout, 0)
    err = float(np.abs(a - b).max() / np.abs(b).max())
    assert err < 2e-5, err
    assert np.array_equal(a, _run(op, B, H, Cin, Cout, 1))       # deterministic across runs (no schedule-dependent result)
