"""GPU, OPT-IN: first hardware check of the SS-form candidate kernel `gather_gemm_ss` (csrc/uad_conv_tc.cu): activations split
once per call into tf32 {hi, lo} images, both operands read from shared memory through descriptors, no converter role.
Written after round 1's GPU budget was spent; never run, so skipped unless the process is started with UAD_TC_SS=7
(1: N = 128 layers, 2: N = 64, 4: N = 32; the launcher and the workspace-size query read the switch once):
    UAD_TC_SS=7 python -m pytest tests/test_gpu_ss_candidate.py -m gpu
UAD_TC_SS=15 additionally feeds the RAW fp32 tensor as the hi operand (only the lo image is written): passes iff kind::tf32
truncates its operands (experiment E1 of tools/ubench/operand_probe.cu) - a failure there is an answer, not a bug.
Compares the tcgen05 path (3xTF32) with the exact-fp32 SIMT path on shapes of both GEMM forms and all three widths."""
import os

import numpy as np
import pytest

from test_gpu_swz_candidate import _run

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get('UAD_TC_SS') not in ('7', '15'), reason='opt-in: UAD_TC_SS=7 (or 15)')]


# GEMM N: conv_fwd / convT_fwd -> Cout, conv_dgrad / convT_dgrad -> Cin
@pytest.mark.parametrize('op,B,H,Cin,Cout', [('conv_fwd', 64, 64, 64, 128), ('conv_fwd', 64, 32, 128, 128), ('conv_dgrad', 64, 64, 128, 128),
                                             ('convT_fwd', 64, 16, 128, 128), ('convT_dgrad', 64, 32, 128, 64), ('conv_fwd', 64, 128, 32, 64),
                                             ('conv_dgrad', 32, 128, 64, 128), ('convT_fwd', 64, 32, 128, 64), ('convT_fwd', 16, 64, 64, 32),
                                             ('convT_dgrad', 16, 128, 32, 32), ('conv_dgrad', 32, 128, 32, 64), ('conv_fwd', 3, 16, 32, 128),
                                             ('convT_fwd', 3, 16, 32, 32)])
def test_ss_form_matches_fp32_simt(op, B, H, Cin, Cout):
    a = _run(op, B, H, Cin, Cout, 1)
    b = _run(op, B, H, Cin, Cout, 0)
    err = float(np.abs(a - b).max() / np.abs(b).max())
    assert err < 2e-5, err
    assert np.array_equal(a, _run(op, B, H, Cin, Cout, 1))       # deterministic across runs (no schedule-dependent result)
