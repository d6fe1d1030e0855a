"""Window arithmetic of the halo-resident gather candidate (csrc/uad_halo.h, included by gather_gemm_tc_np_halo) compiled for the
HOST with g++: for every output-parity class and tap of the transposed form (the tap tables of csrc/uad_conv_api.cu::taps_parity)
the halo row a tile pixel reads equals the zero-padded gather the per-k-block TMA tiles of the shipped kernel deliver."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, 'unsupervised_anomaly_detection_brain_mri_b200', 'csrc')

SHIM = '#include "uad_halo.h"\nextern "C" int halo_row(int th, int tw, int dh, int dw, int TW) { return uad_halo_row(th, tw, dh, dw, TW); }\n'


@pytest.fixture(scope='module')
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp('halo')
    (d / 'shim.cpp').write_text(SHIM)
    subprocess.check_call(['g++', '-O1', '-shared', '-fPIC', '-std=c++17', '-I', HDR, str(d / 'shim.cpp'), '-o', str(d / 'shim.so')])
    return C.CDLL(str(d / 'shim.so'))


def parity_taps(p, q, k=5, lo=1):
    """taps_parity of uad_conv_api.cu: fine pixel (2r + p, 2s + q) <- coarse pixel (r + dh, s + dw)."""
    return [((p - kh + lo) // 2, (q - kw + lo) // 2) for kh in range(k) if (p - kh + lo) % 2 == 0 for kw in range(k) if (q - kw + lo) % 2 == 0]


def test_halo_rows_equal_the_zero_padded_gather(lib):
    rng = np.random.default_rng(0)
    H, W, TH, TW = 16, 32, 8, 16
    img = rng.standard_normal((H, W))
    taps = [t for p in range(2) for q in range(2) for t in parity_taps(p, q)]
    assert len(taps) == 25 and {t for t in taps} == {(a, b) for a in (-1, 0, 1) for b in (-1, 0, 1)}      # 25 taps, 9 distinct shifts
    for r0 in range(0, H, TH):
        for s0 in range(0, W, TW):
            halo = np.zeros((TH + 2) * (TW + 2))                       # TMA box at (s0 - 1, r0 - 1), zero fill outside the image
            for hr in range(TH + 2):
                for hc in range(TW + 2):
                    y, x = r0 - 1 + hr, s0 - 1 + hc
                    if 0 <= y < H and 0 <= x < W:
                        halo[hr * (TW + 2) + hc] = img[y, x]
            for dh, dw in set(taps):
                for th in range(TH):
                    for tw in range(TW):
                        y, x = r0 + th + dh, s0 + tw + dw
                        want = img[y, x] if 0 <= y < H and 0 <= x < W else 0.0
                        row = lib.halo_row(th, tw, dh, dw, TW)
                        assert 0 <= row < (TH + 2) * (TW + 2) and halo[row] == want


def test_halo_shared_memory_budget_allows_two_ctas_per_sm():
    """Launcher arithmetic (uad_conv_tc.cu, UAD_TC_HALO branch) at N = 32: four 8 KB weight stages, staging rows, one 23 KB halo per
    32-channel block of the gathered tensor."""
    N = 32
    for cblks, two_ctas in ((1, True), (2, True), (4, False)):
        stage = 2 * N * 128
        halo = (10 * 18 * 128 + 1023) // 1024 * 1024
        off = (4 * stage + 256 + 3 * N * 4 + 4 * 32 * (N + 4) * 4 + 1023) // 1024 * 1024
        smem = 1024 + off + cblks * halo + 64
        assert smem <= 200 * 1024
        assert (2 * (smem + 1024) <= 228 * 1024) == two_ctas
