"""Index arithmetic of the plane-resident Form-W candidate kernel (csrc/uad_wgrad_tiles.h, included by wgrad_tc2) compiled for
the HOST with g++ and checked against the tap geometry the shipped kernel uses (taps_full in csrc/uad_conv_api.cu: dh = kh - 1;
wgrad_tc: plane = ((dh & 1) << 1) | (dw & 1), window start = ((dh >> 1) + 1, (dw >> 1) + 1))."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, 'unsupervised_anomaly_detection_brain_mri_b200', 'csrc')

SHIM = r'''
#include "uad_wgrad_tiles.h"
extern "C" {
int tap(int tile, int plane) { return uad_wt_tap(tile, plane); }
int a_column(int tile, int r) { return uad_wt_a_column(tile, r); }
}
'''


@pytest.fixture(scope='module')
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp('wt')
    (d / 'shim.cpp').write_text(SHIM)
    subprocess.check_call(['g++', '-O1', '-shared', '-fPIC', '-std=c++17', '-I', HDR, str(d / 'shim.cpp'), '-o', str(d / 'shim.so')])
    return C.CDLL(str(d / 'shim.so'))


def test_every_tap_has_exactly_one_tile_and_plane(lib):
    seen = {}
    for tile in range(9):
        for plane in range(4):
            t = lib.tap(tile, plane)
            if t >= 0:
                assert t not in seen
                seen[t] = (tile, plane)
    assert sorted(seen) == list(range(25))
    for t, (tile, plane) in seen.items():                  # the shipped kernel's geometry for the same tap
        dh, dw = t // 5 - 1, t % 5 - 1
        assert plane == ((dh & 1) << 1) | (dw & 1)
        assert (tile // 3, tile % 3) == ((dh >> 1) + 1, (dw >> 1) + 1)
    full = [sum(lib.tap(tile, p) >= 0 for p in range(4)) for tile in range(9)]
    assert sorted(full) == [1, 2, 2, 2, 2, 4, 4, 4, 4]


def test_window_columns_reproduce_the_stride2_gather(lib):
    """Brute force on a random fine image: sum over the 4 x 8 block of fine[2R + dh, 2S + dw] * o[R, S] read through the
    plane-copy columns equals the direct gather, for every tap."""
    rng = np.random.default_rng(0)
    GH, GW = 24, 40
    fine = rng.standard_normal((GH, GW))
    r0, s0 = 4, 8                                             # block origin (coarse), interior so that no zero fill is involved
    o = rng.standard_normal((4, 8))
    planes = np.zeros((4, 60))
    for pl in range(4):                                       # what the TMA box (32 ch, 10, 1, 6, 1) at (s0 - 1, r0 - 1) delivers
        for hr in range(6):
            for hc in range(10):
                planes[pl, hr * 10 + hc] = fine[2 * (r0 - 1 + hr) + (pl >> 1), 2 * (s0 - 1 + hc) + (pl & 1)]
    for tile in range(9):
        for pl in range(4):
            t = lib.tap(tile, pl)
            if t < 0:
                continue
            dh, dw = t // 5 - 1, t % 5 - 1
            want = sum(fine[2 * (r0 + r) + dh, 2 * (s0 + k) + dw] * o[r, k] for r in range(4) for k in range(8))
            got = 0.0
            for r in range(4):
                c = lib.a_column(tile, r)
                assert 0 <= c and c + 8 <= 60
                got += float(planes[pl, c:c + 8] @ o[r])
            assert abs(got - want) < 1e-12
