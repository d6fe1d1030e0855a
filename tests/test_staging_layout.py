"""The epilogue staging tile's index arithmetic (csrc/uad_staging.h, included by gather_gemm_tc2) compiled for the HOST with g++:
the write -> read round trip returns every element to the lane that stores it, both layouts stay inside their buffer, and every
quarter-warp (the unit in which 128-bit shared-memory accesses are served) touches eight distinct 16-byte bank groups."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, 'unsupervised_anomaly_detection_brain_mri_b200', 'csrc')

SHIM = r'''
#include "uad_staging.h"
extern "C" {
int ld(int swz) { return swz ? uad_stg_ld<true>() : uad_stg_ld<false>(); }
int widx(int swz, int lane, int j) { return swz ? uad_stg_write_index<true>(lane, j) : uad_stg_write_index<false>(lane, j); }
int ridx(int swz, int lane, int it) { return swz ? uad_stg_read_index<true>(lane, it) : uad_stg_read_index<false>(lane, it); }
int rrow(int lane, int it) { return uad_stg_read_row(lane, it); }
}
'''


@pytest.fixture(scope='module')
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp('stg')
    src = d / 'shim.cpp'
    src.write_text(SHIM)
    so = d / 'shim.so'
    subprocess.check_call(['g++', '-O1', '-shared', '-fPIC', '-std=c++17', '-I', HDR, str(src), '-o', str(so)])
    return C.CDLL(str(so))


@pytest.mark.parametrize('swz', [0, 1])
def test_staging_round_trip_and_banks(lib, swz):
    ld = lib.ld(swz)
    assert ld == (32 if swz else 36)
    buf = np.full(32 * ld, -1, np.int64)
    for lane in range(32):                                   # write phase: lane == tile row, value = row * 32 + column
        for j in range(0, 32, 4):
            i = lib.widx(swz, lane, j)
            assert i % 4 == 0 and 0 <= i and i + 4 <= 32 * ld
            assert (buf[i:i + 4] == -1).all()                # no two (row, group) pairs share a slot
            buf[i:i + 4] = lane * 32 + j + np.arange(4)
    seen = set()
    for it in range(8):                                      # read phase: the lane stores row r, columns (lane & 7) * 4 .. + 3
        for lane in range(32):
            r, cq = lib.rrow(lane, it), (lane & 7) * 4
            i = lib.ridx(swz, lane, it)
            assert list(buf[i:i + 4]) == [r * 32 + cq + e for e in range(4)]
            seen.add((r, cq))
    assert len(seen) == 32 * 8                               # every float4 of the tile is stored exactly once
    bank_group = lambda i: (i // 4) % 8                      # noqa: E731   16-byte group within the 128-byte bank row
    for j in range(0, 32, 4):
        for q in range(4):
            assert len({bank_group(lib.widx(swz, lane, j)) for lane in range(8 * q, 8 * q + 8)}) == 8
    for it in range(8):
        for q in range(4):
            assert len({bank_group(lib.ridx(swz, lane, it)) for lane in range(8 * q, 8 * q + 8)}) == 8


def test_swizzled_staging_fits_four_stages_at_n128():
    """Shared-memory budget behind UAD_TC_V2 bit 16 (uad_conv_tc.cu launcher): 1024 alignment slack + stages + barriers +
    epilogue constants + staging + 64 must fit the 227 KB a block may own, with an EVEN ring."""
    for N, swz, want in ((128, 1, 4), (128, 0, 2), (64, 0, 4), (64, 1, 6)):
        stage = 128 * 128 + 2 * N * 128
        tail = 256 + 3 * N * 4 + 8 * 32 * (32 if swz else 36) * 4 + 64
        stages = min(8, ((227 if swz else 226) * 1024 - 1024 - tail) // stage) & ~1
        assert stages == want
        assert 1024 + stages * stage + tail <= 227 * 1024
