"""Helpers for the -m gpu parity tests: call the C ABI with numpy in / numpy out."""
import numpy as np
import torch

from unsupervised_anomaly_detection_brain_mri_b200 import abi  # noqa: F401  (re-exported to the tests)
from unsupervised_anomaly_detection_brain_mri_b200.abi import call, ptr  # noqa: F401

DEV = 'cuda:0'


def dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).to(DEV)


def empty(*shape):
    return torch.full(shape, float('nan'), dtype=torch.float32, device=DEV)


def st():
    return torch.cuda.current_stream().cuda_stream


def workspace(nbytes):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=DEV)


def relerr(a, b):
    """The parity metric of SURVEY 8c: ||a-b||_inf / max(||b||_inf, tiny)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))


def sync():
    torch.cuda.synchronize()


_KEEP = []


def dptr(a, dtype=torch.float32):
    """Upload and return the device pointer, keeping the tensor alive until the test ends (see conftest autouse fixture)."""
    t = dev(a, dtype)
    _KEEP.append(t)
    return t.data_ptr()
