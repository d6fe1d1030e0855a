"""Evaluation post-processing stencils on the device, bit-exact against scipy (the reference's own calls):
brain-mask erosion (utils/Evaluation.py:84-89) and the 5x5x5 median filter (:108-110, applied at :311-312)."""
import numpy as np
import pytest
import scipy.ndimage
import torch

pytestmark = pytest.mark.gpu


def _st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize('N,H,W,it', [(3, 64, 64, 12), (2, 50, 77, 12), (1, 256, 256, 12), (2, 40, 33, 1), (1, 31, 31, 24)])
def test_binary_erosion_cross_matches_scipy(N, H, W, it):
    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    rng = np.random.default_rng(N * 1000 + H)
    yy, xx = np.mgrid[0:H, 0:W]
    m = np.zeros((N, H, W), np.uint8)
    for n in range(N):
        blob = ((yy - H / 2) / (0.45 * H)) ** 2 + ((xx - W / 2) / (0.40 * W)) ** 2 <= 1.0
        holes = rng.uniform(size=(H, W)) < 0.002
        m[n] = (blob & ~holes).astype(np.uint8) * (1 + n)          # non-{0,1} values count as set
    m[0, :, :3] = 1                                                # touches the border: border_value = 0 erodes it
    d_in = torch.from_numpy(m).cuda()
    d_out = torch.empty_like(d_in)
    abi.call('uad_binary_erosion_cross', d_in.data_ptr(), d_out.data_ptr(), N, H, W, it, _st())
    strel = scipy.ndimage.generate_binary_structure(2, 1)
    ref = np.stack([scipy.ndimage.binary_erosion(m[n], structure=strel, iterations=it) for n in range(N)])
    assert np.array_equal(d_out.cpu().numpy().astype(bool), ref)


@pytest.mark.parametrize('Z,H,W', [(12, 40, 40), (5, 17, 70), (2, 9, 9), (24, 64, 64)])
def test_median_filter3d_matches_scipy(Z, H, W):
    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    rng = np.random.default_rng(Z * 100 + W)
    v = rng.random((Z, H, W), dtype=np.float32)
    v[rng.uniform(size=v.shape) < 0.5] = 0.0                       # residual maps are ~half exact zeros
    v[:, : H // 3] = 0.0                                           # constant regions take the early exit
    v[0, -1, -1] = -0.25                                           # a negative value exercises the key transform
    d_in = torch.from_numpy(v).cuda()
    d_out = torch.empty_like(d_in)
    abi.call('uad_median_filter3d_5', d_in.data_ptr(), d_out.data_ptr(), Z, H, W, _st())
    ref = scipy.ndimage.median_filter(v.astype(np.float64), (5, 5, 5))
    assert np.array_equal(d_out.cpu().numpy().astype(np.float64), ref)


def test_score_volume_pipeline_matches_host_reference():
    """erosion -> residual/mask/prior -> median on the device == the reference's per-slice numpy/scipy sequence."""
    from oracle import scoring as OS
    from unsupervised_anomaly_detection_brain_mri_b200.utils import Evaluation
    rng = np.random.default_rng(3)
    Z, S = 14, 96
    from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import make_volume
    vol, seg, skull = make_volume(S, Z, 5, lesions=True)
    rec = np.clip(vol + 0.1 * rng.standard_normal(vol.shape).astype(np.float32), 0, 1).astype(np.float32)
    prior = float(np.quantile(vol, 0.9))
    got = Evaluation.score_volume_on_device(vol, rec, skull, 12, prior, True, True, True, 'cuda:0')
    masks = np.stack([OS.erode_brainmask(skull[s]) for s in range(Z)])
    sub = OS.residual(vol, rec, masks, prior, keep_positive=True, apply_prior=True)
    ref = scipy.ndimage.median_filter(sub, (5, 5, 5))
    assert got.dtype == np.float64 and np.array_equal(got, ref)
