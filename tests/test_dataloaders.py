"""Volume readers and the BrainWeb slice loader (CPU, no GPU, no /root/reference at run time).

What pins them: MINC-1 files are NetCDF classic containers - the test volumes are WRITTEN by scipy.io.netcdf_file (an
independent implementation of the container) with the attributes BrainWeb files carry, and the expected real values are
the MINC-1 rule computed directly; NIfTI-1 headers are checked field by field at their nifti1.h offsets; the loader runs
on a synthetic dataset directory in the reference's layout (dataloaders/BRAINWEB.py:27-56, 209-251)."""
import gzip
import os
import pickle
import struct

import numpy as np
import pytest

from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.BRAINWEB import BRAINWEB
from unsupervised_anomaly_detection_brain_mri_b200.utils import image_utils
from unsupervised_anomaly_detection_brain_mri_b200.utils.MINC import MINC, read_minc1
from unsupervised_anomaly_detection_brain_mri_b200.utils.NII import NII, read_nifti, write_nifti


def write_minc1(path, raw_zyx, imin, imax, valid_range, signtype, steps=(1.0, 1.0, 1.0), starts=(0.0, 0.0, 0.0)):
    """A MINC-1 volume the way BrainWeb ships them: image[zspace,yspace,xspace] + per-slice image-min / image-max."""
    from scipy.io import netcdf_file
    tmp = path[:-3] if path.endswith('.gz') else path
    nc = netcdf_file(tmp, 'w')
    Z, Y, X = raw_zyx.shape
    for name, n, step, start in (('zspace', Z, steps[2], starts[2]), ('yspace', Y, steps[1], starts[1]),
                                 ('xspace', X, steps[0], starts[0])):
        nc.createDimension(name, n)
        v = nc.createVariable(name, 'i', ())
        v.step, v.start = float(step), float(start)
        v.data[...] = 0
    img = nc.createVariable('image', raw_zyx.dtype.char if raw_zyx.dtype.kind == 'i' else 'b', ('zspace', 'yspace', 'xspace'))
    img[:] = raw_zyx.view(np.dtype(f'i{raw_zyx.dtype.itemsize}'))
    img.signtype = signtype
    img.valid_range = np.array(valid_range, np.float64)
    for name, val in (('image-min', imin), ('image-max', imax)):
        val = np.asarray(val, np.float64)
        v = nc.createVariable(name, 'd', ('zspace',) if val.ndim else ())
        if val.ndim:
            v[:] = val
        else:
            v.data[...] = float(val)
    nc.close()
    if path.endswith('.gz'):
        with open(tmp, 'rb') as f, gzip.open(path, 'wb') as g:
            g.write(f.read())
        os.remove(tmp)


# ---------------------------------------------------------------------------------------------------- NIfTI
def test_nifti_round_trip_and_header_fields(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.standard_normal((5, 6, 7)).astype(np.float32)                      # [k, j, i]
    p = str(tmp_path / 'v.nii.gz')
    write_nifti(p, a, origin=(1.5, -2.0, 3.25), spacing=(1.0, 2.0, 0.5))
    raw = gzip.open(p, 'rb').read()
    assert struct.unpack('<i', raw[:4])[0] == 348 and raw[344:348] == b'n+1\x00'
    assert struct.unpack('<8h', raw[40:56])[:4] == (3, 7, 6, 5)                # dim[1..3] = (i, j, k)
    assert struct.unpack('<h', raw[70:72])[0] == 16 and struct.unpack('<h', raw[72:74])[0] == 32
    assert struct.unpack('<f', raw[108:112])[0] == 352.0
    assert len(raw) == 352 + a.size * 4
    data, origin, spacing = read_nifti(p)
    assert data.dtype == np.float64 and data.shape == (5, 6, 7)
    np.testing.assert_array_equal(data, a.astype(np.float64))
    assert origin == (1.5, -2.0, 3.25) and spacing == (1.0, 2.0, 0.5)
    # voxel (i, j, k) is stored i-fastest: element [k, j, i] of the array
    first = np.frombuffer(raw, '<f4', count=8, offset=352)
    np.testing.assert_array_equal(first[:7], a[0, 0, :])
    assert first[7] == a[0, 1, 0]


@pytest.mark.parametrize('dtype', [np.uint8, np.int16, np.int32, np.float64, np.uint16])
def test_nifti_dtypes_preserved(tmp_path, dtype):
    a = (np.arange(2 * 3 * 4).reshape(2, 3, 4) % 100).astype(dtype)
    p = str(tmp_path / 'v.nii')
    write_nifti(p, a)
    data, _, _ = read_nifti(p)
    np.testing.assert_array_equal(data, a.astype(np.float64))


def test_nifti_big_endian_and_scaling(tmp_path):
    a = np.arange(24, dtype=np.int16).reshape(2, 3, 4)
    hdr = bytearray(348)
    struct.pack_into('>i', hdr, 0, 348)
    struct.pack_into('>8h', hdr, 40, 3, 4, 3, 2, 1, 1, 1, 1)
    struct.pack_into('>h', hdr, 70, 4)
    struct.pack_into('>h', hdr, 72, 16)
    struct.pack_into('>8f', hdr, 76, 1, 1, 1, 1, 1, 1, 1, 1)
    struct.pack_into('>3f', hdr, 108, 352.0, 0.5, 10.0)                        # vox_offset, scl_slope, scl_inter
    hdr[344:348] = b'n+1\x00'
    p = str(tmp_path / 'be.nii')
    with open(p, 'wb') as f:
        f.write(bytes(hdr) + b'\0' * 4 + a.astype('>i2').tobytes())
    data, _, _ = read_nifti(p)
    np.testing.assert_array_equal(data, a * 0.5 + 10.0)


def test_nifti_rejects_garbage(tmp_path):
    p = str(tmp_path / 'bad.nii')
    with open(p, 'wb') as f:
        f.write(b'\0' * 400)
    with pytest.raises(IOError):
        read_nifti(p)
    with open(p, 'wb') as f:
        f.write(b'\0' * 10)
    with pytest.raises(IOError):
        read_nifti(p)


def test_nii_wrapper_semantics(tmp_path):
    rng = np.random.default_rng(1)
    a = rng.uniform(0, 100, (4, 5, 6))
    a[0, 0, 0] = np.nan
    v = NII(data=a)
    assert v.data[0, 0, 0] == 0                                                # NaNs removed (reference NII.py:15)
    NII.set_view_mapping({'saggital': 0, 'coronal': 1, 'axial': 2})
    assert (v.num_saggital_slices, v.num_coronal_slices, v.num_axial_slices) == (4, 5, 6)
    assert v.num_slices_along_axis('axial') == 6 and v.shape() == (4, 5, 6)
    np.testing.assert_array_equal(v.get_slice(2, 'axial'), v.data[:, :, 2])
    np.testing.assert_array_equal(v.get_slice(1, 'saggital'), v.data[1])
    w = v.copy()
    w.set_slice(3, np.ones((4, 5)), 'axial')
    assert (w.data[:, :, 3] == 1).all() and not (v.data[:, :, 3] == 1).all()
    sub = np.stack([np.full((4, 5), 7.0), np.full((4, 5), 8.0)])
    w.set_subvolume(1, 3, sub, 'axial')
    assert (w.data[:, :, 1] == 7).all() and (w.data[:, :, 2] == 8).all()
    # normalisation: percentile clip, then scaling / standardisation in float32 (reference NII.py:52-74)
    s = v.copy()
    s.normalize('scaling', lowerpercentile=0.0, upperpercentile=99.8)
    d = v.data.astype(np.float32)
    hi = np.percentile(d, 99.8)
    expect = np.minimum(d, hi)
    expect = expect * (1.0 / expect.max())
    np.testing.assert_array_equal(s.data, expect)
    assert s.data.dtype == np.float32 and abs(float(s.data.max()) - 1.0) < 1e-6
    t = v.copy()
    t.normalize('standardization')
    assert abs(float(t.data.mean())) < 1e-5 and abs(float(t.data.std()) - 1) < 1e-5
    # skull map: thresholded at 0.1 IN PLACE on the map, then multiplied in
    m = NII(data=np.where(a > 50, 0.5, 0.05))
    u = v.copy()
    u.apply_skullmap(m)
    assert set(np.unique(m.data)) <= {0.0, 1.0}
    np.testing.assert_array_equal(u.data, v.data * (np.nan_to_num(a) > 50))
    # save / load / subtract
    p = str(tmp_path / 'v.nii.gz')
    v.save(p)
    r = NII(p)
    np.testing.assert_array_equal(r.data, v.data)
    r.subtract(p)
    assert not r.data.any()
    r.data[:] = 3
    r.set_to_zero()
    assert not r.get_data().any()
    r.denoise()                                                               # constant volume: a fixed point of the curvature flow
    assert not r.get_data().any()


# ---------------------------------------------------------------------------------------------------- MINC-1
def test_minc1_real_value_rule_unsigned_byte(tmp_path):
    """unsigned bytes, valid_range (0, 255), per-slice image-min / image-max: real = (v - 0) / 255 * (max - min) + min."""
    rng = np.random.default_rng(2)
    raw = rng.integers(0, 256, (4, 5, 6)).astype(np.uint8)                     # [z, y, x]
    imin, imax = np.array([0.0, 1.0, -2.0, 0.5]), np.array([10.0, 3.0, 2.0, 0.5])
    p = str(tmp_path / 't2_x.mnc.gz')
    write_minc1(p, raw, imin, imax, (0, 255), 'unsigned', steps=(1.0, 2.0, 3.0), starts=(-90.0, -126.0, -72.0))
    data, origin, spacing = read_minc1(p)
    expect = raw.astype(np.float64) / 255.0 * (imax - imin)[:, None, None] + imin[:, None, None]
    assert data.shape == (6, 5, 4)                                             # [x, y, z]
    np.testing.assert_allclose(data, expect.transpose(2, 1, 0), rtol=0, atol=1e-12)
    assert origin == (-90.0, -126.0, -72.0) and spacing == (1.0, 2.0, 3.0)
    v = MINC(p)
    v.set_view_mapping(BRAINWEB.VIEW_MAPPING)
    assert v.num_slices_along_axis('axial') == 4
    np.testing.assert_allclose(v.get_slice(1, 'axial'), expect[1].T, atol=1e-12)


def test_minc1_signed_short_scalar_range(tmp_path):
    rng = np.random.default_rng(3)
    raw = rng.integers(-100, 4000, (3, 4, 5)).astype(np.int16)
    p = str(tmp_path / 'v.mnc')
    write_minc1(p, raw, 0.0, 8.0, (-100, 4095), 'signed__')
    data, _, _ = read_minc1(p)
    expect = (raw.astype(np.float64) + 100) / 4195.0 * 8.0
    np.testing.assert_allclose(data, expect.transpose(2, 1, 0), atol=1e-12)


def test_minc_label_volume_is_integer_valued(tmp_path):
    """BrainWeb's tissue labels: bytes 0..10 with image-min/max = valid_range -> real values are the labels themselves."""
    raw = (np.arange(2 * 3 * 11).reshape(2, 3, 11) % 11).astype(np.uint8)
    p = str(tmp_path / 'normal.mnc.gz')
    write_minc1(p, raw, 0.0, 10.0, (0, 10), 'unsigned')
    data, _, _ = read_minc1(p)
    np.testing.assert_allclose(data, raw.transpose(2, 1, 0), atol=1e-12)


def test_minc_rejects_other_containers(tmp_path):
    p = str(tmp_path / 'h5.mnc')
    with open(p, 'wb') as f:
        f.write(b'\x89HDF\r\n\x1a\n' + b'\0' * 64)
    with pytest.raises(IOError, match='MINC-2'):
        read_minc1(p)
    with open(p, 'wb') as f:
        f.write(b'garbage' * 10)
    with pytest.raises(IOError):
        read_minc1(p)


def test_minc_opens_nifti_too(tmp_path):
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    p = str(tmp_path / 'v.nii.gz')
    write_nifti(p, a)
    np.testing.assert_array_equal(MINC(p).data, a)


# ---------------------------------------------------------------------------------------------------- image helpers
def test_image_utils():
    img = np.arange(100).reshape(10, 10)
    np.testing.assert_array_equal(image_utils.crop(img, 2, 3, 4, 5), img[2:6, 3:8])
    np.testing.assert_array_equal(image_utils.crop_center(img, 4, 6), img[2:8, 3:7])
    rgb = image_utils.augment_prediction_and_groundtruth_to_image(np.full((2, 2), 0.25), [[1, 1], [0, 0]], [[1, 0], [1, 0]])
    assert rgb.shape == (2, 2, 3)
    assert tuple(rgb[0, 0]) == (0, 1, 0) and tuple(rgb[0, 1]) == (1, 0.5, 0) and tuple(rgb[1, 0]) == (1, 0, 0)
    assert tuple(rgb[1, 1]) == (0.25, 0.25, 0.25)


# ---------------------------------------------------------------------------------------------------- the loader
X, Y, Z = 24, 28, 12


def _dataset_dir(tmp_path, n_normal=4, n_ms=2):
    """<dir>/normal/t2_*.mnc.gz (+ one flair), <dir>/lesions/mild/t2_*.mnc.gz, <dir>/groundtruth/{normal,mild_lesions}.mnc.gz"""
    rng = np.random.default_rng(7)
    root = tmp_path / 'brainweb'
    for sub in ('normal', 'lesions/mild', 'lesions/moderate', 'lesions/severe', 'groundtruth'):
        os.makedirs(root / sub)
    zz, yy, xx = np.mgrid[0:Z, 0:Y, 0:X]
    r = np.sqrt(((yy - Y / 2) / (Y / 2)) ** 2 + ((xx - X / 2) / (X / 2)) ** 2)
    labels = np.zeros((Z, Y, X), np.uint8)
    labels[r < 0.95] = BRAINWEB.LABELS['SKULL']
    labels[r < 0.8] = BRAINWEB.LABELS['CSF']
    labels[r < 0.7] = BRAINWEB.LABELS['GM']
    labels[r < 0.5] = BRAINWEB.LABELS['WM']
    labels[0] = 0                                                              # a blank slice at the bottom
    ms = labels.copy()
    ms[4:8, 10:14, 9:13] = BRAINWEB.LABELS['LESION']
    write_minc1(str(root / 'groundtruth' / 'normal.mnc.gz'), labels, 0.0, 10.0, (0, 10), 'unsigned')
    write_minc1(str(root / 'groundtruth' / 'mild_lesions.mnc.gz'), ms, 0.0, 10.0, (0, 10), 'unsigned')
    raws = {}

    def volume(lab):
        v = (lab.astype(np.float64) * 20 + rng.integers(0, 20, lab.shape)) * (lab > 0)
        return v.astype(np.uint8)

    for i in range(n_normal):
        raws[f't2_n{i}.mnc.gz'] = volume(labels)
        write_minc1(str(root / 'normal' / f't2_n{i}.mnc.gz'), raws[f't2_n{i}.mnc.gz'], np.zeros(Z), np.full(Z, 255.0), (0, 255), 'unsigned')
    write_minc1(str(root / 'normal' / 'flair_n0.mnc.gz'), volume(labels), np.zeros(Z), np.full(Z, 255.0), (0, 255), 'unsigned')
    for i in range(n_ms):
        raws[f't2_m{i}.mnc.gz'] = volume(ms)
        write_minc1(str(root / 'lesions' / 'mild' / f't2_m{i}.mnc.gz'), raws[f't2_m{i}.mnc.gz'], np.zeros(Z), np.full(Z, 255.0), (0, 255), 'unsigned')
    return str(root), labels, ms, raws


def _options(root, **kw):
    o = BRAINWEB.Options()
    o.dir = root
    o.filterProtocol = 'T2'
    o.filterType = ['NORMAL', 'MILDMS']
    o.sliceStart, o.sliceEnd = 0, 140
    o.normalizationMethod = 'scaling'
    o.partition = {'TRAIN': 0.5, 'VAL': 0.25, 'TEST': 0.25}
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def test_brainweb_patients_split_and_slices(tmp_path):
    root, labels, ms, raws = _dataset_dir(tmp_path)
    np.random.seed(0)
    ds = BRAINWEB(_options(root))
    assert [p['name'] for p in ds.patients] == ['t2_n0.mnc.gz', 't2_n1.mnc.gz', 't2_n2.mnc.gz', 't2_n3.mnc.gz', 't2_m0.mnc.gz', 't2_m1.mnc.gz']
    assert [p['type'] for p in ds.patients] == ['NORMAL'] * 4 + ['MILDMS'] * 2
    assert ds.patients[4]['groundtruth_filename'].endswith(os.path.join('groundtruth', 'mild_lesions.mnc.gz'))
    # patients (not slices) are partitioned 3 / 1 / 1 of 6 (floor(0.5*6), max(1, floor(0.25*6)) twice), by NAME
    split = ds.get_patient_split()
    assert [len(split[s]) for s in BRAINWEB.SET_TYPES] == [3, 1, 1]
    assert all(isinstance(n, str) for s in split.values() for n in s)
    names = sum((split[s] for s in BRAINWEB.SET_TYPES), [])
    assert len(set(names)) == 5
    assert sorted(sum((ds.get_patient_idx(s) for s in BRAINWEB.SET_TYPES), [])) == sorted(i for i, p in enumerate(ds.patients) if p['name'] in names)
    with open(ds.split_name(), 'rb') as f:
        assert pickle.load(f) == split
    # slices: Z-1 non-blank axial slices per assigned patient, [N, X, Y, 1] (axial slice of the [x,y,z] array)
    assert ds.images.shape == (5 * (Z - 1), X, Y, 1) and ds.labels.shape == ds.images.shape
    assert ds.images.dtype == np.float32 and ds.num_channels == 1 and (ds.height, ds.width) == (X, Y)
    assert ds.num_examples == 5 * (Z - 1)
    for s, part in enumerate(BRAINWEB.SET_TYPES):
        assert (ds.sets == s).sum() == len(split[part]) * (Z - 1)
        assert ds.get_images(part).shape[0] == len(split[part]) * (Z - 1)
        assert ds.num_batches(4, part) == len(split[part]) * (Z - 1) // 4
    # first assigned patient: slice values = clip at the 99.8th percentile, scale by the max (float32)
    first = next(p for p in ds.patients if p['name'] in names)
    real = raws[first['name']].astype(np.float64) / 255.0 * 255.0
    vol = real.transpose(2, 1, 0).astype(np.float32)
    vol = np.minimum(vol, np.percentile(vol, 99.8))
    vol = vol * (1.0 / vol.max())
    np.testing.assert_allclose(ds.get_image(0)[:, :, 0], vol[:, :, 1], rtol=1e-6)
    # labels are the binary lesion ground truth
    assert set(np.unique(ds.labels)) <= {0.0, 1.0}
    has_ms = any(p['type'] == 'MILDMS' and p['name'] in names for p in ds.patients)
    assert bool(ds.labels.any()) == has_ms
    # a second construction re-uses the stored split
    ds2 = BRAINWEB(_options(root))
    assert ds2.get_patient_split() == split
    np.testing.assert_array_equal(ds2.images, ds.images)


def test_brainweb_volume_preprocessing(tmp_path):
    root, labels, ms, raws = _dataset_dir(tmp_path)
    np.random.seed(1)
    ds = BRAINWEB(_options(root, skullRemoval=True, backgroundRemoval=True))
    patient = ds.patients[4]
    vol, seg, skull = ds.load_volume_and_groundtruth(patient['filtered_files'], patient)
    lab = ms.transpose(2, 1, 0)
    np.testing.assert_array_equal(seg.data, (lab == 10).astype(np.float64))
    brain = ~np.isin(lab, [0, 4, 5, 6, 7, 9])
    np.testing.assert_array_equal(skull.data, brain.astype(np.float64))
    assert not vol.data[~brain].any() and vol.data[brain].any()
    assert vol.data.dtype == np.float32 and abs(float(vol.data.max()) - 1.0) < 1e-6
    assert 'noSkull' in ds.name() and 'noBackground' in ds.name()


def test_brainweb_resolution_pad_resize_crops_and_names(tmp_path):
    root, *_ = _dataset_dir(tmp_path, n_normal=2, n_ms=0)
    np.random.seed(2)
    pad = BRAINWEB(_options(root, filterType=['NORMAL'], sliceResolution=[32, 32]))
    assert pad.images.shape[1:] == (32, 32, 1)
    top, left = (32 - X) // 2, (32 - Y) // 2
    inner = pad.images[0, top:top + X, left:left + Y, 0]
    assert inner.any() and pad.images[0].sum() == pytest.approx(inner.sum())
    assert '_res32x32' in pad.name() and pad.name().startswith('BRAINWEB_p0.5-0.25-0.25')
    small = BRAINWEB(_options(root, filterType=['NORMAL'], sliceResolution=[16, 16]))
    assert small.images.shape[1:] == (16, 16, 1)
    cc = BRAINWEB(_options(root, filterType=['NORMAL'], useCrops=True, cropType='center', cropWidth=8, cropHeight=10))
    assert cc.images.shape[1:] == (10, 8, 1) and 'centercrops8x10' in cc.name()
    rc = BRAINWEB(_options(root, filterType=['NORMAL'], useCrops=True, cropType='random', cropWidth=8, cropHeight=8, numRandomCropsPerSlice=3))
    assert rc.images.shape[0] == 3 * cc.images.shape[0] and rc.images.shape[1:] == (8, 8, 1)
    assert '3cropsPerSlice' in rc.name()
    rot = BRAINWEB(_options(root, filterType=['NORMAL'], rotations=[0, 90]))
    assert rot.images.shape[0] == 2 * cc.images.shape[0]
    few = BRAINWEB(_options(root, filterType=['NORMAL'], numSamples=5))
    assert '_n5' in few.name() and few.num_examples <= 2 * 6


def test_brainweb_cache_is_a_tfrecord(tmp_path):
    root, *_ = _dataset_dir(tmp_path, n_normal=3, n_ms=1)
    np.random.seed(3)
    ds = BRAINWEB(_options(root, cache=True))
    assert os.path.isfile(ds.tfrecord_name()) and ds.tfrecord_name().endswith('.tfrecord')
    again = BRAINWEB(_options(root, cache=True))
    np.testing.assert_array_equal(again.images, ds.images)
    np.testing.assert_array_equal(again.labels, ds.labels)
    np.testing.assert_array_equal(again.sets, ds.sets)
    assert again.get_patient_split() == ds.get_patient_split()


def test_brainweb_next_batch_protocol(tmp_path):
    root, *_ = _dataset_dir(tmp_path, n_normal=4, n_ms=0)
    np.random.seed(4)
    ds = BRAINWEB(_options(root, filterType=['NORMAL'], partition={'TRAIN': 0.5, 'VAL': 0.25, 'TEST': 0.25}))
    n = int((ds.sets == 0).sum())
    train = ds.get_images('TRAIN').copy()
    x, y, m = ds.next_batch(8, set='TRAIN')
    assert x.shape == (8, X, Y, 1) and y.shape == x.shape and m is None and x.dtype == np.float32
    np.testing.assert_array_equal(x, train[:8])                                # first epoch: extraction order (see next_batch)
    seen = [x]
    while ds.epochs_completed['TRAIN'] == 0:
        seen.append(ds.next_batch(8, set='TRAIN')[0])
    # the wrapping batch = tail of the old order + head of the reshuffled set; every slice of the set seen in epoch 0
    total = np.concatenate(seen)[:n]
    np.testing.assert_array_equal(total, train)
    assert seen[-1].shape[0] == 8
    assert ds._index_in_epoch['TRAIN'] == (8 - n % 8) % 8 or n % 8 == 0
    # the reshuffle permutes within the set only
    after = ds.get_images('TRAIN')
    assert sorted(map(float, after.sum(axis=(1, 2, 3)))) == pytest.approx(sorted(map(float, train.sum(axis=(1, 2, 3)))))
    np.testing.assert_array_equal(ds.get_images('VAL'), BRAINWEB(_options(root, filterType=['NORMAL'])).get_images('VAL'))
    # VAL batches without shuffling, with the brain mask derived from the labels
    xv, yv, mv = ds.next_batch(4, shuffle=False, set='VAL', return_brainmask=True)
    assert mv.shape == yv.shape and set(np.unique(mv)) <= {0.0, 1.0}
    noisy = BRAINWEB(_options(root, filterType=['NORMAL'], addInstanceNoise=True))
    xn, _, _ = noisy.next_batch(4, set='TRAIN')
    assert not np.array_equal(xn, noisy.get_images('TRAIN')[:4]) and np.abs(xn - noisy.get_images('TRAIN')[:4]).max() < 0.1


def test_minc1_values_outside_valid_range_are_clamped(tmp_path):
    raw = np.array([[[0, 5, 10, 200, 255]]], np.uint8)                         # valid_range (5, 200)
    p = str(tmp_path / 'c.mnc')
    write_minc1(p, raw, 1.0, 3.0, (5, 200), 'unsigned')
    data, _, _ = read_minc1(p)
    expect = (np.clip(raw.astype(np.float64), 5, 200) - 5) / 195.0 * 2.0 + 1.0
    np.testing.assert_allclose(data, expect.transpose(2, 1, 0), atol=1e-12)


def test_get_datasets_routes_brainweb_directories_to_the_loader(tmp_path):
    """default_config_setup.get_datasets: a BrainWeb tree on disk -> (healthy T2 NORMAL 0.7/0.3/0, SEVEREMS test set), configured
    as the reference does (default_config_setup.py:200-242); no data on disk -> the synthetic generator."""
    from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import SYNTHETIC
    from unsupervised_anomaly_detection_brain_mri_b200.utils import default_config_setup as cfg
    root, labels, ms, _ = _dataset_dir(tmp_path, n_normal=4, n_ms=0)
    rng = np.random.default_rng(11)
    write_minc1(os.path.join(root, 'groundtruth', 'severe_lesions.mnc.gz'), ms, 0.0, 10.0, (0, 10), 'unsigned')
    for i in range(2):
        v = ((ms.astype(np.float64) * 20 + rng.integers(0, 20, ms.shape)) * (ms > 0)).astype(np.uint8)
        write_minc1(os.path.join(root, 'lesions', 'severe', f't2_s{i}.mnc.gz'), v, np.zeros(Z), np.full(Z, 255.0), (0, 255), 'unsigned')
    options = cfg.get_options(batchsize=4, learningrate=1e-4, numEpochs=1, zDim=16, outputWidth=32, outputHeight=32,
                              slices_start=1, slices_end=10, config={'CHECKPOINTDIR': str(tmp_path / 'c'), 'SAMPLEDIR': str(tmp_path / 's'),
                                                                     'BRAINWEBDIR': root})
    options['data']['dir'] = options['globals'][cfg.Dataset.BRAINWEB.value]
    np.random.seed(5)
    hc, pc = cfg.get_datasets(options, cfg.Dataset.BRAINWEB)
    assert isinstance(hc, BRAINWEB) and isinstance(pc, BRAINWEB)
    assert [p['type'] for p in hc.patients] == ['NORMAL'] * 4 and [p['type'] for p in pc.patients] == ['SEVEREMS'] * 2
    assert [len(hc.patients_split[s]) for s in BRAINWEB.SET_TYPES] == [2, 1, 0]         # floor(.7*4), floor(.3*4), 0
    assert [len(pc.patients_split[s]) for s in BRAINWEB.SET_TYPES] == [0, 0, 2]
    assert hc.images.shape == (3 * 9, 32, 32, 1) and pc.images.shape == (2 * 9, 32, 32, 1)
    assert not hc.labels.any() and pc.labels.any()
    assert hc.options.skullRemoval and hc.options.backgroundRemoval and hc.options.cache and os.path.isfile(hc.tfrecord_name())
    assert float(hc.images.min()) == 0.0 and float(hc.images.max()) <= 1.0 + 1e-6
    assert len(pc.get_patient_idx('TEST')) == 2
    vol, seg, skull = pc.load_volume_and_groundtruth(pc.patients[0]['filtered_files'], pc.patients[0])
    assert vol.num_slices_along_axis(pc.options.axis) == Z and seg.data.any()
    options['data']['dir'] = str(tmp_path / 'nothing-here')
    hs, ps = cfg.get_datasets(options, cfg.Dataset.BRAINWEB)
    assert isinstance(hs, SYNTHETIC) and isinstance(ps, SYNTHETIC)


def test_export_patient_volume_writes_nifti_in_native_geometry(tmp_path):
    """options['exportVolumes'] (reference utils/Evaluation.py:323-334): the [slices, H, W] residual sub-volume is resampled back to
    the native in-plane size and written into the patient's volume geometry, plus a binary volume for a numeric threshold."""
    import types

    from unsupervised_anomaly_detection_brain_mri_b200.utils import Evaluation as Ev
    NII.set_view_mapping({'saggital': 2, 'coronal': 1, 'axial': 0})          # [z, y, x] volumes (the MSLUB-style mapping)
    try:
        vol = NII(data=np.zeros((12, 20, 24)))
        vol.origin, vol.spacing = (1.0, 2.0, 3.0), (1.0, 1.0, 2.0)
        sub = np.zeros((6, 10, 12), np.float64)                              # slices 3..8 at half resolution
        sub[2, 4:6, 5:8] = 0.5
        opts = types.SimpleNamespace(sliceStart=3, sliceEnd=9, axis='axial')
        paths = Ev.export_patient_volume(vol, sub, (0.5, 0.5), opts, {'threshold': 0.25}, str(tmp_path), 'patient7')
        assert [os.path.basename(p) for p in paths] == ['patient7.nii.gz', 'patient7.binary.nii.gz']
        cont, binary = NII(paths[0]), NII(paths[1])
        assert cont.data.shape == (12, 20, 24) and cont.origin == (1.0, 2.0, 3.0)
        assert not cont.data[:3].any() and not cont.data[9:].any() and cont.data[5].max() > 0.3
        assert set(np.unique(binary.data)) == {0.0, 1.0} and binary.data[5, 8:12, 10:16].all()
        only = Ev.export_patient_volume(NII(data=np.zeros((12, 20, 24))), sub, (0.5, 0.5), opts, {'threshold': 'bestdice'}, str(tmp_path), 'q')
        assert len(only) == 1
    finally:
        NII.set_view_mapping({'saggital': 0, 'coronal': 1, 'axial': 2})


# ---------------------------------------------------------------------------------------------------- NIfTI lesion datasets
def test_curvature_flow_is_mean_curvature_motion():
    """NII.denoise restates SimpleITK's CurvatureFlow (I_t = kappa |grad I|): on a radially symmetric bump f(r) the exact rate is
    (n - 1) f'(r) / r; flat regions and constants are fixed points; spacing rescales the derivatives."""
    from unsupervised_anomaly_detection_brain_mri_b200.utils.NII import curvature_flow
    n = 97
    y, x = np.mgrid[0:n, 0:n] - 48.0
    r = np.hypot(x, y)
    img = np.exp(-r ** 2 / (2 * 12.0 ** 2))
    out = curvature_flow(img, None, 0.125, 1)
    rate = -r / 144.0 * img / np.maximum(r, 1e-9)
    ring = (r > 5) & (r < 30)
    assert np.abs((out - img)[ring] / 0.125 - rate[ring]).max() < 0.01 * np.abs(rate[ring]).max()
    z, y, x = np.mgrid[0:41, 0:41, 0:41] - 20.0
    r3 = np.sqrt(x * x + y * y + z * z)
    vol = np.exp(-r3 ** 2 / (2 * 7.0 ** 2))
    out3 = curvature_flow(vol, None, 0.125, 1)
    rate3 = 2 * (-r3 / 49.0 * vol) / np.maximum(r3, 1e-9)
    shell = (r3 > 3) & (r3 < 14)
    assert np.abs((out3 - vol)[shell] / 0.125 - rate3[shell]).max() < 0.02 * np.abs(rate3[shell]).max()
    assert np.array_equal(curvature_flow(np.full((5, 6, 7), 3.0)), np.full((5, 6, 7), 3.0))
    ramp = np.tile(np.arange(16.0), (16, 1))                                     # straight level lines: zero curvature
    assert np.abs(curvature_flow(ramp, None, 0.125, 3) - ramp).max() < 1e-12
    half = curvature_flow(img, (2.0, 2.0), 0.125, 1)                              # doubled spacing: every derivative term scales by 1/4
    assert np.abs((half - img)[ring] * 4 - (out - img)[ring]).max() < 1e-12
    v = NII(data=vol)
    v.denoise()
    assert v.data.shape == vol.shape and 0 < np.abs(v.data - vol).max() < 0.05


def _lesion_tree(tmp_path, kind):
    rng = np.random.default_rng(5)
    Zs, Ys, Xs = 10, 20, 24                                                      # [k, j, i] = axial slices first
    names = []

    def volume():
        zz, yy, xx = np.mgrid[0:Zs, 0:Ys, 0:Xs]
        brain = (((yy - Ys / 2) / (Ys / 2 - 1)) ** 2 + ((xx - Xs / 2) / (Xs / 2 - 1)) ** 2) < 1
        brain[0] = False
        img = (100 + 40 * rng.random((Zs, Ys, Xs))) * brain + 5 * rng.random((Zs, Ys, Xs))
        gt = np.zeros((Zs, Ys, Xs), np.float32)
        gt[4:6, 8:12, 10:14] = 1
        return img.astype(np.float32), gt, brain.astype(np.float32)

    root = tmp_path / kind
    for p in range(4):
        img, gt, brain = volume()
        if kind == 'MSLUB':
            d = root / 'data' / f'patient{p:02d}'
            files = {f'patient{p:02d}_FLAIR.aligned.nii.gz': img, f'patient{p:02d}_T1W.aligned.nii.gz': img * 0.5,
                     f'patient{p:02d}_T1WKS.aligned.nii.gz': img, f'patient{p:02d}_T2W.aligned.nii.gz': img,
                     f'patient{p:02d}_consensus_gt.aligned.nii.gz': gt, f'patient{p:02d}_brainmask.aligned.nii.gz': brain}
        elif kind == 'MSISBI2015':
            folder = f'training0{p + 1}'
            d = root / folder / 'preprocessed'
            nm = f'{folder}_01'
            files = {f'{nm}_flair_pp.nii': img, f'{nm}_flair.aligned.nii.gz': img, f'{nm}_mask1.aligned.nii.gz': gt,
                     f'{nm}_skullmap.aligned.nii.gz': brain}
        else:
            d = root / ('UNC_train' if p < 3 else 'CHB_train') / f'case{p:02d}'
            nm = f'case{p:02d}'
            files = {f'{nm}_FLAIR.aligned.nii.gz': img, f'{nm}_T1.aligned.nii.gz': img, f'{nm}_T2.aligned.nii.gz': img,
                     f'{nm}_lesion.aligned.nii.gz': gt, f'{nm}_skullmap.nii.gz': brain}
        os.makedirs(d)
        for fn, arr in files.items():
            write_nifti(str(d / fn), arr)
        names.append(d)
    return str(root)


@pytest.mark.parametrize('kind', ['MSLUB', 'MSISBI2015', 'MSSEG2008'])
def test_lesion_datasets_through_get_datasets(tmp_path, kind):
    from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import SYNTHETIC
    from unsupervised_anomaly_detection_brain_mri_b200.utils import default_config_setup as cfg
    root = _lesion_tree(tmp_path, kind)
    globals_ = {'CHECKPOINTDIR': str(tmp_path / 'c'), 'SAMPLEDIR': str(tmp_path / 's'), 'MSLUBDIR': '', 'MSISBI2015DIR': '', 'MSSEG2008DIR': '',
                'BRAINWEBDIR': ''}
    globals_[kind + 'DIR'] = root
    options = cfg.get_options(batchsize=2, learningrate=1e-4, numEpochs=1, zDim=16, outputWidth=32, outputHeight=32, slices_start=1,
                              slices_end=9, config=globals_)
    member = {'MSLUB': cfg.Dataset.MSLUB, 'MSISBI2015': cfg.Dataset.MSISBI2015, 'MSSEG2008': cfg.Dataset.MSSEG2008_UNC}[kind]
    np.random.seed(3)
    try:
        hc, pc = cfg.get_datasets(options, member)
        assert hc is None and type(pc).__name__ == kind
        n_pat = 3 if kind == 'MSSEG2008' else 4
        assert len(pc.patients) == n_pat and all(len(p['filtered_files']) == 1 for p in pc.patients)       # FLAIR only
        val, test = list(pc.get_patient_idx('VAL')), list(pc.get_patient_idx('TEST'))
        assert len(pc.get_patient_idx('TRAIN')) == 0 and len(val) + len(test) == n_pat and not set(val) & set(test)
        assert pc.images.shape[1:] == (32, 32, 1) and pc.images.dtype == np.float32 and pc.num_channels == 1
        assert pc.images.shape[0] == n_pat * 8 and set(np.unique(pc.labels)) <= {0.0, 1.0} and pc.labels.any()
        assert float(pc.images.max()) <= 1.0 + 1e-6 and float(pc.images.min()) == 0.0          # skull-stripped, scaled
        assert os.path.isfile(pc.tfrecord_name()) and pc.name().endswith('_res32x32_aligned')
        again = cfg.get_datasets(options, member)[1]                                  # second construction: cached TFRecord + stored split
        np.testing.assert_array_equal(again.images, pc.images)
        assert [list(again.get_patient_idx(s)) for s in pc.SET_TYPES] == [list(pc.get_patient_idx(s)) for s in pc.SET_TYPES]
        x, y, m = pc.next_batch(4, set='TEST' if len(test) else 'VAL')                # (shuffles the set in place on its first call)
        assert x.shape == (4, 32, 32, 1) and m.dtype == bool and m.shape == x.shape
        # the evaluation protocol (utils/Evaluation._evaluate): axial slices of the [z, y, x] volume, binary ground truth, brain mask
        patient = pc.patients[int(pc.get_patient_idx('TEST' if len(test) else 'VAL')[0])]
        vol, seg, skull = pc.load_volume_and_groundtruth(patient['filtered_files'][0], patient)
        assert vol.num_slices_along_axis('axial') == 10 and vol.get_slice(3, 'axial').shape == (20, 24)
        assert set(np.unique(seg.data)) == {0.0, 1.0} and skull is not None and not vol.data[0].any()
        options['globals'][kind + 'DIR'] = str(tmp_path / 'absent')
        assert isinstance(cfg.get_datasets(options, member)[1], SYNTHETIC)
    finally:
        NII.set_view_mapping({'saggital': 0, 'coronal': 1, 'axial': 2})
