"""MINC volumes with the interface of reference utils/MINC.py (class MINC(NII)) - without nibabel.

The reference converts every .mnc(.gz) to NIfTI through nibabel and re-reads it with SimpleITK (MINC.py:11-25).  BrainWeb
ships MINC-1 files, i.e. NetCDF *classic* containers, which scipy.io.netcdf_file reads:
  variable 'image' [zspace, yspace, xspace] (byte / short, 'signtype' attribute, 'valid_range' = (vmin, vmax)),
  variables 'image-min' / 'image-max' (scalar or per slice): real = (v - vmin) / (vmax - vmin) * (imax - imin) + imin
  (the MINC-1 real-value rule nibabel's Minc1File applies), dimension variables 'xspace' ... with 'start' / 'step'.
After the NIfTI round trip SimpleITK indexes the array [x, y, z] (MINC axis names), so that is the layout of ``data`` here
and VIEW_MAPPING {'saggital': 0, 'coronal': 1, 'axial': 2} selects xspace / yspace / zspace.
MINC-2 (HDF5) files are not supported (no HDF5 reader in this environment) and raise."""
import gzip
import io

import numpy as np

from .NII import NII


def read_minc1(filename):
    """-> (data [x,y,z] float64 real values, origin (x,y,z), spacing (x,y,z))"""
    from scipy.io import netcdf_file
    opener = gzip.open if str(filename).endswith('.gz') else open
    with opener(filename, 'rb') as f:
        raw = f.read()
    if raw[:3] != b'CDF':
        if raw[:4] == b'\x89HDF':
            raise IOError(f'{filename}: MINC-2 (HDF5) files are not supported - convert with `mincconvert -1`')
        raise IOError(f'{filename}: not a MINC-1 (NetCDF classic) file')
    nc = netcdf_file(io.BytesIO(raw), 'r', mmap=False)
    img = nc.variables['image']
    dims = tuple(img.dimensions)
    if sorted(dims) != ['xspace', 'yspace', 'zspace']:
        raise IOError(f'{filename}: unsupported image dimensions {dims}')
    v = np.array(img[:])
    sign = getattr(img, 'signtype', b'signed__')
    sign = sign.decode() if isinstance(sign, bytes) else str(sign)
    if sign.startswith('unsigned') and v.dtype.kind == 'i':
        v = v.view(np.dtype(f'u{v.dtype.itemsize}').newbyteorder(v.dtype.byteorder))
    v = v.astype(np.float64)
    if hasattr(img, 'valid_range'):
        vmin, vmax = [float(t) for t in np.array(img.valid_range).reshape(-1)[:2]]
    else:
        info = np.iinfo(np.array(img[:]).dtype) if np.array(img[:]).dtype.kind in 'iu' else None
        vmin, vmax = (float(info.min), float(info.max)) if info else (float(v.min()), float(v.max()))
    integer_voxels = np.array(img[:]).dtype.kind in 'iu'        # float voxels ARE the real values (no rescaling)
    if integer_voxels and 'image-min' in nc.variables and 'image-max' in nc.variables and vmax > vmin:
        v = np.clip(v, vmin, vmax)                               # values outside valid_range are clamped first
        imin = np.array(nc.variables['image-min'][...] if nc.variables['image-min'].shape else nc.variables['image-min'].getValue(), np.float64)
        imax = np.array(nc.variables['image-max'][...] if nc.variables['image-max'].shape else nc.variables['image-max'].getValue(), np.float64)
        extra = (1,) * (v.ndim - imin.ndim)             # per-slice (or per-row) scalars broadcast over the trailing image dims
        imin, imax = imin.reshape(imin.shape + extra), imax.reshape(imax.shape + extra)
        v = (v - vmin) / (vmax - vmin) * (imax - imin) + imin
    order = [dims.index(n) for n in ('xspace', 'yspace', 'zspace')]
    data = np.ascontiguousarray(v.transpose(order))
    origin, spacing = [], []
    for n in ('xspace', 'yspace', 'zspace'):
        d = nc.variables.get(n)
        origin.append(float(getattr(d, 'start', 0.0)) if d is not None else 0.0)
        spacing.append(float(getattr(d, 'step', 1.0)) if d is not None else 1.0)
    return data, tuple(origin), tuple(spacing)


class MINC(NII):
    def __init__(self, filename):
        if str(filename).endswith(('.nii', '.nii.gz')):
            NII.__init__(self, filename)
            return
        data, origin, spacing = read_minc1(filename)
        NII.__init__(self, data=data)
        self.origin, self.spacing = origin, spacing
