"""Volume wrapper with the interface of reference utils/NII.py (class NII: .data, VIEW_MAPPING, num_*_slices,
num_slices_along_axis, normalize, apply_skullmap, subtract, get_slice / set_slice / set_subvolume, get_data, set_to_zero,
copy, save) - without SimpleITK.  NIfTI-1 files (.nii / .nii.gz) are read and written with numpy only.

Array convention: ``data`` is indexed [k, j, i] for NIfTI voxel (i, j, k) - what ``sitk.GetArrayFromImage`` returns in the
reference (NII.py:23) - so VIEW_MAPPING {'saggital': 0, 'coronal': 1, 'axial': 2} means the same here.

NIfTI-1 header fields used (nifti1.h): sizeof_hdr@0 (348; byte-swapped 348 => big-endian file), dim@40 (8 x int16),
datatype@70, bitpix@72, pixdim@76 (8 x float32), vox_offset@108, scl_slope@112, scl_inter@116, qoffset_x/y/z@268,
srow_x/y/z@280, magic@344 ('n+1\\0' single file).  denoise() restates SimpleITK's CurvatureFlow (see curvature_flow); visualize() (matplotlib) is not reproduced."""
import copy
import gzip
import struct

import numpy as np

_NIFTI_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
                 768: np.uint32, 1024: np.int64, 1280: np.uint64}
_NIFTI_CODES = {np.dtype(v): k for k, v in _NIFTI_DTYPES.items()}


def _open(filename, mode='rb'):
    return gzip.open(filename, mode) if str(filename).endswith('.gz') else open(filename, mode)


def read_nifti(filename):
    """-> (data [k,j,i] float64 with scl_slope/scl_inter applied, origin (x,y,z), spacing (i,j,k))"""
    with _open(filename) as f:
        raw = f.read()
    if len(raw) < 352:
        raise IOError(f'{filename}: too short for a NIfTI-1 file')
    end = '<'
    if struct.unpack('<i', raw[:4])[0] != 348:
        if struct.unpack('>i', raw[:4])[0] != 348:
            raise IOError(f'{filename}: not a NIfTI-1 file (sizeof_hdr != 348)')
        end = '>'
    if raw[344:347] not in (b'n+1', b'ni1'):
        raise IOError(f'{filename}: bad NIfTI magic {raw[344:348]!r}')
    if raw[344:347] == b'ni1':
        raise IOError(f'{filename}: two-file NIfTI (.hdr/.img) is not supported')
    dim = struct.unpack(end + '8h', raw[40:56])
    datatype, = struct.unpack(end + 'h', raw[70:72])
    pixdim = struct.unpack(end + '8f', raw[76:108])
    vox_offset, slope, inter = struct.unpack(end + '3f', raw[108:120])
    if datatype not in _NIFTI_DTYPES:
        raise IOError(f'{filename}: unsupported NIfTI datatype {datatype}')
    nd = dim[0]
    shape = tuple(int(d) for d in dim[1:1 + nd])
    while len(shape) > 3 and shape[-1] == 1:
        shape = shape[:-1]
    dt = np.dtype(_NIFTI_DTYPES[datatype]).newbyteorder(end)
    n = int(np.prod(shape))
    arr = np.frombuffer(raw, dtype=dt, count=n, offset=int(vox_offset)).reshape(shape[::-1]).astype(np.float64)
    if slope not in (0.0,) and not np.isnan(slope):
        arr = arr * slope + (0.0 if np.isnan(inter) else inter)
    origin = struct.unpack(end + '3f', raw[268:280])
    return arr, origin, tuple(pixdim[1:4])


def write_nifti(filename, data, origin=(0.0, 0.0, 0.0), spacing=(1.0, 1.0, 1.0)):
    """data [k,j,i] -> single-file NIfTI-1 (little endian), dtype preserved when NIfTI has a code for it, else float32."""
    a = np.asarray(data)
    if a.dtype not in _NIFTI_CODES:
        a = a.astype(np.float32)
    a = np.ascontiguousarray(a.astype(a.dtype.newbyteorder('<')))
    hdr = bytearray(348)
    struct.pack_into('<i', hdr, 0, 348)
    shape = a.shape[::-1]
    struct.pack_into('<8h', hdr, 40, len(shape), *(list(shape) + [1] * (7 - len(shape))))
    struct.pack_into('<h', hdr, 70, _NIFTI_CODES[np.dtype(a.dtype.name)])
    struct.pack_into('<h', hdr, 72, a.dtype.itemsize * 8)
    struct.pack_into('<8f', hdr, 76, 1.0, *(list(spacing) + [1.0] * 4))
    struct.pack_into('<3f', hdr, 108, 352.0, 1.0, 0.0)
    struct.pack_into('<h', hdr, 252, 0)                                  # qform_code
    struct.pack_into('<h', hdr, 254, 1)                                  # sform_code = scanner
    struct.pack_into('<3f', hdr, 268, *origin)
    for r, (off, s) in enumerate(zip((280, 296, 312), spacing)):
        row = [0.0, 0.0, 0.0, float(origin[r])]
        row[r] = float(s)
        struct.pack_into('<4f', hdr, off, *row)
    hdr[344:348] = b'n+1\x00'
    with _open(filename, 'wb') as f:
        f.write(bytes(hdr) + b'\x00' * 4 + a.tobytes())


def curvature_flow(image, spacing=None, time_step=0.125, iterations=3):
    """itk::CurvatureFlowImageFilter (N-D): per iteration, for every voxel
        update = [ sum_i Ixx_i * sum_{j != i} Ix_j^2  -  2 sum_{i<j} Ix_i Ix_j Ixy_ij ] / |grad I|^2     (0 where |grad I|^2 < 1e-9)
        I     <- I + time_step * update
    first derivatives 0.5 (I[+1] - I[-1]) / h, second (I[+1] - 2 I + I[-1]) / h^2, cross 0.25 (I[--] - I[-+] - I[+-] + I[++]) / (h_i h_j);
    neighbours outside the image replicate the edge voxel (ZeroFluxNeumannBoundaryCondition)."""
    img = np.array(image, dtype=np.float64, copy=True)
    nd = img.ndim
    h = [1.0] * nd if spacing is None else [float(s) for s in spacing]

    def shifted(p, offsets):
        sl = tuple(slice(1 + o, p.shape[a] - 1 + o) for a, o in enumerate(offsets))
        return p[sl]

    for _ in range(int(iterations)):
        p = np.pad(img, 1, mode='edge')
        zero = [0] * nd
        first, second = [], []
        for i in range(nd):
            plus, minus = list(zero), list(zero)
            plus[i], minus[i] = 1, -1
            first.append(0.5 * (shifted(p, plus) - shifted(p, minus)) / h[i])
            second.append((shifted(p, plus) - 2.0 * img + shifted(p, minus)) / (h[i] * h[i]))
        mag = sum(f * f for f in first)
        update = np.zeros_like(img)
        for i in range(nd):
            update += second[i] * (mag - first[i] * first[i])
        for i in range(nd):
            for j in range(i + 1, nd):
                o = [list(zero) for _ in range(4)]
                o[0][i], o[0][j] = -1, -1
                o[1][i], o[1][j] = -1, 1
                o[2][i], o[2][j] = 1, -1
                o[3][i], o[3][j] = 1, 1
                cross = 0.25 * (shifted(p, o[0]) - shifted(p, o[1]) - shifted(p, o[2]) + shifted(p, o[3])) / (h[i] * h[j])
                update -= 2.0 * first[i] * first[j] * cross
        ok = mag >= 1e-9
        update = np.where(ok, update / np.where(ok, mag, 1.0), 0.0)
        img = img + time_step * update
    return img


class NII:
    VIEW_MAPPING = {'saggital': 0, 'coronal': 1, 'axial': 2}

    def __init__(self, filename=None, data=None):
        self.origin, self.spacing = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
        if filename is not None:
            self.data, self.origin, self.spacing = read_nifti(filename)
        else:
            self.data = np.array(data, dtype=np.float64)
        self.data[np.isnan(self.data)] = 0          # reference NII.py:15

    def save(self, filename):
        write_nifti(filename, self.data, self.origin, self.spacing)

    @property
    def num_saggital_slices(self):
        return self.data.shape[NII.VIEW_MAPPING['saggital']]

    @property
    def num_coronal_slices(self):
        return self.data.shape[NII.VIEW_MAPPING['coronal']]

    @property
    def num_axial_slices(self):
        return self.data.shape[NII.VIEW_MAPPING['axial']]

    @staticmethod
    def set_view_mapping(mapping):
        NII.VIEW_MAPPING = mapping

    def shape(self):
        return self.data.shape

    def num_slices_along_axis(self, axis):
        return self.data.shape[NII.VIEW_MAPPING[axis]]

    def normalize(self, method='scaling', lowerpercentile=None, upperpercentile=None):
        """reference NII.py:52-74: optional percentile clipping, then max-scaling or standardisation (float32)."""
        d = self.data.astype(np.float32)
        lo = np.percentile(d, lowerpercentile) if lowerpercentile is not None else None
        hi = np.percentile(d, upperpercentile) if upperpercentile is not None else None
        if lo is not None:
            d[d < lo] = lo
        if hi is not None:
            d[d > hi] = hi
        if method == 'scaling':
            if d.max() > 0.0:
                d = np.multiply(d, 1.0 / d.max())
        elif method == 'standardization':
            d = d - np.mean(d)
            d = d / np.std(d)
        self.data = d

    def apply_skullmap(self, skullmap):
        m = skullmap.get_data()
        m[m < 0.1] = 0              # in place on the skull map, as the reference does (NII.py:77-80)
        m[m >= 0.1] = 1
        self.data = self.data * m

    def subtract(self, filename):
        self.data = self.data - NII(filename).get_data()

    def _index(self, the_slice, axis):
        idx = [slice(None)] * self.data.ndim
        idx[NII.VIEW_MAPPING[axis]] = the_slice
        return tuple(idx)

    def get_slice(self, the_slice, axis='axial'):
        return self.data[self._index(the_slice, axis)]

    def set_slice(self, the_slice, the_data, axis='axial'):
        self.data[self._index(the_slice, axis)] = the_data

    def set_subvolume(self, slice_start, slice_end, subvolume, axis='axial'):
        for s in range(slice_start, slice_end):     # the first index of the sub-volume is the axis iterated over
            self.set_slice(s, subvolume[s - slice_start, :, :], axis)

    def get_data(self):
        return self.data

    def cast_to_float(self):
        self.data = self.data.astype(np.float64)

    def set_to_zero(self):
        self.data.fill(0.0)

    def denoise(self, time_step=0.125, iterations=3):
        """reference NII.py:82-84: sitk.CurvatureFlow(timeStep=0.125, numberOfIterations=3) - restated, not linked (SimpleITK 1.2.x is
        a third-party dependency that is not installable here; this follows ITK's published itkCurvatureFlowFunction: explicit Euler
        steps of  I_t = kappa |grad I|  with central differences scaled by 1 / spacing, zero-flux (edge-replicating) boundaries and
        a zero update where |grad I|^2 < 1e-9).  NOT validated against ITK itself (none available): stated in DESIGN.md."""
        self.data = curvature_flow(np.asarray(self.data, np.float64), self.spacing_of_axes(), time_step, iterations)

    def spacing_of_axes(self):
        """Voxel spacing per ARRAY axis: data is [k, j, i] for NIfTI voxel (i, j, k) whose pixdim is (di, dj, dk)."""
        sp = tuple(float(v) if v else 1.0 for v in self.spacing)
        return sp[::-1] if self.data.ndim == len(sp) else (1.0,) * self.data.ndim

    def copy(self):
        return copy.deepcopy(self)
