"""Synthetic BrainWeb-shaped dataset with the protocol the hot path consumes (reference dataloaders/BRAINWEB.py:
num_batches :406-409, next_batch :411-478, num_channels :351-352, patients / get_patient_idx / load_volume_and_groundtruth).

The real loaders (MINC / NIfTI I/O) are out of scope; the metric is defined on synthetic slices:
float32 NHWC in [0,1], ~50 % exact-zero background outside an elliptical 'brain', blurred-noise foreground."""
import math

import numpy as np


class Volume:
    """Minimal stand-in for utils/NII.py's NII wrapper: data [Z,H,W]."""

    def __init__(self, data):
        self.data = data

    def shape(self):
        return self.data.shape

    def num_slices_along_axis(self, axis):
        return self.data.shape[0]

    def get_slice(self, s, axis):
        return self.data[s]


def _blur(img, sigma):
    from scipy.ndimage import gaussian_filter
    return gaussian_filter(img, sigma=sigma)


def make_volume(S, Z, seed, lesions=True):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:S, 0:S].astype(np.float32)
    c = (S - 1) / 2.0
    vol = np.zeros((Z, S, S), np.float32)
    seg = np.zeros((Z, S, S), np.uint8)
    for k in range(Z):
        r = 0.75 + 0.25 * math.sin(math.pi * (k + 0.5) / Z)
        mask = ((yy - c) / (0.42 * S * r)) ** 2 + ((xx - c) / (0.36 * S * r)) ** 2 <= 1.0
        g = _blur(rng.uniform(size=(S, S)).astype(np.float32), S / 16.0)
        g = (g - g.min()) / max(float(g.max() - g.min()), 1e-12)
        img = np.clip(0.15 + 0.55 * g + 0.05 * rng.standard_normal((S, S)).astype(np.float32), 0, 1).astype(np.float32)
        if lesions:
            ys, xs = np.nonzero(mask)
            for _ in range(int(rng.integers(0, 3))):
                j = int(rng.integers(0, ys.size))
                rad = float(rng.integers(3, 9)) * S / 256.0 + 1
                disc = ((yy - ys[j]) ** 2 + (xx - xs[j]) ** 2 <= rad * rad) & mask
                img = np.where(disc, np.minimum(img + 0.3, 1.0), img).astype(np.float32)
                seg[k][disc] = 1
        vol[k] = np.where(mask, img, 0.0)
    return vol, seg, (vol > 0).astype(np.uint8)


class SYNTHETIC:
    class Options:
        def __init__(self):
            self.sliceResolution = [128, 128]
            self.numPatients = 4
            self.sliceStart = 20
            self.sliceEnd = 130
            self.axis = 'axial'
            self.partition = {'TRAIN': 0.7, 'VAL': 0.3, 'TEST': 0.0}
            self.lesions = False
            self.seed = 1234
            self.description = ''

    def __init__(self, options=None):
        self.options = options or self.Options()
        o = self.options
        S = o.sliceResolution[0]
        Z = o.sliceEnd - o.sliceStart
        self.patients = []
        self._volumes = {}
        for p in range(o.numPatients):
            name = f'synthetic_{p:03d}'
            self.patients.append({'name': name, 'filtered_files': name, 'idx': p})
            self._volumes[name] = make_volume(S, Z, o.seed + p, lesions=o.lesions)
        imgs = np.concatenate([self._volumes[p['name']][0] for p in self.patients], 0)[..., None]      # [N,H,W,1] f32 in RAM
        masks = np.concatenate([self._volumes[p['name']][2] for p in self.patients], 0)[..., None]
        labels = np.concatenate([self._volumes[p['name']][1] for p in self.patients], 0)[..., None]
        n = imgs.shape[0]
        n_train = int(round(n * o.partition.get('TRAIN', 0.0)))
        n_val = int(round(n * o.partition.get('VAL', 0.0)))
        bounds = {'TRAIN': (0, n_train), 'VAL': (n_train, n_train + n_val), 'TEST': (n_train + n_val, n)}
        self._sets = {k: (imgs[a:b], labels[a:b], masks[a:b]) for k, (a, b) in bounds.items()}
        self._cursor = {k: 0 for k in bounds}
        self._order = {k: np.arange(b - a) for k, (a, b) in bounds.items()}
        self._rng = np.random.default_rng(o.seed)

    @property
    def num_channels(self):
        return 1

    def get_patient_idx(self, split='TEST'):
        n = len(self.patients)
        n_train = int(round(n * self.options.partition.get('TRAIN', 0.0)))
        n_val = int(round(n * self.options.partition.get('VAL', 0.0)))
        return {'TRAIN': list(range(0, n_train)), 'VAL': list(range(n_train, n_train + n_val)),
                'TEST': list(range(n_train + n_val, n))}[split]

    def load_volume_and_groundtruth(self, filename, patient=None):
        vol, seg, skull = self._volumes[filename]
        pad = self.options.sliceStart      # present the sub-volume at its original slice indices
        def padz(a):
            return np.concatenate([np.zeros((pad,) + a.shape[1:], a.dtype), a], 0)
        return Volume(padz(vol)), Volume(padz(seg)), Volume(padz(skull))

    def num_batches(self, batchsize, set='TRAIN'):
        return self._sets[set][0].shape[0] // batchsize

    def next_batch(self, batchsize, shuffle=True, set='TRAIN', return_brainmask=False):
        imgs, labels, masks = self._sets[set]
        n = imgs.shape[0]
        cur = self._cursor[set]
        if cur + batchsize > n:                       # wrap-around: reshuffle (BRAINWEB.py:427-442)
            if shuffle:
                self._order[set] = self._rng.permutation(n)
            cur = 0
        idx = self._order[set][cur:cur + batchsize]
        self._cursor[set] = cur + batchsize
        batch = imgs[idx]
        assert batch.size
        return batch, labels[idx], (masks[idx] if return_brainmask else None)
