"""Shared machinery of the three NIfTI lesion datasets (reference dataloaders/MSLUB.py, MSISBI2015.py, MSSEG2008.py - there three
near-identical 480-line files; the patient discovery differs, everything else is common and lives here):
per-patient TRAIN / VAL / TEST split stored as `split-*.pckl` (patient INDICES, as these three loaders write it), per volume
denoising (NII.denoise), skull stripping with the patient's brain mask, 0 / 99.8 percentile clip + normalisation, then per axial
slice: skip slices whose 90th percentile is below 0.2, zero-pad small slices, zoom to sliceResolution (labels re-binarised at 0.9),
optional crops (random / center / around lesions), TFRecord cache, and the `next_batch` epoch logic (these loaders DO shuffle at
the start of the first epoch; the brain mask of a batch is `images > 0.05`)."""
import math
import os
import pickle

import numpy
import scipy.ndimage

from ..utils.image_utils import crop, crop_center
from ..utils.NII import NII
from ..utils.tfrecord_utils import read_tf_record, write_tf_record


class LesionDataset(object):
    NAME = 'DATASET'
    SET_TYPES = ['TRAIN', 'VAL', 'TEST']

    class Options(object):
        def __init__(self):
            self.dir = os.path.dirname(os.path.realpath(__file__))
            self.numSamples = -1
            self.partition = {'TRAIN': 0.7, 'VAL': 0.2, 'TEST': 0.1}
            self.useCrops = False
            self.cropType = 'random'             # random, center or lesions
            self.numRandomCropsPerSlice = 5
            self.onlyPatchesWithLesions = False
            self.rotations = 0
            self.cropWidth = 128
            self.cropHeight = 128
            self.cache = False
            self.sliceResolution = None          # HxW
            self.addInstanceNoise = False
            self.filterProtocol = None
            self.filterProtocols = []
            self.filterType = 'train'
            self.axis = 'axial'
            self.debug = False
            self.normalizationMethod = 'standardization'
            self.sliceStart = 0
            self.sliceEnd = 155
            self.format = 'raw'                  # raw or aligned
            self.skullStripping = True
            self.viewMapping = {'saggital': 2, 'coronal': 1, 'axial': 0}

    def __init__(self, options=None):
        self.options = options if options is not None else self.Options()
        o = self.options
        self.patients = self._get_patients()
        self._epochs_completed = {s: 0 for s in self.SET_TYPES}
        self._index_in_epoch = {s: 0 for s in self.SET_TYPES}
        self.patientsSplit = self._load_or_make_split()
        if o.cache and os.path.isfile(self.tfrecord_name()):
            self._images, self._labels, self._sets = read_tf_record(self.tfrecord_name())
            self._sets = self._sets.reshape(-1)
        else:
            self._create_numpy_arrays()
            if o.cache:
                write_tf_record(self._images, self._labels, self._sets, self.tfrecord_name())

    # ------------------------------------------------------------------ patients and their split
    def _get_patients(self):
        return type(self).get_patients(self.options)

    def _load_or_make_split(self):
        if os.path.isfile(self.split_name()):
            with open(self.split_name(), 'rb') as f:
                return pickle.load(f)
        n = len(self.patients)
        order = numpy.random.permutation(n)
        split, taken = {}, 0
        for part, share in self.options.partition.items():
            count = math.floor(share * n) if share <= 1.0 else int(share)      # (a share of exactly 1 means "all", as in the reference)
            count = min(count, n - taken)
            split[part] = order[taken:taken + count]
            taken += count
        os.makedirs(self.dir(), exist_ok=True)
        with open(self.split_name(), 'wb') as f:
            pickle.dump(split, f)
        return split

    def get_patient_idx(self, split='TRAIN'):
        return self.patientsSplit[split]

    def get_patient_split(self):
        return self.patientsSplit

    def get_patient(self, i):
        return self.patients[i]

    # ------------------------------------------------------------------ volumes -> slices
    def load_volume_and_groundtruth(self, nii_filename, patient):
        o = self.options
        nii = self._open(nii_filename)
        nii_groundtruth = self._open(patient['groundtruth'])
        nii.denoise()
        nii.set_view_mapping(o.viewMapping)
        nii.data[numpy.isnan(nii.data)] = 0.0
        gt = nii_groundtruth.data
        nii_groundtruth.data = (gt >= 0.9).astype(gt.dtype)                    # binary ground truth
        nii_skullmap = None
        if o.skullStripping:
            try:
                nii_skullmap = self._open(patient['skullmap'])
                nii_skullmap.set_view_mapping(o.viewMapping)
                nii.apply_skullmap(nii_skullmap)
            except (IOError, OSError):
                print(f'{self.NAME}: Failed to open file ' + patient['skullmap'] + ', skipping skullremoval')
        nii.normalize(method=o.normalizationMethod, lowerpercentile=0, upperpercentile=99.8)
        return nii, nii_groundtruth, nii_skullmap

    def _open(self, filename):
        return NII(filename)

    def gather_data(self, patient, nii_filename):
        o = self.options
        images, labels = [], []
        nii, nii_seg, _ = self.load_volume_and_groundtruth(nii_filename, patient)
        for s in range(o.sliceStart, min(o.sliceEnd, nii.num_slices_along_axis(o.axis))):
            if 0 < o.numSamples < len(images):
                break
            img, seg = nii.get_slice(s, o.axis), nii_seg.get_slice(s, o.axis)
            if numpy.percentile(img, 90) < 0.2:                               # "empty" slice
                continue
            if o.sliceResolution is not None:
                py = max(o.sliceResolution[0] - img.shape[0], 0)
                px = max(o.sliceResolution[1] - img.shape[1], 0)
                if py or px:
                    pads = ((py // 2, py - py // 2), (px // 2, px - px // 2))
                    img, seg = numpy.pad(img, pads, 'constant'), numpy.pad(seg, pads, 'constant')
                img = scipy.ndimage.zoom(img, float(o.sliceResolution[0]) / float(img.shape[0]))
                seg = scipy.ndimage.zoom(seg, float(o.sliceResolution[0]) / float(seg.shape[0]), mode='nearest')
                seg = (seg >= 0.9).astype(seg.dtype)
            if not o.useCrops:
                images.append(img)
                labels.append(seg)
            elif o.cropType == 'random':
                xs = numpy.random.randint(0, high=img.shape[1] - o.cropWidth, size=o.numRandomCropsPerSlice)
                ys = numpy.random.randint(0, high=img.shape[0] - o.cropHeight, size=o.numRandomCropsPerSlice)
                for x, y in zip(xs, ys):
                    images.append(crop(img, y, x, o.cropHeight, o.cropWidth))
                    labels.append(crop(img, y, x, o.cropHeight, o.cropWidth))      # (sic: the reference crops the image twice)
            elif o.cropType == 'center':
                images.append(crop_center(img, o.cropWidth, o.cropHeight))
                labels.append(crop_center(seg, o.cropWidth, o.cropHeight))
            elif o.cropType == 'lesions':                                      # one crop around the centroid of every lesion component
                cc, n = scipy.ndimage.label(seg, structure=numpy.ones((3, 3)))
                for cy, cx in scipy.ndimage.center_of_mass(seg, cc, range(1, n + 1)):
                    cy = min(max(cy, o.cropHeight // 2), img.shape[0] - o.cropHeight // 2)
                    cx = min(max(cx, o.cropWidth // 2), img.shape[1] - o.cropWidth // 2)
                    ic = crop(img, int(cy) - o.cropHeight // 2, int(cx) - o.cropWidth // 2, o.cropHeight, o.cropWidth)
                    sc = crop(seg, int(cy) - o.cropHeight // 2, int(cx) - o.cropWidth // 2, o.cropHeight, o.cropWidth)
                    if ic.shape[0] == o.cropHeight and ic.shape[1] == o.cropWidth:
                        images.append(ic)
                        labels.append(sc)
        return images, labels

    def _create_numpy_arrays(self):
        images, labels, sets = [], [], []
        for p, patient in enumerate(self.patients):
            part = next((s for s in self.SET_TYPES if p in self.patientsSplit.get(s, ())), None)
            if part is None:
                continue
            for nii_filename in patient['filtered_files']:
                im, lb = self.gather_data(patient, nii_filename)
                images += im
                labels += lb
                sets += [self.SET_TYPES.index(part)] * len(im)
        self._images = numpy.array(images).astype(numpy.float32)
        self._labels = numpy.array(labels).astype(numpy.float32)
        if self._images.ndim < 4:
            self._images = numpy.expand_dims(self._images, 3)
        if self._labels.ndim < 4:
            self._labels = numpy.expand_dims(self._labels, 3)
        self._sets = numpy.array(sets).astype(numpy.int32)

    # ------------------------------------------------------------------ accessors / names
    images = property(lambda self: self._images)
    labels = property(lambda self: self._labels)
    sets = property(lambda self: self._sets)
    num_examples = property(lambda self: self._images.shape[0])
    width = property(lambda self: self._images.shape[2])
    height = property(lambda self: self._images.shape[1])
    num_channels = property(lambda self: self._images.shape[3])
    epochs_completed = property(lambda self: self._epochs_completed)

    def get_images(self, set=None):
        return self._images[numpy.where(self._sets == self.SET_TYPES.index(set))[0]]

    def get_image(self, i):
        return self._images[i, :, :, :]

    def get_label(self, i):
        return self._labels[i, :, :, :]

    def name(self):
        o = self.options
        n = self.NAME
        if o.numSamples > 0:
            n += '_n{}'.format(o.numSamples)
        n += '_p{}-{}'.format(o.partition['TRAIN'], o.partition['VAL'])
        if o.useCrops:
            n += '_{}crops{}x{}'.format(o.cropType, o.cropWidth, o.cropHeight)
            if o.cropType == 'random':
                n += '_{}cropsPerSlice'.format(o.numRandomCropsPerSlice)
        if o.sliceResolution is not None:
            n += '_res{}x{}'.format(o.sliceResolution[0], o.sliceResolution[1])
        return n + '_{}'.format(o.format)

    def dir(self):
        return self.options.dir

    def split_name(self):
        p = self.options.partition
        return os.path.join(self.dir(), 'split-{}-{}-{}.pckl'.format(p['TRAIN'], p['VAL'], p['TEST']))

    def pckl_name(self):
        return os.path.join(self.dir(), self.name() + '.pckl')

    def tfrecord_name(self):
        return os.path.join(self.dir(), self.name() + '.tfrecord')

    # ------------------------------------------------------------------ batches
    def _members(self, set):
        return numpy.where(self._sets == self.SET_TYPES.index(set))[0]

    def num_batches(self, batchsize, set='TRAIN'):
        return len(self._members(set)) // batchsize

    def _shuffle(self, members):
        perm = numpy.random.permutation(len(members))
        self._images[members] = self._images[members[perm]]
        self._labels[members] = self._labels[members[perm]]

    def next_batch(self, batch_size, shuffle=True, set='TRAIN', return_brainmask=True):
        members = self._members(set)
        n = len(members)
        start = self._index_in_epoch[set]
        if self._epochs_completed[set] == 0 and start == 0 and shuffle:
            self._shuffle(members)
        if start + batch_size > n:
            self._epochs_completed[set] += 1
            rest_i, rest_l = self._images[members[start:n]], self._labels[members[start:n]]
            if shuffle:
                self._shuffle(members)
            end = batch_size - (n - start)
            self._index_in_epoch[set] = end
            images = numpy.concatenate((rest_i, self._images[members[0:end]]), axis=0)
            labels = numpy.concatenate((rest_l, self._labels[members[0:end]]), axis=0)
        else:
            end = start + batch_size
            self._index_in_epoch[set] = end
            images, labels = self._images[members[start:end]], self._labels[members[start:end]]
        if self.options.addInstanceNoise:
            images = images + numpy.random.normal(0, 0.01, images.shape)
        assert images.size, 'The batch is empty!'
        assert labels.size, 'The labels of the current batch are empty!'
        return images, labels, (images > 0.05) if return_brainmask else None
