"""BrainWeb slice dataset with the interface of reference dataloaders/BRAINWEB.py (class BRAINWEB, Options attributes,
patients / patients_split, load_volume_and_groundtruth, get_patient_idx, num_batches, next_batch, name / *_name helpers) -
on the numpy-only volume readers of utils/MINC.py / utils/NII.py and the TensorFlow-free TFRecord cache.

Directory layout (reference Options :27-56, get_patients :209-251):
    <dir>/normal/*.mnc.gz, <dir>/lesions/{mild,moderate,severe}/*.mnc.gz, <dir>/groundtruth/{normal,mild_lesions,...}.mnc.gz
Pipeline per patient (:262-298, :124-186): optional skull / background removal from the tissue labels, percentile clip
(0, 99.8) + normalisation, lesion label -> binary ground truth, then per axial slice in [sliceStart, sliceEnd): skip blank
slices, resize (cv2) or zero-pad to sliceResolution, optional rotations / crops.  Patients (not slices) are split into
TRAIN / VAL / TEST (:84-107); the split is stored as `split-<p>.pckl` (a plain dict of patient-name lists, as the reference
writes it) and the slices as `<name>.tfrecord` when options.cache is set.
Not carried over: the pickled copy of the dataset object (`<name>.pckl`) - the TFRecord + split file are the cache here -
and the matplotlib / imageio helpers (visualize, export_slices)."""
import glob
import math
import os
import pickle

import numpy

from ..utils.image_utils import crop, crop_center
from ..utils.MINC import MINC
from ..utils.tfrecord_utils import read_tf_record, write_tf_record


class BRAINWEB(object):
    FILTER_TYPES = ['NORMAL', 'MILDMS', 'MODERATEMS', 'SEVEREMS']
    SET_TYPES = ['TRAIN', 'VAL', 'TEST']
    LABELS = {'BACKGROUND': 0, 'CSF': 1, 'GM': 2, 'WM': 3, 'FAT': 4, 'MUSCLE': 5, 'SKIN': 6, 'SKULL': 7, 'GLIALMATTER': 8,
              'CONNECTIVE': 9, 'LESION': 10}
    VIEW_MAPPING = {'saggital': 0, 'coronal': 1, 'axial': 2}
    PROTOCOL_MAPPINGS = {'FLAIR': 'flair*', 'T2': 't2*'}
    NON_BRAIN = ('FAT', 'MUSCLE', 'SKIN', 'SKULL', 'CONNECTIVE')
    GT_FILES = {'NORMAL': 'normal.mnc.gz', 'MILDMS': 'mild_lesions.mnc.gz', 'MODERATEMS': 'moderate_lesions.mnc.gz',
                'SEVEREMS': 'severe_lesions.mnc.gz'}

    class Options(object):
        def __init__(self):
            self.description = None
            self.dir = os.path.dirname(os.path.realpath(__file__))
            self.folderNormal = 'normal'
            self.folderMildMS = os.path.join('lesions', 'mild')
            self.folderModerateMS = os.path.join('lesions', 'moderate')
            self.folderSevereMS = os.path.join('lesions', 'severe')
            self.folderGT = 'groundtruth'
            self.numSamples = -1
            self.partition = {'TRAIN': 0.6, 'VAL': 0.15, 'TEST': 0.25}
            self.sliceStart = 20
            self.sliceEnd = 140
            self.useCrops = False
            self.cropType = 'random'            # random or center
            self.numRandomCropsPerSlice = 5
            self.rotations = [0]
            self.cropWidth = 128
            self.cropHeight = 128
            self.cache = False
            self.sliceResolution = None          # HxW
            self.addInstanceNoise = False        # a little Gaussian noise on every sampled batch
            self.filterProtocol = None           # 'T2' or 'FLAIR'
            self.filterType = None               # subset of FILTER_TYPES
            self.axis = 'axial'
            self.debug = False
            self.normalizationMethod = 'standardization'
            self.skullRemoval = False
            self.backgroundRemoval = False

    def __init__(self, options=None):
        self.options = options if options is not None else BRAINWEB.Options()
        o = self.options
        self.patients = BRAINWEB.get_patients(o)
        self._epochs_completed = {s: 0 for s in BRAINWEB.SET_TYPES}
        self._index_in_epoch = {s: 0 for s in BRAINWEB.SET_TYPES}
        self.patients_split = self._load_or_make_split()
        if o.cache and os.path.isfile(self.tfrecord_name()):
            self._images, self._labels, self._sets = read_tf_record(self.tfrecord_name())
            self._sets = self._sets.reshape(-1)
        else:
            self._extract_slices()
            if o.cache:
                write_tf_record(self._images, self._labels, self._sets, self.tfrecord_name())

    # ------------------------------------------------------------------ patients and their split
    @staticmethod
    def get_patients(options):
        folders = (('NORMAL', options.folderNormal), ('MILDMS', options.folderMildMS), ('MODERATEMS', options.folderModerateMS),
                   ('SEVEREMS', options.folderSevereMS))
        pattern = (BRAINWEB.PROTOCOL_MAPPINGS[options.filterProtocol] if options.filterProtocol else '*') + '.mnc.gz'
        wanted = options.filterType if options.filterType is not None else BRAINWEB.FILTER_TYPES
        patients = []
        for kind, folder in folders:
            if kind not in wanted:
                continue
            for path in sorted(glob.glob(os.path.join(options.dir, folder, pattern))):
                patients.append({'name': os.path.basename(path), 'type': kind, 'fullpath': path, 'filtered_files': path,
                                 'groundtruth_filename': os.path.join(options.dir, options.folderGT, BRAINWEB.GT_FILES[kind])})
        return patients

    def _load_or_make_split(self):
        if os.path.isfile(self.split_name()):
            with open(self.split_name(), 'rb') as f:
                split = pickle.load(f)
            return self._names_of(split)
        n = len(self.patients)
        order = numpy.random.permutation(n)
        split, taken = {}, 0
        for part, share in self.options.partition.items():
            count = max(1, math.floor(share * n)) if 1.0 >= share > 0.0 else int(share)
            count = min(count, n - taken)
            split[part] = order[taken:taken + count]
            taken += count
        split = self._names_of(split)
        os.makedirs(self.dir(), exist_ok=True)
        with open(self.split_name(), 'wb') as f:
            pickle.dump(split, f)
        return split

    def _names_of(self, split):
        """Patient indices (the reference's old split format) -> patient file names (its OS-agnostic format)."""
        out = {}
        for part, members in split.items():
            members = list(members)
            out[part] = [m if isinstance(m, str) else self.patients[int(m)]['name'] for m in members]
        for part in BRAINWEB.SET_TYPES:
            out.setdefault(part, [])
        return out

    def get_patient_idx(self, split='TRAIN'):
        return [i for i, p in enumerate(self.patients) if p['name'] in self.patients_split[split]]

    def get_patient_split(self):
        return self.patients_split

    # ------------------------------------------------------------------ volumes -> slices
    def load_volume_and_groundtruth(self, minc_filename, patient):
        o = self.options
        vol = MINC(patient['fullpath'])
        vol.set_view_mapping(BRAINWEB.VIEW_MAPPING)
        seg = MINC(patient['groundtruth_filename'])
        skullmap = MINC(patient['groundtruth_filename'])
        skullmap.data = numpy.ones_like(skullmap.data)
        tissue = numpy.rint(seg.data).astype(numpy.int64)
        if o.skullRemoval:
            for name in BRAINWEB.NON_BRAIN:
                skullmap.data[tissue == BRAINWEB.LABELS[name]] = 0
        if o.backgroundRemoval:
            skullmap.data[tissue == BRAINWEB.LABELS['BACKGROUND']] = 0
        seg.data = (tissue == BRAINWEB.LABELS['LESION']).astype(seg.data.dtype)          # binary lesion ground truth
        if o.skullRemoval or o.backgroundRemoval:
            vol.apply_skullmap(skullmap)
        # 99.8th percentile as in Nyul et al., "New variants of a method of MRI scale standardization", TMI 19(2), 2000
        vol.normalize(method=o.normalizationMethod, lowerpercentile=0.0, upperpercentile=99.8)
        return vol, seg, skullmap

    def _fit(self, img, seg):
        """Down-sample (cv2) or zero-pad a slice pair to options.sliceResolution."""
        res = self.options.sliceResolution
        if res is None:
            return img, seg
        if img.shape[0] > res[0] or img.shape[1] > res[1]:
            import cv2
            return (cv2.resize(img, tuple(res)), cv2.resize(seg, tuple(res), interpolation=cv2.INTER_NEAREST))
        top, left = (res[0] - img.shape[0]) // 2, (res[1] - img.shape[1]) // 2
        out_i, out_s = numpy.zeros(res, img.dtype), numpy.zeros(res, seg.dtype)
        out_i[top:top + img.shape[0], left:left + img.shape[1]] = img
        out_s[top:top + seg.shape[0], left:left + seg.shape[1]] = seg
        return out_i, out_s

    def _extract_slices(self):
        o = self.options
        images, labels, sets = [], [], []
        for patient in self.patients:
            part = next((s for s in BRAINWEB.SET_TYPES if patient['name'] in self.patients_split[s]), None)
            if part is None:
                continue
            set_id = BRAINWEB.SET_TYPES.index(part)
            vol, seg, _ = self.load_volume_and_groundtruth(patient['filtered_files'], patient)
            for s in range(o.sliceStart, min(o.sliceEnd, vol.num_slices_along_axis(o.axis))):
                if 0 < o.numSamples < len(images):
                    break
                img, lab = vol.get_slice(s, o.axis), seg.get_slice(s, o.axis)
                if numpy.unique(img).size == 1:                # blank slice
                    continue
                img, lab = self._fit(img, lab)
                for angle in o.rotations:
                    if angle != 0:
                        from scipy.ndimage import rotate
                        img_r, lab_r = rotate(img, angle, reshape=False), rotate(lab, angle, reshape=False, mode='nearest')
                    else:
                        img_r, lab_r = img, lab
                    if o.useCrops and o.cropType == 'random':
                        xs = numpy.random.randint(0, high=img_r.shape[1] - o.cropWidth, size=o.numRandomCropsPerSlice)
                        ys = numpy.random.randint(0, high=img_r.shape[0] - o.cropHeight, size=o.numRandomCropsPerSlice)
                        for x, y in zip(xs, ys):
                            images.append(crop(img_r, y, x, o.cropHeight, o.cropWidth))
                            labels.append(crop(img_r, y, x, o.cropHeight, o.cropWidth))    # (sic: the reference crops the image twice, :167)
                            sets.append(set_id)
                    elif o.useCrops and o.cropType == 'center':
                        images.append(crop_center(img_r, o.cropWidth, o.cropHeight))
                        labels.append(crop_center(lab_r, o.cropWidth, o.cropHeight))
                        sets.append(set_id)
                    elif not o.useCrops:
                        images.append(img_r)
                        labels.append(lab_r)
                        sets.append(set_id)
        self._images = numpy.array(images).astype(numpy.float32)
        self._labels = numpy.array(labels).astype(numpy.float32)
        if self._images.ndim < 4:
            self._images = numpy.expand_dims(self._images, 3)
        if self._labels.ndim < 4:
            self._labels = numpy.expand_dims(self._labels, 3)
        self._sets = numpy.array(sets).astype(numpy.int32)

    # ------------------------------------------------------------------ accessors (reference :311-352)
    @property
    def images(self):
        return self._images

    @property
    def labels(self):
        return self._labels

    @property
    def sets(self):
        return self._sets

    def get_images(self, set=None):
        return self._images[numpy.where(self._sets == BRAINWEB.SET_TYPES.index(set))[0]]

    def get_image(self, i):
        return self._images[i, :, :, :]

    def get_label(self, i):
        return self._labels[i, :, :, :]

    @property
    def num_examples(self):
        return self._images.shape[0]

    @property
    def width(self):
        return self._images.shape[2]

    @property
    def height(self):
        return self._images.shape[1]

    @property
    def num_channels(self):
        return self._images.shape[3]

    @property
    def epochs_completed(self):
        return self._epochs_completed

    # ------------------------------------------------------------------ file names (reference :358-387)
    def name(self):
        o = self.options
        n = 'BRAINWEB'
        if o.description:
            n += '_{}'.format(o.description)
        if o.numSamples > 0:
            n += '_n{}'.format(o.numSamples)
        n += '_p{}-{}-{}'.format(o.partition['TRAIN'], o.partition['VAL'], o.partition['TEST'])
        if o.useCrops:
            n += '_{}crops{}x{}'.format(o.cropType, o.cropWidth, o.cropHeight)
            if o.cropType == 'random':
                n += '_{}cropsPerSlice'.format(o.numRandomCropsPerSlice)
        if o.sliceResolution is not None:
            n += '_res{}x{}'.format(o.sliceResolution[0], o.sliceResolution[1])
        if o.skullRemoval:
            n += '_noSkull'
        if o.backgroundRemoval:
            n += '_noBackground'
        return n

    def dir(self):
        return self.options.dir

    def pckl_name(self):
        return os.path.join(self.dir(), self.name() + '.pckl')

    def tfrecord_name(self):
        return os.path.join(self.dir(), self.name() + '.tfrecord')

    def split_name(self):
        p = self.options.partition
        return os.path.join(self.dir(), 'split-{}-{}-{}.pckl'.format(p['TRAIN'], p['VAL'], p['TEST']))

    # ------------------------------------------------------------------ batches (reference :406-478)
    def _members(self, set):
        return numpy.where(self._sets == BRAINWEB.SET_TYPES.index(set))[0]

    def num_batches(self, batchsize, set='TRAIN'):
        return len(self._members(set)) // batchsize

    def _shuffle(self, members):
        perm = numpy.random.permutation(len(members))
        self._images[members] = self._images[members[perm]]
        self._labels[members] = self._labels[members[perm]]

    def next_batch(self, batch_size, shuffle=True, set='TRAIN', return_brainmask=False):
        """The next batch_size slices of a set; an epoch boundary completes the batch from the reshuffled set."""
        members = self._members(set)
        n = len(members)
        start = self._index_in_epoch[set]
        # (the reference's first-epoch shuffle, :419, compares the per-set DICT with 0 and so never runs: the first epoch
        #  is served in extraction order; kept, so that a seeded run draws the same batches)
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
        if not return_brainmask:
            return images, labels, None
        masks = numpy.copy(labels)
        for name in BRAINWEB.NON_BRAIN + ('BACKGROUND',):
            masks[masks == BRAINWEB.LABELS[name]] = 0
        masks[masks > 0] = 1
        return images, labels, masks
