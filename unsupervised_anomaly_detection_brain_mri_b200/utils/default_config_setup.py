"""options / dataset / config plumbing (mirror of reference utils/default_config_setup.py:13-271, pure Python there too)."""
import json
import os
from enum import Enum

from ..dataloaders.BRAINWEB import BRAINWEB
from ..dataloaders.MSISBI2015 import MSISBI2015
from ..dataloaders.MSLUB import MSLUB
from ..dataloaders.MSSEG2008 import MSSEG2008
from ..dataloaders.SYNTHETIC import SYNTHETIC

base_path = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class Dataset(Enum):
    BRAINWEB = 'BRAINWEBDIR'
    MSSEG2008_UNC = 'MSSEG2008DIR'
    MSSEG2008_CHB = 'MSSEG2008DIR'
    MSISBI2015 = 'MSISBI2015DIR'
    MSLUB = 'MSLUBDIR'
    Brainweb = 'BRAINWEBDIR'      # reference run.py:67,73,84,90 spells it this way; accept both
    SYNTHETIC = 'SYNTHETICDIR'


def get_options(batchsize, learningrate, numEpochs, zDim, outputWidth, outputHeight, slices_start=20, slices_end=130,
                numMonteCarloSamples=0, config=None):
    options = {}
    if config:
        options["globals"] = config
    else:
        path = os.path.join(base_path, "config.default.json")
        if os.path.isfile(path):
            with open(path, 'r') as f:
                options["globals"] = json.load(f)
        else:
            options["globals"] = {"CHECKPOINTDIR": "checkpoints", "SAMPLEDIR": "samples"}
    options['debug'] = False
    options['data'] = {}
    options['train'] = {}
    options['train']['checkpointDir'] = options["globals"]["CHECKPOINTDIR"]
    options['train']['samplesDir'] = options["globals"]["SAMPLEDIR"]
    options['train']['batchsize'] = batchsize
    options['train']['learningrate'] = learningrate
    options['train']['numEpochs'] = numEpochs
    options['train']['zDim'] = zDim
    options['train']['snapshotAfter'] = 1000
    options['train']['outputWidth'] = outputWidth
    options['train']['outputHeight'] = outputHeight
    options['train']['useTensorboard'] = True
    options['train']['useMatplotlib'] = False
    options['train']['tensorboardPort'] = 9001
    options['sliceStart'] = slices_start
    options['sliceEnd'] = slices_end
    options['threshold'] = 'bestdice'
    options['exportVolumes'] = False
    options['exportPRC'] = True
    options['exportROC'] = True
    options['numMonteCarloSamples'] = numMonteCarloSamples
    options['keepOnlyPositiveResiduals'] = True
    options['applyHyperIntensityPrior'] = True
    options['medianFiltering'] = True
    options['erodeBrainmask'] = True
    return options


def get_synthetic_dataset_options(options, lesions, num_patients=None):
    o = SYNTHETIC.Options()
    o.sliceResolution = [options['train']['outputHeight'], options['train']['outputWidth']]
    o.sliceStart = options['sliceStart']
    o.sliceEnd = options['sliceEnd']
    o.lesions = lesions
    if num_patients is not None:
        o.numPatients = num_patients
    elif 'numPatients' in options.get('data', {}):
        o.numPatients = options['data']['numPatients']
    return o


def get_datasets(options, dataset: Dataset = Dataset.BRAINWEB):
    """(healthy train/val set, lesion test set) - reference default_config_setup.py:60-72.
    BRAINWEB with a real data directory (``<dir>/normal/*.mnc.gz`` present) goes through the BrainWeb loader exactly as the
    reference configures it (:200-242); MSLUB / MSISBI2015 / MSSEG2008 with data on disk return (None, lesion set) as there
    (:63-70, :78-194).  Without data on disk every member maps to the synthetic BrainWeb-shaped generator (healthy slices for
    training, slices with lesions + labels for testing), which is also what the metric is defined on."""
    if dataset in (Dataset.BRAINWEB, Dataset.Brainweb) and has_brainweb_data(options.get('data', {}).get('dir')):
        return get_Brainweb_healthy_dataset(options), get_Brainweb_lesion_dataset(options)
    g = options.get('globals', {})
    if dataset == Dataset.MSLUB and os.path.isdir(os.path.join(g.get('MSLUBDIR', ''), 'data')):
        return None, get_MSLUB_dataset(options)
    if dataset == Dataset.MSISBI2015 and os.path.isdir(os.path.join(g.get('MSISBI2015DIR', ''), 'training01')):
        return None, get_MSISBI2015_dataset(options)
    if dataset.name.startswith('MSSEG2008') and os.path.isdir(os.path.join(g.get('MSSEG2008DIR', ''), dataset.name[-3:] + '_train')):
        return None, get_MSSEG2008_dataset(options, dataset.name[-3:])
    hc = get_synthetic_dataset_options(options, lesions=False)
    hc.partition = {'TRAIN': 0.7, 'VAL': 0.3, 'TEST': 0.0}
    pc = get_synthetic_dataset_options(options, lesions=True, num_patients=options.get('data', {}).get('numTestPatients', 2))
    pc.partition = {'TRAIN': 0.0, 'VAL': 0.0, 'TEST': 1.0}
    pc.seed = 4321
    return SYNTHETIC(hc), SYNTHETIC(pc)


def has_brainweb_data(directory):
    import glob
    return bool(directory) and bool(glob.glob(os.path.join(directory, BRAINWEB.Options().folderNormal, '*.mnc.gz')))


def _lesion_options(cls, options, directory, partition):
    """The option block the reference repeats for its three NIfTI lesion sets (default_config_setup.py:87-114, 129-155, 169-194)."""
    o = cls.Options()
    o.description = ''
    o.debug = options['debug']
    o.dir = directory
    o.useCrops = False
    o.cropType = 'center'
    o.cropWidth = options['train']['outputWidth']
    o.cropHeight = options['train']['outputHeight']
    o.numRandomCropsPerSlice = 5
    o.rotations = [0]
    o.partition = partition
    o.sliceResolution = [options['train']['outputHeight'], options['train']['outputWidth']]
    o.cache = True
    o.numSamples = -1
    o.addInstanceNoise = False
    o.axis = 'axial'
    o.filterProtocols = ['FLAIR']
    o.normalizationMethod = 'scaling'
    o.skullStripping = True
    o.sliceStart = options['sliceStart']
    o.sliceEnd = options['sliceEnd']
    o.format = 'aligned'
    return o


def get_MSLUB_dataset_options(options):
    return _lesion_options(MSLUB, options, options['globals']['MSLUBDIR'], {'TRAIN': 0, 'VAL': 5, 'TEST': 25})


def get_MSLUB_dataset(options):
    return MSLUB(get_MSLUB_dataset_options(options))


def get_MSISBI2015_dataset_options(options):
    o = _lesion_options(MSISBI2015, options, options['globals']['MSISBI2015DIR'], {'TRAIN': 0, 'VAL': 5, 'TEST': 15})
    o.filterType = 'train'
    return o


def get_MSISBI2015_dataset(options):
    return MSISBI2015(get_MSISBI2015_dataset_options(options))


def get_MSSEG2008_dataset_options(options, filter_sanner):
    o = _lesion_options(MSSEG2008, options, options['globals']['MSSEG2008DIR'], {'TRAIN': 0, 'VAL': 2, 'TEST': 8})
    o.filterScanner = filter_sanner          # 'UNC' or 'CHB'
    o.filterType = 'train'
    return o


def get_MSSEG2008_dataset(options, filter_sanner):
    return MSSEG2008(get_MSSEG2008_dataset_options(options, filter_sanner))


def get_Brainweb_healthy_dataset(options):
    return BRAINWEB(get_Brainweb_dataset_options(options))


def get_Brainweb_lesion_dataset(options):
    dataset_options = get_Brainweb_dataset_options(options)
    dataset_options.partition = {'TRAIN': 0.0, 'VAL': 0.0, 'TEST': 1.0}      # patients with lesions: only for testing
    dataset_options.filterType = 'SEVEREMS'
    dataset_options.rotations = [0]
    return BRAINWEB(dataset_options)


def get_Brainweb_dataset_options(options):
    """reference default_config_setup.py:217-242, field for field."""
    o = BRAINWEB.Options()
    o.description = ""
    o.debug = options['debug']
    o.dir = options['data']['dir']
    o.useCrops = False
    o.cropType = 'center'
    o.cropWidth = options['train']['outputWidth']
    o.cropHeight = options['train']['outputHeight']
    o.numRandomCropsPerSlice = 5
    o.rotations = [0]
    o.partition = {'TRAIN': 0.7, 'VAL': 0.3, 'TEST': 0.0}
    o.sliceResolution = [options['train']['outputHeight'], options['train']['outputWidth']]
    o.cache = True
    o.numSamples = -1
    o.addInstanceNoise = False
    o.axis = 'axial'
    o.filterType = 'NORMAL'
    o.filterProtocol = 'T2'
    o.normalizationMethod = 'scaling'
    o.skullRemoval = True
    o.sliceStart = options['sliceStart']
    o.sliceEnd = options['sliceEnd']
    o.backgroundRemoval = True
    o.registerTo = None
    return o


def get_config(trainer, options, optimizer, intermediateResolutions, dropout_rate, dataset):
    config = trainer.Config()
    config.dataset = type(dataset).__name__
    config.description = ''
    config.numChannels = dataset.num_channels
    config.batchsize = options['train']['batchsize']
    config.checkpointDir = options['train']['checkpointDir']
    config.snapShotAfter = options['train']['snapshotAfter']
    config.sampleDir = options['train']['samplesDir']
    config.learningrate = options['train']['learningrate']
    config.numEpochs = options['train']['numEpochs']
    config.zDim = options['train']['zDim']
    config.beta1 = 0.5
    config.outputHeight = options['train']['outputHeight']
    config.outputWidth = options['train']['outputWidth']
    config.useTensorboard = options['train']['useTensorboard']
    config.useMatplotlib = options['train']['useMatplotlib']
    config.tensorboardPort = options['train']['tensorboardPort']
    config.debugGradients = options['debug']
    config.optimizer = optimizer
    config.intermediateResolutions = intermediateResolutions
    config.weightRegularization = 0.0
    config.dropout_rate = dropout_rate
    config.dropout = False
    config.l1_weight = 1.0
    config.options = options
    return config
