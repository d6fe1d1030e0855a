#!/usr/bin/env python
"""Drop-in driver with the reference's CLI (reference run.py:119-152): same flags, same call sequence
(options -> datasets -> config -> Trainer(sess, config, network) -> train -> evaluate), no TensorFlow."""
import argparse
import importlib
import json
import os
import sys

PKG = 'unsupervised_anomaly_detection_brain_mri_b200'
base_path = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, base_path)

from unsupervised_anomaly_detection_brain_mri_b200.utils.Evaluation import determine_threshold_on_labeled_patients, evaluate  # noqa: E402
from unsupervised_anomaly_detection_brain_mri_b200.utils.default_config_setup import (Dataset, get_config, get_datasets,  # noqa: E402
                                                                                       get_options)


def main(args):
    # trainer class name == file name, network function name == file name (reference run.py:21-24)
    trainer = getattr(importlib.import_module(f'{PKG}.trainers.{args.trainer}'), args.trainer)
    network = getattr(importlib.import_module(f'{PKG}.models.{args.model}'), args.model)
    with open(os.path.join(base_path, args.config), 'r') as f:
        json_config = json.load(f)
    dataset = Dataset.BRAINWEB
    options = get_options(batchsize=args.batchsize, learningrate=args.lr, numEpochs=args.numEpochs, zDim=args.zDim,
                          outputWidth=args.outputWidth, outputHeight=args.outputHeight, slices_start=args.slices_start,
                          slices_end=args.slices_end, numMonteCarloSamples=args.numMonteCarloSamples, config=json_config)
    options['data']['dir'] = options["globals"][dataset.value]
    if args.numPatients:
        options['data']['numPatients'] = args.numPatients
    dataset_hc, dataset_pc = get_datasets(options, dataset=dataset)
    config = get_config(trainer=trainer, options=options, optimizer=args.optimizer,
                        intermediateResolutions=list(args.intermediateResolutions), dropout_rate=0.2, dataset=dataset_hc)
    for arg in vars(args):                      # additional Config parameters (reference run.py:45-47)
        if hasattr(config, arg):
            setattr(config, arg, getattr(args, arg))
    model = trainer(None, config, network=network)     # `sess` is accepted and ignored
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        model.enable_data_parallel()
    model.train(dataset_hc)

    if not args.threshold:
        if args.ds:
            evaluate_optimal(model, options, args.ds)
            return
        for prior in (False, True):
            options['applyHyperIntensityPrior'] = prior
            for ds in (Dataset.BRAINWEB, Dataset.MSLUB, Dataset.MSISBI2015):
                evaluate_optimal(model, options, ds)
    if args.threshold and args.ds:
        evaluate_with_threshold(model, options, args.threshold, args.ds)
    else:
        options['applyHyperIntensityPrior'] = False
        datasetBrainweb = get_evaluation_dataset(options, Dataset.BRAINWEB)
        _bestDiceVAL, _threshVAL = determine_threshold_on_labeled_patients([datasetBrainweb], model, options, description='VAL')
        print(f"Optimal threshold on MS Lesion Validation Set without optimal postprocessing: {_threshVAL} (Dice-Score {_bestDiceVAL})")
        for ds in (Dataset.BRAINWEB, Dataset.MSLUB, Dataset.MSISBI2015):
            evaluate_with_threshold(model, options, _threshVAL, ds)


def evaluate_with_threshold(model, options, threshold, dataset):
    options['applyHyperIntensityPrior'] = False
    options['threshold'] = threshold
    evaluation_dataset = get_evaluation_dataset(options, dataset)
    res = evaluate(evaluation_dataset, model, options, description=f'{type(evaluation_dataset).__name__}-{dataset.name}-VALthresh_{threshold}',
                   epoch=str(options['train']['numEpochs']))
    print(f'[{dataset.name}] DICE @ {threshold}: {res["DICE"]:.4f}')


def evaluate_optimal(model, options, dataset):
    prior = "_wPrior" if options['applyHyperIntensityPrior'] else ''
    evaluation_dataset = get_evaluation_dataset(options, dataset)
    res = evaluate(evaluation_dataset, model, options, epoch=str(options['train']['numEpochs']),
                   description=f'{type(evaluation_dataset).__name__}-{dataset.name}_upperbound_{options["threshold"]}{prior}')
    print(f'[{dataset.name}{prior}] best DICE {res["bestDiceScore"]:.4f} @ {res["bestThreshold"]:.6f}')


def get_evaluation_dataset(options, dataset=Dataset.BRAINWEB):
    options['data']['dir'] = options["globals"].get(dataset.value, '')
    return get_datasets(options, dataset=dataset)[1]


# the reference's command line (reference run.py:119-152): short flag, long flag, default, type, help
FLAGS = [
    ('-c', '--config', 'config.default.json', str, 'config-path'),
    ('-b', '--batchsize', 8, int, 'batchsize'),
    ('-l', '--lr', 0.0001, float, 'learning rate'),
    ('-E', '--numEpochs', 1000, int, 'how many epochs to train'),
    ('-z', '--zDim', 128, int, 'Latent dimension'),
    ('-w', '--outputWidth', 128, int, 'Output width'),
    ('-g', '--outputHeight', 128, int, 'Output height'),
    ('-o', '--optimizer', 'ADAM', str, 'Can be either ADAM, SGD or RMSProp'),
    ('-s', '--slices_start', 20, int, 'slices start'),
    ('-e', '--slices_end', 130, int, 'slices end'),
    ('-t', '--trainer', 'AE', str, 'Can be every class from trainers directory'),
    ('-m', '--model', 'autoencoder', str, 'Can be every class from models directory'),
    ('-O', '--threshold', None, float, 'Use predefined ThreshOld'),
    ('-d', '--ds', None, lambda name: Dataset[name], 'Only evaluate on given dataset'),
    ('-n', '--numMonteCarloSamples', 0, int, 'Amount of Monte Carlos Samples during restoration'),
    ('-G', '--use_gradient_based_restoration', False, bool, 'only for ceVAE'),
    ('-L', '--restore_lr', 1e-3, float, 'only for VAE_You / GMVAE'),
    ('-S', '--restore_steps', 150, int, 'only for VAE_You / GMVAE'),
    ('-T', '--tv_lambda', -1.0, float, 'only for VAE_You / GMVAE'),
    ('-K', '--kappa', 1.0, float, 'only for GANs'),
    ('-M', '--scale', 10.0, float, 'only for GANs'),
    ('-R', '--rho', 1.0, float, 'only for ConstrainedAAE'),
    ('-C', '--dim_c', 9, int, 'only for GMVAE'),
    ('-Z', '--dim_z', 128, int, 'only for GMVAE'),
    ('-W', '--dim_w', 1, int, 'only for GMVAE'),
    ('-A', '--c_lambda', 1, int, 'only for GMVAE'),
]


def build_parser():
    parser = argparse.ArgumentParser(description='Framework')
    for short, long_, default, type_, help_ in FLAGS:
        parser.add_argument(short, long_, default=default, type=type_, help=help_)
    parser.add_argument('-i', '--intermediateResolutions', default=(8, 8), type=int, nargs=2, help='Spatial Bottleneck resolution')
    parser.add_argument('--numPatients', default=0, type=int, help='synthetic dataset size (patients x 110 slices)')
    return parser


if __name__ == '__main__':
    main(build_parser().parse_args())
