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


if __name__ == '__main__':
    args = argparse.ArgumentParser(description='Framework')
    args.add_argument('-c', '--config', default='config.default.json', type=str, help='config-path')
    args.add_argument('-b', '--batchsize', default=8, type=int, help='batchsize')
    args.add_argument('-l', '--lr', default=0.0001, type=float, help='learning rate')
    args.add_argument('-E', '--numEpochs', default=1000, type=int, help='how many epochs to train')
    args.add_argument('-z', '--zDim', default=128, type=int, help='Latent dimension')
    args.add_argument('-w', '--outputWidth', default=128, type=int, help='Output width')
    args.add_argument('-g', '--outputHeight', default=128, type=int, help='Output height')
    args.add_argument('-o', '--optimizer', default='ADAM', type=str, help='Can be either ADAM, SGD or RMSProp')
    args.add_argument('-i', '--intermediateResolutions', default=(8, 8), type=int, nargs=2, help='Spatial Bottleneck resolution')
    args.add_argument('-s', '--slices_start', default=20, type=int, help='slices start')
    args.add_argument('-e', '--slices_end', default=130, type=int, help='slices end')
    args.add_argument('-t', '--trainer', default='AE', type=str, help='Can be every class from trainers directory')
    args.add_argument('-m', '--model', default='autoencoder', type=str, help='Can be every class from models directory')
    args.add_argument('-O', '--threshold', default=None, type=float, help='Use predefined ThreshOld')
    args.add_argument('-d', '--ds', default=None, type=lambda s: Dataset[s], help='Only evaluate on given dataset')
    args.add_argument('-n', '--numMonteCarloSamples', default=0, type=int, help='Amount of Monte Carlos Samples during restoration')
    args.add_argument('-G', '--use_gradient_based_restoration', default=False, type=bool, help='only for ceVAE')
    args.add_argument('-L', '--restore_lr', default=1e-3, type=float, help='only for VAE_You / GMVAE')
    args.add_argument('-S', '--restore_steps', default=150, type=int, help='only for VAE_You / GMVAE')
    args.add_argument('-T', '--tv_lambda', default=-1.0, type=float, help='only for VAE_You / GMVAE')
    args.add_argument('-K', '--kappa', default=1.0, type=float, help='only for GANs')
    args.add_argument('-M', '--scale', default=10.0, type=float, help='only for GANs')
    args.add_argument('-R', '--rho', default=1.0, type=float, help='only for ConstrainedAAE')
    args.add_argument('-C', '--dim_c', default=9, type=int, help='only for GMVAE')
    args.add_argument('-Z', '--dim_z', default=128, type=int, help='only for GMVAE')
    args.add_argument('-W', '--dim_w', default=1, type=int, help='only for GMVAE')
    args.add_argument('-A', '--c_lambda', default=1, type=int, help='only for GMVAE')
    args.add_argument('--numPatients', default=0, type=int, help='synthetic dataset size (patients x 110 slices)')
    main(args.parse_args())
