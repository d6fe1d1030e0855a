#!/usr/bin/env python
"""Mirror of reference mains/main_fAnoGAN.py: same hyper-parameters and call sequence, no TensorFlow session."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from unsupervised_anomaly_detection_brain_mri_b200.models.fanogan import fanogan
from unsupervised_anomaly_detection_brain_mri_b200.trainers.fAnoGAN import fAnoGAN
from unsupervised_anomaly_detection_brain_mri_b200.utils import Evaluation
from unsupervised_anomaly_detection_brain_mri_b200.utils.default_config_setup import Dataset, get_config, get_datasets, get_options

dataset = Dataset.BRAINWEB
options = get_options(batchsize=8, learningrate=0.001, numEpochs=1, zDim=128, outputWidth=128, outputHeight=128)
options['data']['dir'] = options["globals"][dataset.value]
datasetHC, datasetPC = get_datasets(options, dataset=dataset)
config = get_config(trainer=fAnoGAN, options=options, optimizer='ADAM', intermediateResolutions=[16, 16], dropout_rate=0.1, dataset=datasetHC)

config.kappa = 1.0          # weight of the feature-matching term of the encoder loss
config.scale = 10.0         # gradient-penalty weight of the critic loss

# both training phases (WGAN-GP, then the encoder on the frozen generator / critic) run inside train()
model = fAnoGAN(None, config, network=fanogan)
model.train(datasetHC)

Evaluation.evaluate(datasetPC, model, options, description=f"{type(datasetHC).__name__}-{options['threshold']}", epoch=str(options['train']['numEpochs']))
