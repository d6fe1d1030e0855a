#!/usr/bin/env python
"""Mirror of reference mains/main_GMVAE_spatial.py: same hyper-parameters and call sequence, no TensorFlow session."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from unsupervised_anomaly_detection_brain_mri_b200.models.gaussian_mixture_variational_autoencoder_spatial import gaussian_mixture_variational_autoencoder_spatial
from unsupervised_anomaly_detection_brain_mri_b200.trainers.GMVAE_spatial import GMVAE_spatial
from unsupervised_anomaly_detection_brain_mri_b200.utils import Evaluation
from unsupervised_anomaly_detection_brain_mri_b200.utils.default_config_setup import Dataset, get_config, get_datasets, get_options

dataset = Dataset.BRAINWEB
options = get_options(batchsize=8, learningrate=5e-5, numEpochs=1, zDim=128, outputWidth=128, outputHeight=128)
options['data']['dir'] = options["globals"][dataset.value]
datasetHC, datasetPC = get_datasets(options, dataset=dataset)
config = get_config(trainer=GMVAE_spatial, options=options, optimizer='ADAM', intermediateResolutions=[8, 8], dropout_rate=0.1, dataset=datasetHC)

config.dim_c = 9
config.dim_z = 1
config.dim_w = 1
config.c_lambda = 1
config.restore_lr = 1e-3
config.restore_steps = 150
config.tv_lambda = -1.0

# Create an instance of the model and train it
model = GMVAE_spatial(None, config, network=gaussian_mixture_variational_autoencoder_spatial)

# Train it
model.train(datasetHC)

# Evaluate
Evaluation.evaluate(datasetPC, model, options, description=f"{type(datasetHC).__name__}-{options['threshold']}", epoch=str(options['train']['numEpochs']))
