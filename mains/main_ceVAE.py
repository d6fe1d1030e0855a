#!/usr/bin/env python
"""Mirror of reference mains/main_ceVAE.py: same hyper-parameters and call sequence, no TensorFlow session."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from unsupervised_anomaly_detection_brain_mri_b200.models import context_encoder_variational_autoencoder
from unsupervised_anomaly_detection_brain_mri_b200.trainers.ceVAE import ceVAE
from unsupervised_anomaly_detection_brain_mri_b200.utils import Evaluation
from unsupervised_anomaly_detection_brain_mri_b200.utils.default_config_setup import Dataset, get_config, get_datasets, get_options

dataset = Dataset.BRAINWEB
options = get_options(batchsize=8, learningrate=0.0001, numEpochs=3, zDim=128, outputWidth=128, outputHeight=128)
options['data']['dir'] = options["globals"][dataset.value]
datasetHC, datasetPC = get_datasets(options, dataset=dataset)
config = get_config(trainer=ceVAE, options=options, optimizer='ADAM', intermediateResolutions=[8, 8], dropout_rate=0.1, dataset=datasetHC)

config.use_gradient_based_restoration = 0.1

# Create an instance of the model and train it
model = ceVAE(None, config, network=context_encoder_variational_autoencoder.context_encoder_variational_autoencoder)

# Train it
model.train(datasetHC)

# Evaluate
Evaluation.evaluate(datasetPC, model, options, description=f"{type(datasetHC).__name__}-{options['threshold']}", epoch=str(options['train']['numEpochs']))
