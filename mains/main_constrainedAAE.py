#!/usr/bin/env python
"""Mirror of reference mains/main_constrainedAAE.py: same hyper-parameters and call sequence, no TensorFlow session."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from unsupervised_anomaly_detection_brain_mri_b200.models.constrained_adversarial_autoencoder import constrained_adversarial_autoencoder
from unsupervised_anomaly_detection_brain_mri_b200.trainers.ConstrainedAAE import ConstrainedAAE
from unsupervised_anomaly_detection_brain_mri_b200.utils import Evaluation
from unsupervised_anomaly_detection_brain_mri_b200.utils.default_config_setup import Dataset, get_config, get_datasets, get_options

dataset = Dataset.BRAINWEB
options = get_options(batchsize=8, learningrate=0.001, numEpochs=1, zDim=128, outputWidth=128, outputHeight=128)
options['data']['dir'] = options["globals"][dataset.value]
datasetHC, datasetPC = get_datasets(options, dataset=dataset)
config = get_config(trainer=ConstrainedAAE, options=options, optimizer='ADAM', intermediateResolutions=[16, 16], dropout_rate=0.1, dataset=datasetHC)

config.scale = 1  # 10.0
config.rho = 1.0

# Create an instance of the model and train it
model = ConstrainedAAE(None, config, network=constrained_adversarial_autoencoder)

# Train it
model.train(datasetHC)

# Evaluate
Evaluation.evaluate(datasetPC, model, options, description=f"{type(datasetHC).__name__}-{options['threshold']}", epoch=str(options['train']['numEpochs']))
