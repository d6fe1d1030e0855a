"""Base trainer: config, checkpoint save/load, optimiser factory (mirror of reference trainers/DLMODEL.py:11-152).

Checkpoints keep the reference's directory layout and side files (``<dir>/<model_dir>/<modelname>.model-<step>``,
``Config-<step>.json``, ``Curves.npy``, DLMODEL.py:63-84) but store the weights as ``.npz`` keyed by the TF variable
names instead of a TF tensor bundle.  As in the reference, Adam moments are NOT checkpointed (the Saver is created
before the optimiser there: AE.py:21 vs :32)."""
import json
import os
import re
from abc import abstractmethod

import numpy as np


class DLMODEL(object):
    class Config(object):
        def __init__(self):
            self.modelname = ''
            self.model_config = {}
            self.checkpointDir = None
            self.description = ''
            self.batchsize = 6
            self.useTensorboard = True
            self.tensorboardPort = 8008
            self.useMatplotlib = False
            self.debugGradients = False
            self.tfSummaryAfter = 100
            self.dataset = ''
            self.beta1 = 0.5

    def __init__(self, sess, config=None):
        self.sess = sess                      # accepted for drop-in compatibility, never used
        self.config = config if config is not None else self.Config()
        self.variables = {}
        self.curves = {}
        self.handles = {}
        self.saver = None
        self.losses = None
        self.engine = None

    @abstractmethod
    def train(self, dataset):
        """Train a Deep Neural Network"""

    @property
    def model_dir(self):
        return "{}_d{}_b{}_{}".format(self.config.modelname, self.config.dataset, self.config.batchsize, self.config.description)

    # ------------------------------------------------------------------ checkpoints
    def _weights(self):
        return self.engine.fp.to_numpy()

    def _load_weights(self, values):
        self.engine.fp.load(values)

    def save(self, checkpoint_dir, step):
        model_name = self.config.modelname + ".model"
        checkpoint_dir = os.path.join(checkpoint_dir, self.model_dir)
        os.makedirs(checkpoint_dir, exist_ok=True)
        path = os.path.join(checkpoint_dir, f'{model_name}-{step}')
        np.savez(path + '.npz', **{k.replace('/', '|'): v for k, v in self._weights().items()})
        with open(os.path.join(checkpoint_dir, 'checkpoint'), 'w') as f:
            f.write(f'model_checkpoint_path: "{model_name}-{step}"\n')
        with open(os.path.join(checkpoint_dir, 'Config-{}.json'.format(step)), 'w') as outfile:
            try:
                json.dump(self.config.__dict__, outfile)
            except Exception:
                print("Failed to save config json")
        np.save(os.path.join(checkpoint_dir, 'Curves.npy'), self.curves)

    def load(self, checkpoint_dir, iteration=None):
        print(" [*] Reading checkpoints...")
        checkpoint_dir = os.path.join(checkpoint_dir, self.model_dir)
        curves_file = os.path.join(checkpoint_dir, 'Curves.npy')
        if os.path.isfile(curves_file):
            self.curves = np.load(curves_file, allow_pickle=True).item()
        ckpt_name = None
        if iteration is not None:
            ckpt_name = self.config.modelname + '.model-' + str(iteration)
        else:
            index = os.path.join(checkpoint_dir, 'checkpoint')
            if os.path.isfile(index):
                m = re.search(r'model_checkpoint_path: "([^"]+)"', open(index).read())
                if m:
                    ckpt_name = os.path.basename(m.group(1))
        if ckpt_name and os.path.isfile(os.path.join(checkpoint_dir, ckpt_name + '.npz')):
            with np.load(os.path.join(checkpoint_dir, ckpt_name + '.npz')) as z:
                self._load_weights({k.replace('|', '/'): z[k] for k in z.files})
            counter = int(next(re.finditer(r'(\d+)(?!.*\d)', ckpt_name)).group(0))
            print(" [*] Success to read {}".format(ckpt_name))
            return True, counter
        print(" [*] Failed to find a checkpoint")
        return False, 0

    @staticmethod
    def create_optimizer(loss=None, var_list=(), learningrate=0.001, type='ADAM', beta1=0.05, momentum=0.9, name='optimizer',
                         minimize=True, scope=None):
        """Returns the optimiser description the engine executes.  Only ADAM (the reference's default and the only type
        any mains/ script uses) has a fused kernel; other types raise like an invalid type does in the reference."""
        if type != 'ADAM':
            raise ValueError('Invalid optimizer type (the B200 path implements ADAM)')
        return {'type': 'ADAM', 'learningrate': learningrate, 'beta1': beta1, 'beta2': 0.999, 'epsilon': 1e-8, 'name': name}

    def get_number_of_trainable_params(self):
        if self.engine is None:
            return 0
        scopes = {}
        for name, shape in self.engine.specs.items():
            scopes[name.split('/')[0]] = scopes.get(name.split('/')[0], 0) + int(np.prod(shape))
        for scope, n in scopes.items():
            print(f'#Params in {scope}: {n}')
        total = sum(scopes.values())
        print(f'#Params in total: {total}')
        return total
