"""Base trainer: config, checkpoint save/load, optimiser factory (mirror of reference trainers/DLMODEL.py:11-152).

Checkpoints keep the reference's directory layout and side files (``<dir>/<model_dir>/<modelname>.model-<step>``,
``Config-<step>.json``, ``Curves.npy``, DLMODEL.py:63-84) but store the weights as ``.npz`` keyed by the TF variable
names instead of a TF tensor bundle.  As in the reference, Adam moments are NOT checkpointed (the Saver is created
before the optimiser there: AE.py:21 vs :32).  ``export_tf_checkpoint`` / ``import_tf_checkpoint`` additionally write /
read the reference's own on-disk format (tf.train.Saver V2 tensor bundles, ``utils/tf_checkpoint.py``)."""
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
        if iteration is None and not (ckpt_name and os.path.isfile(os.path.join(checkpoint_dir, ckpt_name + '.npz'))):
            # the state file may name a step that exists only as an exported TF bundle (export_tf_checkpoint shares the
            # `checkpoint` file, as tf.train.Saver would): fall back to the newest native checkpoint instead of restarting from 0
            pat = re.compile(re.escape(self.config.modelname) + r'\.model-(\d+)\.npz$')
            have = sorted((int(m.group(1)), f[:-4]) for f in (os.listdir(checkpoint_dir) if os.path.isdir(checkpoint_dir) else [])
                          for m in [pat.match(f)] if m)
            if have:
                ckpt_name = have[-1][1]
        if ckpt_name and os.path.isfile(os.path.join(checkpoint_dir, ckpt_name + '.npz')):
            with np.load(os.path.join(checkpoint_dir, ckpt_name + '.npz')) as z:
                self._load_weights({k.replace('|', '/'): z[k] for k in z.files})
            counter = int(next(re.finditer(r'(\d+)(?!.*\d)', ckpt_name)).group(0))
            print(" [*] Success to read {}".format(ckpt_name))
            return True, counter
        print(" [*] Failed to find a checkpoint")
        return False, 0

    # ------------------------------------------------------------------ TensorFlow checkpoint files (tf.train.Saver V2)
    def export_tf_checkpoint(self, checkpoint_dir, step, with_optimizer=False, suffix_scheme='per_scope'):
        """Writes `<checkpoint_dir>/<model_dir>/<modelname>.model-<step>.{index,data-00000-of-00001}` + the `checkpoint`
        state file in the format the reference's `self.saver.save(...)` produces (trainers/DLMODEL.py:66-74), so a
        TensorFlow installation of the reference can `load()` weights trained here.  Variables: the trainable ones and the
        (frozen) BatchNormalization moving statistics - exactly the set the reference's Saver holds (it is created before
        the optimiser, AE.py:21 vs :32, so the Adam slots are NOT in its files); with_optimizer=True adds
        `<var>/Adam`, `<var>/Adam_1`, `beta1_power`, `beta2_power` as a Saver created after `minimize()` would."""
        from ..utils import tf_checkpoint as tfc
        eng = self.engine
        weights = self._weights()
        m = v = None
        if with_optimizer and hasattr(eng, 'adam_step') and isinstance(getattr(eng, 't', None), int):
            if getattr(eng, 'peer', None) is not None:       # fused peer optimiser: every rank holds the moments of its shard only
                eng.peer.gather_adam_state(eng.fp.m, eng.fp.v)   # (a collective - every rank has to make this call)
            m, v = eng.fp.to_numpy(eng.fp.m), eng.fp.to_numpy(eng.fp.v)
        variables = tfc.saver_variables(weights, m, v, step=int(getattr(eng, 't', 0)) if m is not None else 0,
                                        beta1=float(getattr(self.config, 'beta1', 0.5)), beta2=0.999)
        # un-named layers: tf.compat.v1.layers (what the reference's graphs are built from) numbers them per variable scope;
        # suffix_scheme='graph' keeps this code's graph-wide numbering (import accepts either, tf_checkpoint.resolve_layer_names)
        if suffix_scheme == 'per_scope':
            variables = tfc.rename_prefixes(variables, tfc.per_scope_layer_names(list(weights)))
        elif suffix_scheme != 'graph':
            raise ValueError(f"suffix_scheme must be 'per_scope' or 'graph', not {suffix_scheme!r}")
        directory = os.path.join(checkpoint_dir, self.model_dir)
        name = f'{self.config.modelname}.model-{step}'
        tfc.write_bundle(os.path.join(directory, name), variables)
        tfc.update_checkpoint_state(directory, name)
        return os.path.join(directory, name)

    def import_tf_checkpoint(self, prefix_or_dir, load_optimizer=True):
        """Loads a TensorFlow checkpoint of the reference (a `...model-<step>` prefix, or a directory holding a `checkpoint`
        state file) into the engine: weights by TF variable name, Adam moments when present.  Returns the step parsed
        from the name (reference DLMODEL.load :104).  Raises if a variable is missing, a shape differs, or the file carries
        BatchNormalization moving statistics other than the 0 / 1 the reference never updates."""
        from ..utils import tf_checkpoint as tfc
        prefix = prefix_or_dir
        if os.path.isdir(prefix_or_dir):
            prefix = tfc.latest_checkpoint(prefix_or_dir)
            if prefix is None:
                raise FileNotFoundError(f'no checkpoint state file in {prefix_or_dir}')
        variables = tfc.read_bundle(prefix)
        eng = self.engine
        wanted = list(eng.specs.keys())
        if any(n not in variables for n in wanted):
            # same layers, other automatic suffixes (per-scope vs whole-graph counters): pair them by creation order
            variables = tfc.rename_layers(variables, tfc.resolve_layer_names(wanted, variables.keys()))
        missing = [n for n in wanted if n not in variables]
        if missing:
            raise KeyError(f'{prefix}: variables missing from the checkpoint: {missing[:4]}{"..." if len(missing) > 4 else ""}')
        for n in wanted:
            if tuple(variables[n].shape) != tuple(eng.specs[n]):
                raise ValueError(f'{prefix}: {n} has shape {variables[n].shape}, the graph needs {tuple(eng.specs[n])}')
        weights, m, v, frozen = tfc.split_saver_variables(variables, wanted)
        if not frozen:
            raise ValueError(f'{prefix}: BatchNormalization moving statistics differ from 0 / 1 - the frozen-BN kernels '
                             f'(SURVEY App. A.3) would not reproduce this checkpoint')
        self._load_weights(weights)
        if load_optimizer and m is not None and hasattr(eng, 'adam_step') and isinstance(getattr(eng, 't', None), int):
            eng.fp.load(m, buf=eng.fp.m)
            eng.fp.load(v, buf=eng.fp.v)
            if 'beta1_power' in variables:                      # beta1_power = beta1^(t+1)  ->  t
                import math
                b1 = float(getattr(self.config, 'beta1', 0.5))
                t = int(round(math.log(float(variables['beta1_power'])) / math.log(b1))) - 1 if 0 < b1 < 1 else 0
                eng.t = max(t, 0)
                eng.step_dev.fill_(eng.t)
        mm = re.search(r'(\d+)(?!.*\d)', os.path.basename(prefix))
        return int(mm.group(0)) if mm else 0

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
