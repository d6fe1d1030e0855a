"""f-AnoGAN trainer surface (mirror of reference trainers/fAnoGAN.py).

Implemented on the device this round: construction, checkpoint save/load, ``reconstruct`` (x_enc = sigmoid(G(E(x))),
fAnoGAN.py:220-239) and therefore residual scoring through utils/Evaluation.  NOT yet implemented: ``train`` - the
WGAN-GP critic step needs the gradient of ||d D(x_hat)/d x_hat|| w.r.t. the critic weights, i.e. a double backward
through conv / LayerNormalization (fAnoGAN.py:55-57); it raises NotImplementedError rather than training something
that is not the reference's objective."""
import os
from datetime import datetime

import numpy as np
import torch

from .. import abi
from ..fanogan_engine import FanoganEngine
from ..models.customlayers import Placeholder
from ..utils.logger import Logger, Phase  # noqa: F401
from .AEMODEL import AEMODEL
from .DLMODEL import DLMODEL


class fAnoGAN(DLMODEL):
    class Config(AEMODEL.Config):
        def __init__(self):
            super().__init__('fAnoGAN')
            self.kappa = 1.0
            self.scale = 10.0

    def __init__(self, sess, config=None, network=None):
        super().__init__(sess, config if config is not None else self.Config())
        self.losses = {}
        self.network = network
        cfg = self.config
        self.x = Placeholder([None, cfg.outputHeight, cfg.outputWidth, cfg.numChannels], 'x')
        self.z = Placeholder([None, cfg.zDim], 'z')
        self.dropout = Placeholder([], 'dropout')
        self.dropout_rate = Placeholder([], 'dropout_rate')
        self.outputs = self.network(self.z, self.x, dropout_rate=self.dropout_rate, dropout=self.dropout, config=cfg)
        self.reconstruction = self.outputs['x_enc']
        self.generated = self.outputs['x_']
        self.graph = self.reconstruction.graph
        self.checkpointDir = os.path.join(cfg.checkpointDir or 'checkpoints', self.network.__name__)
        self.logDir = os.path.join(os.getcwd(), 'logs', self.network.__name__, self.model_dir, datetime.now().strftime('%Y%m%d_%H%M%S'))
        self.device = getattr(cfg, 'device', None) or f'cuda:{int(os.environ.get("LOCAL_RANK", 0))}'
        self.math_mode = int(getattr(cfg, 'math_mode', abi.MATH_TC_3XTF32))
        torch.cuda.set_device(self.device)
        g = self.graph
        self.engine = FanoganEngine(g.S, g.C, g.zDim, g.res, batch=cfg.batchsize, device=self.device, math_mode=self.math_mode,
                                    seed=int(getattr(cfg, 'seed', 1)))
        self._eval = {}
        self.get_number_of_trainable_params()
        self.saver = self

    @property
    def model_dir(self):
        return "{}_d{}_s{}x{}_{}_b{}_z{}_{}".format(self.config.modelname, self.config.dataset, self.config.outputWidth,
                                                    self.config.outputHeight, self.network.__name__, self.config.batchsize,
                                                    self.config.zDim, self.config.description)

    def load_checkpoint(self):
        ok, counter = self.load(self.checkpointDir)
        return counter if ok else 0

    def sample_z(self):
        return np.random.normal(size=[self.config.batchsize, self.config.zDim])     # float64, as fAnoGAN.py:241-242

    def train(self, dataset):
        raise NotImplementedError('f-AnoGAN WGAN-GP training (trainers/fAnoGAN.py:45-210) is not implemented on the B200 path yet: '
                                  'the gradient penalty needs a double backward through conv / LayerNormalization. '
                                  'Weights can be loaded with .load(); .reconstruct() and Evaluation.evaluate() run on the device.')

    def _engine_for(self, n):
        if n not in self._eval:
            g = self.graph
            e = FanoganEngine(g.S, g.C, g.zDim, g.res, batch=n, device=self.device, math_mode=self.math_mode)
            e.fp = self.engine.fp            # share the weights
            self._eval[n] = e
        return self._eval[n]

    def reconstruct(self, x, dropout=False):
        if x.ndim < 4:
            x = np.expand_dims(x, 0)
        x = np.ascontiguousarray(x, np.float32)
        N = x.shape[0]
        chunk = min(N, int(getattr(self.config, 'evalBatchsize', 64)))
        eng = self._engine_for(chunk)
        rec = np.empty_like(x)
        rate = self.config.dropout_rate if dropout else 0.0
        keep = 1.0 / (1.0 - rate) if rate > 0 else 1.0
        for i in range(0, N, chunk):
            xb = x[i:i + chunk]
            n = xb.shape[0]
            if n < chunk:
                xb = np.concatenate([xb, np.zeros((chunk - n,) + xb.shape[1:], np.float32)], 0)
            eng.set_inputs(xb)
            m1 = m2 = None
            if rate > 0:                                   # MC-dropout: both Dropout sites of the graph are live
                m1 = torch.empty(chunk, eng.zDim, device=eng.device)
                m2 = torch.empty(chunk, eng.flat, device=eng.device)
                st = torch.cuda.current_stream().cuda_stream
                seed = int(np.random.randint(0, 2 ** 31))
                abi.call('uad_dropout_mask', m1.data_ptr(), m1.numel(), float(rate), seed, 0, None, st)
                abi.call('uad_dropout_mask', m2.data_ptr(), m2.numel(), float(rate), seed, 1 << 40, None, st)
            z = eng.encode(m1, keep)
            rec[i:i + n] = eng.generate(z, m2, keep).cpu().numpy()[:n]
        results = {'reconstruction': rec}
        results['l1err'] = np.sum(np.abs(x - rec))
        results['l2err'] = np.sum(np.sqrt((x - rec) ** 2))
        return results
