"""Adversarial-autoencoder trainer (mirror of reference trainers/AAE.py) on the CUDA engine.

Per mini-batch (AAE.py:84-124): ``d_iters = 20`` optim_ae steps while epoch <= 5 (one afterwards), 20 optim_dis steps (latent
WGAN-GP critic) and one optim_gen step (Encoder scope only); every sess.run feeds a fresh z ~ N(0,1) (get_feed_dict) and draws
fresh dropout / epsilon noise.  Validation fetches the reconstruction and every loss with dropout off (:139-156) and stops early
on the reconstruction loss.

STATUS: the engine's call sequences are verified on CPU against the oracle (tests/test_engine_emulated.py); the first run on
hardware is still to come (tests/test_gpu_aae.py, opt-in)."""
from collections import defaultdict
from math import inf

import numpy as np

from ..aae_engine import AdversarialAEEngine
from ..utils.logger import Phase
from . import trainer_utils
from .AEMODEL import AEMODEL, indicate_early_stopping, update_log_dicts
from .fAnoGAN import fAnoGAN


class AAE(fAnoGAN):
    ENGINE = AdversarialAEEngine
    REC_KEY = 'x_hat'

    class Config(AEMODEL.Config):
        def __init__(self):
            super().__init__('AAE')
            self.scale = 10.0

    def _engine_kwargs(self):
        return dict(scale=float(self.config.scale))

    def sample_z(self):
        return np.random.normal(size=[self.config.batchsize, self.config.zDim])      # AAE.py:188-189

    def _feed(self, batch):
        """get_feed_dict (AAE.py:161-167): x <- batch, z <- sample_z()."""
        self.engine.set_inputs(np.ascontiguousarray(batch, np.float32))
        self.engine.set_latent(self.sample_z().astype(np.float32))

    def train(self, dataset):
        cfg, eng = self.config, self.engine
        eng.scale = float(cfg.scale)
        self.variables = list(eng.specs.keys())
        lr, rate = float(cfg.learningrate), float(cfg.dropout_rate)
        graphs = bool(getattr(cfg, 'useCudaGraph', True))
        kw = dict(dropout_rate=rate, allreduce=getattr(self, '_allreduce', None), world=getattr(self, 'world', 1), use_graph=graphs)
        verbose = bool(getattr(cfg, 'verbose', True))
        d_iters = int(getattr(cfg, 'd_iters', 20))
        best_cost = inf
        last_improvement = 0
        last_epoch = self.load_checkpoint()
        for epoch in range(last_epoch, cfg.numEpochs):
            phase = Phase.TRAIN
            scalars, visuals = defaultdict(list), []
            num_batches = dataset.num_batches(cfg.batchsize, set=phase.value)
            for idx in range(num_batches):
                batch, _, _ = dataset.next_batch(cfg.batchsize, set=phase.value)
                run = {}
                for _ in range(d_iters if epoch <= 5 else 1):
                    self._feed(batch)
                    run = dict(eng.step_ae(lr, dropout=True, **kw))
                run['reconstruction'] = eng.br[0].xhat.cpu().numpy()
                run['L1'] = eng.br[0].l1.cpu().numpy()
                for _ in range(d_iters):
                    self._feed(batch)
                    run['disc_loss'] = eng.step_disc(lr, dropout=True, **kw)['disc_loss']
                self._feed(batch)
                run['gen_loss'] = eng.step_gen(lr, dropout=True, **kw)['gen_loss']
                if verbose:
                    print(f'Epoch ({phase.value}): [{epoch:2d}] [{idx:4d}/{num_batches:4d}] loss: {run["reconstructionLoss"]:.8f},'
                          f' gen_loss: {run["gen_loss"]:.8f}, disc_loss: {run["disc_loss"]:.8f}')
                update_log_dicts(*trainer_utils.get_summary_dict(batch, run), scalars, visuals)
            self.log_to_tensorboard(epoch, scalars, visuals, phase)
            last_epoch += 1
            self.save(self.checkpointDir, last_epoch)

            phase = Phase.VAL
            scalars, visuals = defaultdict(list), []
            num_batches = dataset.num_batches(cfg.batchsize, set=phase.value)
            for idx in range(num_batches):
                batch, _, _ = dataset.next_batch(cfg.batchsize, set=phase.value)
                self._feed(batch)
                run = dict(eng.step_ae(lr, dropout=False, train=False, **kw))
                run['reconstruction'] = eng.br[0].xhat.cpu().numpy()
                run['L1'] = eng.br[0].l1.cpu().numpy()
                run.update(eng.step_disc(lr, dropout=False, apply=False, **kw))       # **self.losses (AAE.py:148)
                if verbose:
                    print(f'Epoch ({phase.value}): [{epoch:2d}] [{idx:4d}/{num_batches:4d}] loss: {run["loss"]:.8f}')
                update_log_dicts(*trainer_utils.get_summary_dict(batch, run), scalars, visuals)
            self.log_to_tensorboard(epoch, scalars, visuals, phase)
            best_cost, last_improvement, stop = indicate_early_stopping(scalars['reconstructionLoss'], best_cost, last_improvement)
            if stop:
                print('Early stopping was triggered due to no improvement over the last 5 epochs')
                break

    def reconstruct(self, x, dropout=False):
        if x.ndim < 4:
            x = np.expand_dims(x, 0)
        x = np.ascontiguousarray(x, np.float32)
        N = x.shape[0]
        chunk = min(N, int(getattr(self.config, 'evalBatchsize', 64)))
        eng = self._engine_for(chunk)
        rec = np.empty_like(x)
        rate = self.config.dropout_rate if dropout else 0.0
        for i in range(0, N, chunk):
            xb = x[i:i + chunk]
            n = xb.shape[0]
            if n < chunk:
                xb = np.concatenate([xb, np.zeros((chunk - n,) + xb.shape[1:], np.float32)], 0)
            eng.set_inputs(xb)
            eng.draw_noise(rate > 0, rate)                 # MC-dropout: both bottleneck Dropout sites are live
            eng.forward(training=False, dropout_rate=rate, branches=[0], need_l1=False)
            rec[i:i + n] = eng.br[0].xhat.cpu().numpy()[:n]
        results = {'reconstruction': rec}
        results['l1err'] = np.sum(np.abs(x - rec))
        results['l2err'] = np.sum(np.sqrt((x - rec) ** 2))
        return results
