"""AnoVAEGAN trainer (mirror of reference trainers/AnoVAEGAN.py) on the CUDA engine.

Per mini-batch (AnoVAEGAN.py:94-149): one optim_vae step (Encoder + Generator on L1 + kl_weight * KL), one optim_gen step
(Generator on -mean(D(out))) and ``d_iters = 5`` optim_dis steps (WGAN-GP critic), every sess.run drawing fresh N(0,1) /
dropout / interpolation noise; then a validation pass with early stopping on the reconstruction loss (:163-192).
``reconstruct`` = out with a fresh eps (the graph's tf.random_normal is live at inference too, anovaegan.py:35).

STATUS: the engine's call sequences are verified on CPU against the oracle (tests/test_engine_emulated.py); the first run on
hardware is still to come (tests/test_gpu_anovaegan.py, opt-in)."""
from collections import defaultdict
from math import inf

import numpy as np
import torch

from .. import abi
from ..anovaegan_engine import AnoVaeGanEngine
from ..utils.logger import Phase
from . import trainer_utils
from .AEMODEL import AEMODEL, indicate_early_stopping, update_log_dicts
from .fAnoGAN import fAnoGAN


class AnoVAEGAN(fAnoGAN):
    ENGINE = AnoVaeGanEngine
    REC_KEY = 'out'

    class Config(AEMODEL.Config):
        def __init__(self):
            super().__init__('AnoVAEGAN')
            self.scale = 10.0
            self.kappa = 1.0
            self.kl_weight = 1.0

    def _network_outputs(self):
        return self.network(self.x, dropout_rate=self.dropout_rate, dropout=self.dropout, config=self.config)

    def _engine_kwargs(self):
        return dict(kl_weight=float(self.config.kl_weight), scale=float(self.config.scale))

    def _feed(self, batch):
        self.engine.set_inputs(np.ascontiguousarray(batch, np.float32))

    def train(self, dataset):
        cfg, eng = self.config, self.engine
        eng.kl_weight, eng.scale = float(cfg.kl_weight), float(cfg.scale)
        eng.enable_training()
        self.variables = list(eng.specs.keys())
        lr, rate = float(cfg.learningrate), float(cfg.dropout_rate)
        graphs = bool(getattr(cfg, 'useCudaGraph', True))
        kw = dict(dropout_rate=rate, allreduce=getattr(self, '_allreduce', None), world=getattr(self, 'world', 1), use_graph=graphs)
        verbose = bool(getattr(cfg, 'verbose', True))
        best_cost = inf
        last_improvement = 0
        last_epoch = self.load_checkpoint()
        for epoch in range(last_epoch, cfg.numEpochs):
            phase = Phase.TRAIN
            scalars, visuals = defaultdict(list), []
            d_iters = 5
            num_batches = dataset.num_batches(cfg.batchsize, set=phase.value)
            for idx in range(num_batches):
                batch, _, _ = dataset.next_batch(cfg.batchsize, set=phase.value)
                self._feed(batch)
                run = dict(eng.step_vae(lr, dropout=True, **kw))
                run['reconstruction'] = eng.x_gen.cpu().numpy()
                run['L1'] = eng.l1.cpu().numpy()
                run.update(eng.step_gen(lr, dropout=True, **kw))
                for _ in range(d_iters):
                    run.update(eng.step_disc(lr, dropout=True, **kw))
                if verbose:
                    print(f'Epoch ({phase.value}): [{epoch:2d}] [{idx:4d}/{num_batches:4d}] gen_loss: {run["gen_loss"]:.8f}, '
                          f'disc_loss: {run["disc_loss"]:.8f}, reconstructionLoss: {run["reconstructionLoss"]:.8f}')
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
                run = dict(eng.step_vae(lr, dropout=False, train=False, **kw))
                run['reconstruction'] = eng.x_gen.cpu().numpy()
                run['L1'] = eng.l1.cpu().numpy()
                if verbose:
                    print(f'Epoch ({phase.value}): [{epoch:2d}] [{idx:4d}/{num_batches:4d}] reconstructionLoss: '
                          f'{run["reconstructionLoss"]:.8f}')
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
        keep = 1.0 / (1.0 - rate) if rate > 0 else 1.0
        st = torch.cuda.current_stream().cuda_stream
        for i in range(0, N, chunk):
            xb = x[i:i + chunk]
            n = xb.shape[0]
            if n < chunk:
                xb = np.concatenate([xb, np.zeros((chunk - n,) + xb.shape[1:], np.float32)], 0)
            eng.set_inputs(xb)
            seed = int(np.random.randint(0, 2 ** 31))
            abi.call('uad_randn', eng.eps.data_ptr(), eng.eps.numel(), seed, 0, None, st)
            masks = (None, None, None)
            if rate > 0:                                   # MC-dropout: all three Dropout applications of the graph are live
                masks = (torch.empty(chunk, eng.zDim, device=eng.device), torch.empty(chunk, eng.zDim, device=eng.device),
                         torch.empty(chunk, eng.flat, device=eng.device))
                for j, m in enumerate(masks):
                    abi.call('uad_dropout_mask', m.data_ptr(), m.numel(), float(rate), seed, (j + 1) << 40, None, st)
            rec[i:i + n] = eng.reconstruct(masks, keep).cpu().numpy()[:n]
        results = {'reconstruction': rec}
        results['l1err'] = np.sum(np.abs(x - rec))
        results['l2err'] = np.sum(np.sqrt((x - rec) ** 2))
        return results
