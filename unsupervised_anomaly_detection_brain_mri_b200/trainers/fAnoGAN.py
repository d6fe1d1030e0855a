"""f-AnoGAN trainer (mirror of reference trainers/fAnoGAN.py) on the CUDA engine.

``train``: phase A = WGAN-GP (per mini-batch one generator step and ``d_iters = 5`` critic steps, fAnoGAN.py:87-140), phase B
= encoder training (izi_f) with validation and early stopping (fAnoGAN.py:142-210); ``reconstruct`` = sigmoid(G(E(x)))
(fAnoGAN.py:220-239).  The three optimisers are tf.train.AdamOptimizer(lr, beta1=0.5, beta2=0.9) on the scope-contiguous
slices of the flat parameter buffer (fAnoGAN.py:71-77)."""
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

    ENGINE = FanoganEngine          # what a sibling trainer on the same stacks overrides (trainers/AnoVAEGAN.py)
    REC_KEY = 'x_enc'

    def _network_outputs(self):
        return self.network(self.z, self.x, dropout_rate=self.dropout_rate, dropout=self.dropout, config=self.config)

    def _engine_kwargs(self):
        return dict(kappa=float(self.config.kappa), scale=float(self.config.scale))

    def __init__(self, sess, config=None, network=None):
        super().__init__(sess, config if config is not None else self.Config())
        self.losses = {}
        self.network = network
        cfg = self.config
        self.x = Placeholder([None, cfg.outputHeight, cfg.outputWidth, cfg.numChannels], 'x')
        self.z = Placeholder([None, cfg.zDim], 'z')
        self.dropout = Placeholder([], 'dropout')
        self.dropout_rate = Placeholder([], 'dropout_rate')
        self.outputs = self._network_outputs()
        self.reconstruction = self.outputs[self.REC_KEY]
        self.generated = self.outputs.get('x_')
        self.graph = self.reconstruction.graph
        self.checkpointDir = os.path.join(cfg.checkpointDir or 'checkpoints', self.network.__name__)
        self.logDir = os.path.join(os.getcwd(), 'logs', self.network.__name__, self.model_dir, datetime.now().strftime('%Y%m%d_%H%M%S'))
        self.device = getattr(cfg, 'device', None) or f'cuda:{int(os.environ.get("LOCAL_RANK", 0))}'
        self.math_mode = int(getattr(cfg, 'math_mode', abi.MATH_TC_3XTF32))
        torch.cuda.set_device(self.device)
        g = self.graph
        self.engine = self.ENGINE(g.S, g.C, g.zDim, g.res, batch=cfg.batchsize, device=self.device, math_mode=self.math_mode,
                                  seed=int(getattr(cfg, 'seed', 1)), **self._engine_kwargs())
        self._eval = {}
        self.world, self._allreduce = 1, None
        self.logger = Logger(self.sess, self.logDir, enabled=bool(getattr(cfg, 'useTensorboard', False)))
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

    def sample_z(self, batch_size=None):
        return np.random.normal(size=[batch_size if batch_size else self.config.batchsize, self.config.zDim])   # fAnoGAN.py:241-242

    def enable_data_parallel(self):
        """Shard mini-batches over ranks; one all-reduce of the updated scope's gradient slice per train op."""
        from .. import dist as udist
        udist.init_from_env()
        self.world = udist.world_size()
        if self.world > 1:
            udist.broadcast_(self.engine.fp.params, src=0)
            self._allreduce = udist.allreduce_sum_
            # fused reduce-scatter + Adam + all-gather over NVLink peer memory per train op (UAD_PEER_ADAM=0: NCCL + Adam);
            # the subclasses with per-optimiser Adam slots (AnoVAEGAN) keep the NCCL form
            if (os.environ.get('UAD_PEER_ADAM', '1') != '0' and udist.dist.get_backend() == 'nccl' and type(self.engine).__name__ == 'FanoganEngine'):
                self.engine.enable_peer_optimizer()

    def _feed(self, batch):
        """get_feed_dict (fAnoGAN.py:212-218): x <- batch, z <- sample_z()."""
        self.engine.set_inputs(np.ascontiguousarray(batch, np.float32))
        self._z = self.sample_z()
        self.engine.set_latent(self._z.astype(np.float32))

    def train(self, dataset):
        from collections import defaultdict
        from math import inf

        from . import trainer_utils
        from .AEMODEL import indicate_early_stopping, update_log_dicts
        cfg, eng = self.config, self.engine
        eng.kappa, eng.scale = float(cfg.kappa), float(cfg.scale)
        eng.enable_training()
        self.variables = list(eng.specs.keys())
        lr, rate = float(cfg.learningrate), float(cfg.dropout_rate)
        graphs = bool(getattr(cfg, 'useCudaGraph', True))                 # replay each train op as one CUDA graph launch
        kw = dict(dropout_rate=rate, dropout=True, allreduce=getattr(self, '_allreduce', None), world=getattr(self, 'world', 1),
                  use_graph=graphs)
        verbose = bool(getattr(cfg, 'verbose', True))
        all_losses = bool(getattr(cfg, 'fetchAllLosses', True))
        best_cost = inf
        last_improvement = 0
        last_epoch = self.load_checkpoint()

        for epoch in range(last_epoch, cfg.numEpochs):                    # ---- TRAINING WGAN (fAnoGAN.py:87-140)
            phase = Phase.TRAIN
            scalars, visuals = defaultdict(list), []
            d_iters = 5
            num_batches = dataset.num_batches(cfg.batchsize, set=phase.value)
            for idx in range(num_batches):
                batch, _, _ = dataset.next_batch(cfg.batchsize, set=phase.value)
                self._feed(batch)
                run = dict(eng.step_gen(lr, **kw))
                for _ in range(d_iters):
                    self._feed(batch)                                     # every sess.run draws a fresh z (fAnoGAN.py:125)
                    run.update(eng.step_disc(lr, **kw))
                run['generated'] = eng.x_gen.cpu().numpy()
                if verbose:
                    print(f'Epoch ({phase.value} WGAN): [{epoch:2d}] [{idx:4d}/{num_batches:4d}]'
                          f' gen_loss: {run["gen_loss"]:.8f}, disc_loss: {run["disc_loss"]:.8f}')
                update_log_dicts(*trainer_utils.get_summary_dict(batch, run, visualization_keys=['generated']), scalars, visuals)
            self.log_to_tensorboard(epoch, scalars, visuals, phase, name='wgan_x')
            last_epoch += 1
            self.save(self.checkpointDir, last_epoch)

        def encoder_pass(phase, epoch):
            scalars, visuals = defaultdict(list), []
            num_batches = dataset.num_batches(cfg.batchsize, set=phase.value)
            for idx in range(num_batches):
                batch, _, _ = dataset.next_batch(cfg.batchsize, set=phase.value)
                self._feed(batch)
                train = phase == Phase.TRAIN
                run = dict(eng.step_enc(lr, dropout_rate=rate, dropout=train, allreduce=kw['allreduce'], world=kw['world'],
                                        train=train, use_graph=graphs))
                run['reconstruction'] = eng.x_enc.cpu().numpy()
                run['L1'] = eng.l1.cpu().numpy()
                if train:
                    run['z_enc'], run['z'] = eng.z_enc.cpu().numpy(), self._z
                if all_losses:                                            # **self.losses (fAnoGAN.py:157,190)
                    run.update(eng.wgan_scalars(rate, dropout=train, use_graph=graphs))
                if verbose:
                    tag = f'{phase.value} Encoder' if train else phase.value
                    print(f'Epoch ({tag}): [{epoch:2d}] [{idx:4d}/{num_batches:4d}]  reconstructionLoss: '
                          f'{run["reconstructionLoss"]:.8f}')
                update_log_dicts(*trainer_utils.get_summary_dict(batch, run), scalars, visuals)
            self.log_to_tensorboard(epoch, scalars, visuals, phase)
            return scalars

        for epoch in range(last_epoch, 2 * cfg.numEpochs):                # ---- TRAINING / VALIDATION Encoder (:142-210)
            encoder_pass(Phase.TRAIN, epoch)
            last_epoch += 1
            self.save(self.checkpointDir, last_epoch)
            val = encoder_pass(Phase.VAL, epoch)
            best_cost, last_improvement, stop = indicate_early_stopping(val['reconstructionLoss'], best_cost, last_improvement)
            if stop:
                print('Early stopping was triggered due to no improvement over the last 5 epochs')
                break

    def log_to_tensorboard(self, epoch, scalars, visuals, phase, name='x'):
        for key in scalars.keys():
            scalars[key] = np.mean(scalars[key])
        vis = [v for v in visuals if v is not None]
        summaries = dict(scalars)
        if vis:
            summaries[name] = np.vstack(vis)[:50]
        self.logger.summarize(epoch, phase=phase, summaries_dict=summaries)

    def _engine_for(self, n):
        if n not in self._eval:
            g = self.graph
            e = self.ENGINE(g.S, g.C, g.zDim, g.res, batch=n, device=self.device, math_mode=self.math_mode, **self._engine_kwargs())
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
