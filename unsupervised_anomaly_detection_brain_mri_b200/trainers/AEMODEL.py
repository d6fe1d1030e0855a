"""AE base trainer (mirror of reference trainers/AEMODEL.py:12-79) driving the CUDA engine instead of a tf.Session."""
import os
from abc import ABC
from collections import defaultdict
from datetime import datetime
from math import inf

import numpy as np
import torch

from .. import abi
from ..engine import ConvAutoencoderEngine
from ..models.customlayers import Placeholder
from ..utils.logger import Logger, Phase
from . import trainer_utils
from .DLMODEL import DLMODEL


class AEMODEL(DLMODEL, ABC):
    class Config(DLMODEL.Config):
        def __init__(self, modelname='AE'):
            super().__init__()
            self.modelname = modelname
            self.intermediateResolutions = [8, 8]
            self.outputWidth = 256
            self.outputHeight = 256
            self.numChannels = 3
            self.dropout = False
            self.dropout_rate = 0.2
            self.zDim = 128

    TWO_INPUTS = False          # ceVAE feeds (x, x_ce)
    REC_KEY = 'x_hat'           # output key of the reconstruction (GMVAE: 'xz_mu')
    MASKED_INPUT = False        # CE feeds the patch-masked batch and scores against the plain one (one-input graph)
    LOSS_KEYS = ('loss',)

    def __init__(self, sess, config=None, network=None):
        super().__init__(sess, config if config is not None else self.Config())
        self.losses = {}
        self.network = network
        self.dropout = Placeholder([], 'dropout')
        self.dropout_rate = Placeholder([], 'dropout_rate')
        self.checkpointDir = os.path.join(self.config.checkpointDir or 'checkpoints', self.network.__name__)
        self.logDir = os.path.join(os.getcwd(), 'logs', self.network.__name__, self.model_dir, datetime.now().strftime('%Y%m%d_%H%M%S'))
        self.logger = Logger(self.sess, self.logDir, enabled=bool(getattr(self.config, 'useTensorboard', False)))
        cfg = self.config
        self.x = Placeholder([None, cfg.outputHeight, cfg.outputWidth, cfg.numChannels], 'x')
        if self.TWO_INPUTS:
            self.x_ce = Placeholder([None, cfg.outputHeight, cfg.outputWidth, cfg.numChannels], 'x_ce')
            self.outputs = self.network(self.x, self.x_ce, dropout_rate=self.dropout_rate, dropout=self.dropout, config=cfg)
        else:
            self.outputs = self.network(self.x, dropout_rate=self.dropout_rate, dropout=self.dropout, config=cfg)
        self.reconstruction = self.outputs[self.REC_KEY]
        self.graph = self.reconstruction.graph
        # device / data-parallel context (one process per GPU; world > 1 when launched under torchrun)
        self.device = getattr(cfg, 'device', None) or f'cuda:{int(os.environ.get("LOCAL_RANK", 0))}'
        self.math_mode = int(getattr(cfg, 'math_mode', abi.MATH_TC_3XTF32))
        self.world = 1
        self._allreduce = None
        torch.cuda.set_device(self.device)
        self.engine = ConvAutoencoderEngine(self.graph.arch, self.graph.S, self.graph.C, self.graph.zDim, self.graph.res,
                                            batch=cfg.batchsize, device=self.device, math_mode=self.math_mode,
                                            seed=int(getattr(cfg, 'seed', 1)),
                                            keep_preact=bool(getattr(cfg, 'keep_preact', False)), **self._engine_extra())
        self._eval_engines = {}
        self._pinned = {}
        self.get_number_of_trainable_params()
        self.saver = self           # reference attribute; save/load live on the trainer itself

    def _engine_extra(self):
        """Architecture-specific engine arguments (GMVAE: dim_w, dim_c, c_lambda)."""
        return {}

    # ------------------------------------------------------------------ data parallel
    def enable_data_parallel(self):
        """Shard mini-batches over ranks; one NCCL all-reduce on the flat gradient buffer per step."""
        from .. import dist as udist
        udist.init_from_env()
        self.world = udist.world_size()
        if self.world > 1:
            udist.broadcast_(self.engine.fp.params, src=0)
            self._allreduce = udist.allreduce_sum_
            # one fused kernel over NVLink peer memory (reduce-scatter + Adam + all-gather) instead of NCCL all-reduce + Adam;
            # UAD_PEER_ADAM=0 keeps the NCCL form
            if os.environ.get('UAD_PEER_ADAM', '1') != '0' and udist.dist.get_backend() == 'nccl' and hasattr(self.engine, 'enable_peer_optimizer'):
                self.engine.enable_peer_optimizer()

    # ------------------------------------------------------------------ reference helpers
    def log_to_tensorboard(self, epoch, scalars, visuals, phase: Phase, name='x'):
        for key in scalars.keys():
            scalars[key] = np.mean(scalars[key])
        vis = [v for v in visuals if v is not None]
        summaries = dict(scalars)
        if vis:
            summaries[name] = np.vstack(vis)[:50]
        self.logger.summarize(epoch, phase=phase, summaries_dict=summaries)

    def load_checkpoint(self):
        could_load, checkpoint_counter = self.load(self.checkpointDir)
        if could_load:
            last_epoch = checkpoint_counter
            print(" [*] Load SUCCESS")
        else:
            last_epoch = 0
            print(" [!] Load failed...")
        return last_epoch

    @property
    def model_dir(self):
        return "{}_d{}_s{}x{}_{}_b{}_z{}_{}".format(self.config.modelname, self.config.dataset, self.config.outputWidth,
                                                    self.config.outputHeight, self.network.__name__, self.config.batchsize,
                                                    self.config.zDim, self.config.description)

    # ------------------------------------------------------------------ the step (replaces sess.run)
    def _stage(self, key, arr):
        """Host batch -> pinned staging buffer -> device input buffer (async H2D on the compute stream)."""
        arr = np.ascontiguousarray(arr, np.float32)
        buf = self._pinned.get((key, arr.shape))
        if buf is None:
            buf = torch.empty(arr.shape, dtype=torch.float32).pin_memory()
            self._pinned[(key, arr.shape)] = buf
        buf.numpy()[...] = arr
        return buf

    def _prefetch(self, key, arr):
        """Start the host->device copy of a FUTURE batch on a side stream (pinned staging, double-buffered) so it overlaps
        the step that is about to run; run_batch() picks it up by identity of the host array."""
        if arr is None:
            return
        if not hasattr(self, '_copy_stream'):
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._pf = {}
            self._pf_flip = {}
        flip = self._pf_flip.get(key, 0)
        self._pf_flip[key] = flip ^ 1
        pinned = self._stage((key, flip), arr)
        dev = self._pinned.get(('dev', key, flip, pinned.shape))
        if dev is None:
            dev = torch.empty(pinned.shape, dtype=torch.float32, device=self.device)
            self._pinned[('dev', key, flip, pinned.shape)] = dev
        ev = torch.cuda.Event()
        with torch.cuda.stream(self._copy_stream):
            dev.copy_(pinned, non_blocking=True)
            ev.record()
        self._pf[key] = (id(arr), dev, ev)

    def _fetch_pinned(self, key, dev):
        """Start the device->host copy of a result tensor into a pinned buffer (two per key, alternating) on the compute stream and
        return the numpy view; the caller synchronises (the scalar-loss read does) before looking at it."""
        flip = self._pf_flip_out.get(key, 0) if hasattr(self, '_pf_flip_out') else 0
        if not hasattr(self, '_pf_flip_out'):
            self._pf_flip_out = {}
        self._pf_flip_out[key] = flip ^ 1
        buf = self._pinned.get(('out', key, flip, tuple(dev.shape)))
        if buf is None:
            buf = torch.empty(tuple(dev.shape), dtype=dev.dtype).pin_memory()
            self._pinned[('out', key, flip, tuple(dev.shape))] = buf
        buf.copy_(dev, non_blocking=True)
        return buf.numpy()

    def _feed(self, key, arr):
        """Device-side source for this batch: the prefetched copy if one was started for exactly this array, else a
        pinned host buffer (async H2D on the compute stream)."""
        pf = getattr(self, '_pf', {}).pop(key, None)
        if pf is not None and pf[0] == id(arr):
            torch.cuda.current_stream(self.device).wait_event(pf[2])
            return pf[1]
        return self._stage(key, arr)

    def run_batch(self, batch, phase: Phase, batch_ce=None, fetch_maps=False, want_anomaly=False, prefetch=None,
                  prefetch_ce=None):
        """One ``sess.run`` of the reference's process() loop (AE.py:70-83): feed a host batch, run the step on the GPU,
        fetch the scalar losses (and, on request, the reconstruction / L1 maps the reference fetches every step).
        ``prefetch`` (optional): the NEXT host batch - its H2D copy is overlapped with this step."""
        eng, cfg = self.engine, self.config
        eng.set_inputs(self._feed('x', batch), None if batch_ce is None else self._feed('x_ce', batch_ce))
        if phase == Phase.TRAIN:
            eng.train_step(cfg.learningrate, beta1=cfg.beta1, dropout_rate=cfg.dropout_rate, dropout=True,
                           allreduce=self._allreduce, world=self.world, want_anomaly=want_anomaly,
                           use_graph=bool(getattr(cfg, 'use_cuda_graph', True)))
        else:
            eng.draw_noise(False, 0.0)
            eng.forward(training=False, dropout_rate=0.0)
        self._prefetch('x', prefetch)                 # the GPU is busy with the step: stage the next batch meanwhile
        self._prefetch('x_ce', prefetch_ce)
        maps = None
        if fetch_maps:                                # pinned, asynchronous D2H behind the step; the scalar read below waits for all
            maps = {k: self._fetch_pinned(k, t) for k, t in (('reconstruction', eng.br[0].xhat), ('L1', eng.br[0].l1))}
        run = dict(eng.losses())                      # device -> host read of the step's scalars
        if maps is not None:
            run.update(maps)                          # views of a two-deep ring of pinned buffers: valid until the second-next fetch
        run = {k: (np.float32(v) if np.ndim(v) == 0 else v) for k, v in run.items()}
        return run

    def _eval_engine(self, n):
        n = int(n)
        if n not in self._eval_engines:
            g = self.graph
            self._eval_engines[n] = ConvAutoencoderEngine(g.arch, g.S, g.C, g.zDim, g.res, batch=n, device=self.device,
                                                          math_mode=self.math_mode, share_params=self.engine.fp, **self._engine_extra())
        return self._eval_engines[n]

    def reconstruct(self, x, dropout=False):
        """Forward-only pass (AE.py:92-110).  Accepts [H,W,C] or a whole stack [N,H,W,C] (batched on the device)."""
        if x.ndim < 4:
            x = np.expand_dims(x, 0)
        x = np.ascontiguousarray(x, np.float32)
        N = x.shape[0]
        chunk = min(N, int(getattr(self.config, 'evalBatchsize', 128)))
        eng = self._eval_engine(chunk)
        rec = np.empty_like(x)
        rate = self.config.dropout_rate if dropout else 0.0
        for i in range(0, N, chunk):
            xb = x[i:i + chunk]
            n = xb.shape[0]
            if n < chunk:
                xb = np.concatenate([xb, np.zeros((chunk - n,) + xb.shape[1:], np.float32)], 0)
            eng.set_inputs(xb, xb if self.TWO_INPUTS else None)
            eng._keep = 1.0 / (1.0 - rate) if rate > 0 else 1.0
            eng.draw_noise(dropout, rate)
            eng.forward(training=False, dropout_rate=rate, branches=[0], need_l1=False)
            rec[i:i + n] = eng.br[0].xhat.cpu().numpy()[:n]
        results = {'reconstruction': rec}
        results['l1err'] = np.sum(np.abs(x - rec))
        results['l2err'] = np.sum(np.sqrt((x - rec) ** 2))
        return results

    # ------------------------------------------------------------------ epoch loops (AE.py:23-90)
    def _make_ce_batch(self, batch, brainmasks, phase):
        return None

    def train(self, dataset):
        self.variables = list(self.engine.specs.keys())
        self.optim = self.create_optimizer(None, var_list=self.variables, learningrate=self.config.learningrate,
                                           beta1=self.config.beta1, type=self.config.optimizer)
        best_cost = inf
        last_improvement = 0
        last_epoch = self.load_checkpoint()
        from .. import dist as udist
        for epoch in range(last_epoch, self.config.numEpochs):
            self.process(dataset, epoch, Phase.TRAIN, self.optim)
            last_epoch += 1
            if udist.rank() == 0:                          # data parallel: identical weights on every rank, one writer
                self.save(self.checkpointDir, last_epoch)
            val_scalars = self.process(dataset, epoch, Phase.VAL)
            # data parallel: every rank must take the same early-stopping decision (a rank that leaves the loop alone would
            # leave the others blocked in the gradient all-reduce) - decide on the mean validation loss over the ranks
            val_loss = val_scalars['loss']
            if self.world > 1:
                val_loss = udist.mean_over_ranks(float(np.mean(val_loss)), self.device)
            best_cost, last_improvement, stop = indicate_early_stopping(val_loss, best_cost, last_improvement)
            if stop:
                print('Early stopping was triggered due to no improvement over the last 5 epochs')
                break

    def process(self, dataset, epoch, phase: Phase, optim=None, visualization_keys=None):
        scalars = defaultdict(list)
        visuals = []
        num_batches = dataset.num_batches(self.config.batchsize, set=phase.value)
        if self.world > 1:                                 # every rank steps the same number of times (one all-reduce per step)
            from .. import dist as udist
            num_batches = int(-udist.max_over_ranks(-num_batches, self.device))
        every = int(getattr(self.config, 'fetchMapsEvery', 0))      # the reference fetches the full maps EVERY step
        verbose = bool(getattr(self.config, 'verbose', True))
        def fetch():
            if self.TWO_INPUTS or self.MASKED_INPUT:
                b, _, masks = dataset.next_batch(self.config.batchsize, return_brainmask=True, set=phase.value)
                return b, self._make_ce_batch(b, masks, phase)
            b, _, _ = dataset.next_batch(self.config.batchsize, set=phase.value)
            return b, None
        nxt = fetch() if num_batches > 0 else None
        for idx in range(0, num_batches):
            batch, batch_ce = nxt
            nxt = fetch() if idx + 1 < num_batches else (None, None)      # look-ahead: its H2D overlaps this step
            fetch_m = every > 0 and idx % every == 0
            run = self.run_batch(batch, phase, batch_ce=batch_ce, fetch_maps=fetch_m, want_anomaly=self.TWO_INPUTS,
                                 prefetch=nxt[0], prefetch_ce=nxt[1])
            if verbose:
                print(f'Epoch ({phase.value}): [{epoch:2d}] [{idx:4d}/{num_batches:4d}] loss: {run["loss"]:.8f}')
            update_log_dicts(*trainer_utils.get_summary_dict(batch, run, visualization_keys), scalars, visuals)
        self.log_to_tensorboard(epoch, scalars, visuals, phase)
        return scalars


def update_log_dicts(scalars, visuals, train_scalars, train_visuals):
    for k, v in list(scalars.items()):
        train_scalars[k].append(v)
    train_visuals.append(visuals)


def indicate_early_stopping(current_cost, best_cost, last_improvement):
    if current_cost < best_cost:
        best_cost = current_cost
        last_improvement = 0
        return best_cost, last_improvement, False
    else:
        last_improvement += 1
        if last_improvement >= 5:
            return best_cost, last_improvement, True
        return best_cost, last_improvement, False
