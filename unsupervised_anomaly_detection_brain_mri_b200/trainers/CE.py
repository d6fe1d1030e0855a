"""Context-encoder trainer and its patch masking (mirror of reference trainers/CE.py).

``CE`` trains the one-input autoencoder graph on the patch-masked batch against the PLAIN batch (CE.py:21,34,83-90): the
engine's reconstruction target is decoupled from its input (engine.set_target).  Validation and ``reconstruct`` feed the
plain batch on both sides."""
import random

import numpy as np

from ..utils.logger import Phase
from .AEMODEL import AEMODEL


def retrieve_masked_batch(batch, brainmasks):
    """1-3 random 20x20 patches inside each sample's brain bounding box are zeroed.

    Reference quirk preserved (CE.py:130-138): the loop variable shadows the mask array, so the mask that is
    finally multiplied is the LAST sample's [H,W,C] mask, broadcast over the whole batch."""
    def retrieve_brain_range(brainmask):
        pixels = np.argwhere(brainmask).T
        return (min(pixels[0]), max(pixels[0])), (min(pixels[1]), max(pixels[1]))

    brain_ranges = [retrieve_brain_range(bm) for bm in brainmasks]
    masks = np.ones(batch.shape)
    last = masks[0]
    for sample_mask, brain_range in zip(masks, brain_ranges):
        last = sample_mask
        for _ in range(random.randint(1, 3)):
            size_w, size_h = 20, 20
            if brain_range[0][0] < brain_range[0][1] - size_w and brain_range[1][0] < brain_range[1][1] - size_h:
                px = random.randint(brain_range[0][0], brain_range[0][1] - size_w)
                py = random.randint(brain_range[1][0], brain_range[1][1] - size_h)
                sample_mask[px:px + size_w, py:py + size_h] = 0
    return (batch * last).astype(batch.dtype)


class CE(AEMODEL):
    MASKED_INPUT = True

    class Config(AEMODEL.Config):
        def __init__(self):
            super().__init__('CE')

    def _make_ce_batch(self, batch, brainmasks, phase):
        return retrieve_masked_batch(batch, brainmasks) if phase == Phase.TRAIN else batch      # CE.py:88

    def run_batch(self, batch, phase, batch_ce=None, fetch_maps=False, want_anomaly=False, prefetch=None, prefetch_ce=None):
        eng, cfg = self.engine, self.config
        eng.set_target(self._stage('x', batch))
        eng.set_inputs(self._stage('x_ce', batch if batch_ce is None else batch_ce))
        if phase == Phase.TRAIN:
            eng.train_step(cfg.learningrate, beta1=cfg.beta1, dropout_rate=cfg.dropout_rate, dropout=True, allreduce=self._allreduce,
                           world=self.world, use_graph=bool(getattr(cfg, 'use_cuda_graph', True)))
        else:
            eng.draw_noise(False, 0.0)
            eng.forward(training=False, dropout_rate=0.0)
        run = dict(eng.losses())
        if fetch_maps:
            run['reconstruction'] = eng.br[0].xhat.cpu().numpy()
            run['L1'] = eng.br[0].l1.cpu().numpy()
        return {k: (np.float32(v) if np.ndim(v) == 0 else v) for k, v in run.items()}
