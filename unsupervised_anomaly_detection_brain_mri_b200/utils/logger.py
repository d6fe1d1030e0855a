"""Phase enum + scalar/image summaries (mirror of reference utils/logger.py:8-60) on torch's TensorBoard writer."""
import os
from enum import Enum

import numpy as np


class Phase(Enum):
    TRAIN = 'TRAIN'
    VAL = 'VAL'
    TEST = 'TEST'


class Logger:
    def __init__(self, sess, summary_dir, enabled=True):
        self.summary_dir = summary_dir
        self.enabled = enabled
        self._writers = {}

    def _writer(self, phase):
        if phase not in self._writers:
            from torch.utils.tensorboard import SummaryWriter
            self._writers[phase] = SummaryWriter(os.path.join(self.summary_dir, phase.value))
        return self._writers[phase]

    def summarize(self, step, phase: Phase = Phase.TRAIN, scope='', summaries_dict=None):
        if not self.enabled or summaries_dict is None:
            return
        if not isinstance(phase, Phase):
            raise ValueError(f'Illegal Argument for summarizer: {phase}')
        w = self._writer(phase)
        for tag, value in summaries_dict.items():
            if value is None:
                continue
            value = np.asarray(value)
            if value.ndim <= 1:
                w.add_scalar(tag, float(value.reshape(-1)[0]), step)
            else:
                img = value.astype(np.float32)
                if img.ndim == 3:
                    img = img[..., None]
                w.add_images(tag, np.clip(img / 255.0, 0, 1), step, dataformats='NHWC')
        w.flush()
