"""Scalar / image-strip extraction from one step's results (mirror of reference trainers/trainer_utils.py:6-18)."""
import numpy as np


def normalize(x):
    """min-max to [0,1] (reference utils/utils.py:74-75 uses cv2.normalize NORM_MINMAX)."""
    x = np.asarray(x, np.float32)
    lo, hi = float(x.min()), float(x.max())
    return (x - lo) / (hi - lo) if hi > lo else np.zeros_like(x)


def get_summary_dict(batch, run, visualization_keys=None, *others):
    if visualization_keys is None:
        visualization_keys = ['reconstruction', 'L1']
    scalars = {k: v for k, v in run.items() if np.ndim(v) == 0 and v is not None and not (isinstance(v, float) and v != v)}
    visuals = None
    if all(k in run and run[k] is not None for k in visualization_keys):
        visuals = np.asarray([255 * np.hstack([normalize(batch[i]), *[normalize(run[key][i]) for key in visualization_keys],
                                               *[normalize(e[i]) for e in others]]) for i in range(len(batch))])
    return scalars, visuals
