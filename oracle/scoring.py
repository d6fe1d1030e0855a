"""numpy restatement of the residual-map anomaly scoring (TEST INFRASTRUCTURE ONLY; PARITY UNPINNED).

Follows utils/Evaluation.py:84-89 (brain-mask erosion), :282-291 (residual, mask, hyper-intensity prior, float64
sub-volume), :453-457 (threshold mask) and trainers/Metrics.py:10-14,67-72,138-162,134 (Dice, recursive best-Dice
threshold search, argmax).  numpy semantics are restated literally (dtype promotion included) because the masks and
the argmax must be bit-exact.
"""
from __future__ import annotations

import numpy as np
import scipy.ndimage


def erode_brainmask(mask2d, iterations=12):
    """Evaluation.py:84-89: 4-neighbour cross structuring element, border_value=0."""
    strel = scipy.ndimage.generate_binary_structure(2, 1)
    return scipy.ndimage.binary_erosion(np.squeeze(mask2d), structure=strel, iterations=iterations)


def residual(x, x_rec, mask, prior_quantile, keep_positive=True, apply_prior=True):
    """Evaluation.py:282-291 for a stack of slices.

    x, x_rec: float32 [N,H,W]; mask: bool/int [N,H,W] (already eroded if wanted); prior_quantile: python float (f64).
    Returns the float64 sub-volume whose values are exactly float32-representable."""
    x = x.astype(np.float32)
    x_rec = x_rec.astype(np.float32)
    if keep_positive:
        d = np.maximum(x - x_rec, 0)
    else:
        d = np.abs(x - x_rec)
    d = np.multiply(mask.astype(bool), d)                 # stays float32
    if apply_prior:
        d = d.copy()
        d[x < prior_quantile] = 0                         # float32 < float64 compare -> promoted to f64
    sub = np.zeros(d.shape, np.float64)
    sub[...] = d
    return sub


def threshold_mask(diffs, t):
    """Evaluation.py:453-457: ``diffs > t`` with diffs float64 and t a python float."""
    return diffs > t


def dice(P, G):
    """Metrics.py:67-72."""
    psum = np.sum(P.flatten())
    gsum = np.sum(G.flatten())
    pgsum = np.sum(np.multiply(P.flatten(), G.flatten()))
    with np.errstate(divide='ignore', invalid='ignore'):
        return (2 * pgsum) / (psum + gsum)


def counts(diffs, labels, t):
    """(sum P*G, sum P, sum G) as python ints for one threshold."""
    P = np.where(diffs > t, 1, 0)
    return int(np.sum(P * labels)), int(np.sum(P)), int(np.sum(labels))


def xfrange(start, stop, step):
    i = 0
    while start + i * step < stop:
        yield start + i * step
        i += 1


def best_dice_search(predictions, labels, granularity=10):
    """Metrics.py:138-162 + :134: decade-refinement search; returns (best_score, best_threshold, all_threshs, all_scores)."""

    def level(start, stop, decimal):
        th, sc = [], []
        recursed = False
        if decimal == granularity:
            return th, sc
        for i, t in enumerate(xfrange(start, stop, (1.0 / (10.0 ** decimal)))):
            score = dice(np.where(predictions > t, 1, 0), labels)
            if i >= 2 and score <= sc[i - 1] and not recursed:
                sth, ssc = level(th[i - 2], t, decimal + 1)
                th.extend(sth)
                sc.extend(ssc)
                recursed = True
            sc.append(score)
            th.append(t)
        return th, sc

    th, sc = level(0, 1.0, 1)
    pairs = sorted(zip(th, sc))
    th, sc = list(zip(*pairs))
    idx = int(np.argmax(sc))
    return sc[idx], th[idx], th, sc
