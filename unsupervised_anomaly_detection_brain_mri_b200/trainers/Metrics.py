"""Dice / confusion metrics and the recursive best-Dice threshold search (mirror of reference trainers/Metrics.py).

The per-threshold work - ``predictions > t`` over the whole stacked volume and the three integer sums Dice needs -
runs on the GPU (``uad_threshold_counts``: float64 compare, int64 counts => bit-identical to numpy); the host only
replays the reference's recursion (Metrics.py:138-162) and argmax (:134) on those counts.  ROC / PRC stay on sklearn.
Plotting is out of scope."""
import ctypes

import numpy as np
import torch

from .. import abi


def xfrange(start, stop, step):
    i = 0
    while start + i * step < stop:
        yield start + i * step
        i += 1


def dice(P, G):
    psum = np.sum(P.flatten())
    gsum = np.sum(G.flatten())
    pgsum = np.sum(np.multiply(P.flatten(), G.flatten()))
    with np.errstate(divide='ignore', invalid='ignore'):
        return (2 * pgsum) / (psum + gsum)


def confusion_matrix(P, G):
    P, G = P.flatten().astype(bool), G.flatten().astype(bool)
    tp = np.sum(P & G)
    fp = np.sum(P & ~G)
    fn = np.sum(~P & G)
    tn = np.sum(~P & ~G)
    return tp, fp, tn, fn


def tpr(P, G):
    tp, fp, tn, fn = confusion_matrix(P, G)
    return tp / (tp + fn)


def fpr(P, G):
    tp, fp, tn, fn = confusion_matrix(P, G)
    return fp / (fp + tn)


def precision(P, G):
    tp, fp, tn, fn = confusion_matrix(P, G)
    return tp / (tp + fp)


def recall(P, G):
    return tpr(P, G)


def vd(P, G):
    tps = np.multiply(P.flatten(), G.flatten())
    return np.sum(np.abs(np.logical_xor(tps, G.flatten()))) / np.sum(G.flatten())


def compute_roc(predictions, labels, filename=None, plottitle="ROC Curve"):
    from sklearn.metrics import auc, roc_curve
    _fpr, _tpr, _ = roc_curve(labels.astype(int), predictions)
    return auc(_fpr, _tpr), _fpr, _tpr, _


def compute_prc(predictions, labels, filename=None, plottitle="Precision-Recall Curve"):
    from sklearn.metrics import average_precision_score, precision_recall_curve
    precisions, recalls, thresholds = precision_recall_curve(labels.astype(int), predictions)
    return average_precision_score(labels.astype(int), predictions), precisions, recalls, thresholds


class DeviceScorer:
    """Holds the stacked residual volume and labels on the GPU and answers (sum P*G, sum P, sum G) per threshold."""
    MAX_THR = 32

    def __init__(self, predictions, labels, device=None, allreduce=None):
        """allreduce: in-place sum over ranks of an int64 device tensor (data-parallel scoring shards by volume; the Dice
        counts are integers, so the reduced result - hence argmax - is identical for every world size)."""
        abi.lib()
        self.allreduce = allreduce
        device = device or f'cuda:{torch.cuda.current_device()}'
        if isinstance(predictions, torch.Tensor):
            self.diff = predictions.reshape(-1).to(device=device, dtype=torch.float32)
        else:
            p32 = np.ascontiguousarray(predictions, np.float32).reshape(-1)
            if not np.array_equal(p32.astype(np.float64), np.asarray(predictions, np.float64).reshape(-1)):
                raise ValueError('predictions are not exactly float32-representable; the bit-exact device compare '
                                 '((double)d_f32 > t) needs the float32 residuals the scoring path produces')
            self.diff = torch.from_numpy(p32).to(device)
        if isinstance(labels, torch.Tensor):
            self.label = (labels.reshape(-1) != 0).to(device=device, dtype=torch.uint8)
        else:
            self.label = torch.from_numpy(np.ascontiguousarray(np.asarray(labels).reshape(-1) != 0).astype(np.uint8)).to(device)
        self.n = self.diff.numel()
        self.counts = torch.zeros(3 * self.MAX_THR, dtype=torch.int64, device=device)

    def counts_for(self, thresholds, mask_out=None):
        out = []
        for i in range(0, len(thresholds), self.MAX_THR):
            chunk = [float(t) for t in thresholds[i:i + self.MAX_THR]]
            arr = (ctypes.c_double * len(chunk))(*chunk)
            abi.call('uad_threshold_counts', self.diff.data_ptr(), self.label.data_ptr(), self.n, arr, len(chunk),
                     self.counts.data_ptr(), None if (mask_out is None or i > 0) else mask_out.data_ptr(),
                     torch.cuda.current_stream().cuda_stream)
            if self.allreduce is not None:
                self.allreduce(self.counts[:3 * len(chunk)])
            c = self.counts[:3 * len(chunk)].cpu().numpy().reshape(-1, 3)
            out.extend((int(a), int(b), int(g)) for a, b, g in c)
        return out

    def dice_scores(self, thresholds):
        res = []
        for pg, p, g in self.counts_for(thresholds):
            with np.errstate(divide='ignore', invalid='ignore'):
                res.append(np.float64(2 * pg) / np.float64(p + g))     # numpy int64 sums then true division -> float64
        return res

    def threshold_mask(self, t):
        mask = torch.empty(self.n, dtype=torch.uint8, device=self.diff.device)
        self.counts_for([t], mask_out=mask)
        return mask


def compute_dice_score(predictions, labels, granularity, scorer=None):
    scorer = scorer or DeviceScorer(predictions, labels)

    def inner_compute_dice_curve_recursive(start, stop, decimal):
        _threshs, _scores = [], []
        had_recursion = False
        if decimal == granularity:
            return _threshs, _scores
        level = list(xfrange(start, stop, (1.0 / (10.0 ** decimal))))
        level_scores = scorer.dice_scores(level)                 # one device pass per refinement level
        for i, (t, score) in enumerate(zip(level, level_scores)):
            if i >= 2 and score <= _scores[i - 1] and not had_recursion:
                _subthreshs, _subscores = inner_compute_dice_curve_recursive(_threshs[i - 2], t, decimal + 1)
                _threshs.extend(_subthreshs)
                _scores.extend(_subscores)
                had_recursion = True
            _scores.append(score)
            _threshs.append(t)
        return _threshs, _scores

    threshs, scores = inner_compute_dice_curve_recursive(0, 1.0, 1)
    sorted_pairs = sorted(zip(threshs, scores))
    threshs, scores = list(zip(*sorted_pairs))
    return scores, threshs


def compute_dice_curve_recursive(predictions, labels, filename=None, plottitle="DICE Curve", granularity=5, scorer=None):
    scores, threshs = compute_dice_score(predictions, labels, granularity, scorer=scorer)
    bestthresh_idx = np.argmax(scores)
    return scores[bestthresh_idx], threshs[bestthresh_idx]


def combined_predictive_uncertainty(p, sigmas, axis=-1, log_var=False):
    if log_var:
        sigmas = np.exp(sigmas)
    return np.mean(np.square(p), axis=axis) - np.square(np.mean(p, axis=axis)) + np.mean(sigmas, axis=axis)
