"""float64 naive-loop micro-oracles (TEST INFRASTRUCTURE ONLY; PARITY UNPINNED - see oracle/__init__.py).

These are the ground truth that pins ``tf_graph_cpu``'s use of torch conv primitives to the TF-1.15 SAME
semantics of SURVEY.md Appendix A.  Pure numpy loops over taps (vectorised over batch/space) - small sizes only.
"""
from __future__ import annotations

import math

import numpy as np


def conv2d_same_s2(x, w, b):
    """y[b,i,j,co] = sum_{kh,kw,ci} x[b,2i+kh-lo,2j+kw-lo,ci] w[kh,kw,ci,co] + b  with TF SAME padding
    (models/customlayers.py:21; SURVEY A.1: k=5,s=2, even input -> pad 1 before / 2 after)."""
    x = x.astype(np.float64)
    w = w.astype(np.float64)
    B, H, W, Ci = x.shape
    k = w.shape[0]
    Ho, Wo = -(-H // 2), -(-W // 2)
    pt = max((Ho - 1) * 2 + k - H, 0)
    lo = pt // 2
    xp = np.zeros((B, H + pt, W + pt, Ci))
    xp[:, lo:lo + H, lo:lo + W] = x
    y = np.zeros((B, Ho, Wo, w.shape[3]))
    for kh in range(k):
        for kw in range(k):
            patch = xp[:, kh:kh + 2 * Ho:2, kw:kw + 2 * Wo:2, :]
            y += patch @ w[kh, kw]
    return y + b.astype(np.float64)


def conv2dT_same_s2(x, K, b):
    """out[b,2i+kh-1,2j+kw-1,co] += x[b,i,j,ci] K[kh,kw,co,ci], full (2n+3) output cropped 1 / 2
    (models/customlayers.py:34)."""
    x = x.astype(np.float64)
    K = K.astype(np.float64)
    B, H, W, Ci = x.shape
    k = K.shape[0]
    Co = K.shape[2]
    full = np.zeros((B, 2 * H + k - 2, 2 * W + k - 2, Co))
    for kh in range(k):
        for kw in range(k):
            full[:, kh:kh + 2 * H:2, kw:kw + 2 * W:2, :] += x @ K[kh, kw].T
    return full[:, 1:1 + 2 * H, 1:1 + 2 * W, :] + b.astype(np.float64)


def bn_frozen(x, gamma, beta, eps=1e-3):
    return x.astype(np.float64) * (gamma.astype(np.float64) / math.sqrt(1.0 + eps)) + beta.astype(np.float64)


def lrelu(x, alpha=0.3):
    return np.where(x > 0, x, alpha * x)


def kl_per_sample(mu, log_sigma):
    """trainers/VAE.py:38 literally: 0.5*sum(mu^2 + sigma^2 - log(sigma^2) - 1)."""
    mu = mu.astype(np.float64)
    s = np.exp(log_sigma.astype(np.float64))
    return 0.5 * np.sum(mu ** 2 + s ** 2 - np.log(s ** 2) - 1.0, axis=1)


def adam_tf(p, g, m, v, t, lr, b1=0.5, b2=0.999, eps=1e-8):
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    return p - lr_t * m / (np.sqrt(v) + eps), m, v


def layernorm_hw(x, gamma_hw, beta_hw, eps=1e-3):
    """tf.keras.layers.LayerNormalization(axis=[1,2]) on NHWC: stats over (H,W) per (b,c); gamma/beta [H,W]."""
    x = x.astype(np.float64)
    mu = x.mean(axis=(1, 2), keepdims=True)
    var = x.var(axis=(1, 2), keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * gamma_hw[None, :, :, None] + beta_hw[None, :, :, None]
