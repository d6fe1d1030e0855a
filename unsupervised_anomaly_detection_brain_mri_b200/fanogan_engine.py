"""Device executor of the f-AnoGAN forward paths (reference models/fanogan.py:11-84): Encoder -> z_enc (tanh),
Generator -> x_enc = sigmoid(G(z)) with LayerNormalization([1,2]) blocks, Discriminator feature stack + Dense(1).

Scope this round: the forward / reconstruct / scoring path (trainers/fAnoGAN.py:220-239 + Evaluation).  The WGAN-GP
training step (fAnoGAN.py:50-77: gradient penalty = double backward through conv / LayerNorm) is the next §8 row."""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from . import abi
from .abi import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, OP_CONV_FWD, OP_CONVT_FWD, call, ptr
from .engine import BN_C, KSIZE, LRELU_ALPHA, FlatParams, glorot_init, stack_plan

LN_EPS = 1e-3


def _ln(k):
    return 'layer_normalization' if k == 0 else f'layer_normalization_{k}'


def _bn(k):
    return 'batch_normalization' if k == 0 else f'batch_normalization_{k}'


def param_specs(S, C=1, zDim=128, res=8):
    """TF variable names -> shapes for the three scopes Encoder / Generator / Discriminator (selected by substring in
    trainers/fAnoGAN.py:71-73, so the flat buffer is laid out scope-contiguous)."""
    n, enc, dec = stack_plan(S, res)
    sp = OrderedDict()
    cin = C
    for i, co in enumerate(enc):
        sp[f'Encoder/enc_conv2D_{i}/kernel'] = (KSIZE, KSIZE, cin, co)
        sp[f'Encoder/enc_conv2D_{i}/bias'] = (co,)
        sp[f'Encoder/{_bn(i)}/gamma'] = (co,)
        sp[f'Encoder/{_bn(i)}/beta'] = (co,)
        cin = co
    cb = cin // 8
    flat = res * res * cb
    sp['Encoder/conv2d/kernel'] = (1, 1, cin, cb)
    sp['Encoder/conv2d/bias'] = (cb,)
    sp['Encoder/dense/kernel'] = (flat, zDim)
    sp['Encoder/dense/bias'] = (zDim,)
    sp['Generator/conv2d_1/kernel'] = (1, 1, cb, cin)
    sp['Generator/conv2d_1/bias'] = (cin,)
    sp['Generator/dense_1/kernel'] = (zDim, flat)
    sp['Generator/dense_1/bias'] = (flat,)
    ln, s = 0, res
    sp[f'Generator/{_ln(ln)}/gamma'] = (s, s)
    sp[f'Generator/{_ln(ln)}/beta'] = (s, s)
    ln += 1
    for i, co in enumerate(dec):
        sp[f'Generator/dec_Conv2DT_{i}/kernel'] = (KSIZE, KSIZE, co, cin)
        sp[f'Generator/dec_Conv2DT_{i}/bias'] = (co,)
        s *= 2
        sp[f'Generator/{_ln(ln)}/gamma'] = (s, s)
        sp[f'Generator/{_ln(ln)}/beta'] = (s, s)
        ln += 1
        cin = co
    sp['Generator/dec_Conv2D_final/kernel'] = (1, 1, cin, C)
    sp['Generator/dec_Conv2D_final/bias'] = (C,)
    cin, s = C, S
    for i, co in enumerate(enc):
        sp[f'Discriminator/enc_conv2D_{i}/kernel'] = (KSIZE, KSIZE, cin, co)
        sp[f'Discriminator/enc_conv2D_{i}/bias'] = (co,)
        s //= 2
        sp[f'Discriminator/{_ln(ln)}/gamma'] = (s, s)
        sp[f'Discriminator/{_ln(ln)}/beta'] = (s, s)
        ln += 1
        cin = co
    sp['Discriminator/dense_2/kernel'] = (cin, 1)
    sp['Discriminator/dense_2/bias'] = (1,)
    return sp


class FanoganEngine:
    def __init__(self, S, C=1, zDim=128, res=8, batch=8, device='cuda:0', math_mode=abi.MATH_TC_3XTF32, seed=1):
        if C != 1:
            raise NotImplementedError('numChannels == 1 only (all reference datasets are single-channel)')
        abi.lib()
        self.S, self.C, self.zDim, self.res, self.B = S, C, zDim, res, batch
        self.device = torch.device(device)
        self.math_mode = math_mode
        self.n, self.enc_ch, self.dec_ch = stack_plan(S, res)
        self.cb = self.enc_ch[-1] // 8
        self.flat = res * res * self.cb
        self.specs = param_specs(S, C, zDim, res)
        self.fp = FlatParams(self.specs, self.device)
        init = glorot_init(self.specs, seed)
        self.fp.load(init)
        self._alloc()

    def _new(self, *shape):
        return torch.empty(*shape, dtype=torch.float32, device=self.device)

    def _alloc(self):
        B, S = self.B, self.S
        self.x = self._new(B, S, S, 1)
        self.enc_a, self.dis_z, self.dis_a = [], [], []
        s = S
        for co in self.enc_ch:
            s //= 2
            self.enc_a.append(self._new(B, s, s, co))
            self.dis_z.append(self._new(B, s, s, co))
            self.dis_a.append(self._new(B, s, s, co))
        r = self.res
        self.zb = self._new(B, r, r, self.cb)
        self.z_pre = self._new(B, self.zDim)
        self.z_enc = self._new(B, self.zDim)
        self.d = self._new(B, self.flat)
        self.zr = self._new(B, r, r, self.enc_ch[-1])
        self.ar = self._new(B, r, r, self.enc_ch[-1])
        self.gen_z, self.gen_a = [], []
        s = r
        for co in self.dec_ch:
            s *= 2
            self.gen_z.append(self._new(B, s, s, co))
            self.gen_a.append(self._new(B, s, s, co))
        self.g_pre = self._new(B, S, S, 1)
        self.x_enc = self._new(B, S, S, 1)
        self.d_out = self._new(B, r, r, 1)
        L = abi.lib()
        need = 1 << 20
        s, cin = S, 1
        for co in self.enc_ch:
            need = max(need, L.uad_conv_workspace_bytes(OP_CONV_FWD, B, s, s, cin, co, KSIZE, self.math_mode))
            need = max(need, L.uad_layernorm_hw_workspace_bytes(B, (s // 2) ** 2, co))
            s //= 2
            cin = co
        need = max(need, L.uad_layernorm_hw_workspace_bytes(B, s * s, cin))
        for co in self.dec_ch:
            need = max(need, L.uad_conv_workspace_bytes(OP_CONVT_FWD, B, s, s, cin, co, KSIZE, self.math_mode))
            need = max(need, L.uad_layernorm_hw_workspace_bytes(B, (2 * s) ** 2, co))
            s *= 2
            cin = co
        r2 = r * r
        for (M, K, N) in ((B * r2, self.enc_ch[-1], self.cb), (B * r2, self.cb, self.enc_ch[-1]), (B, self.flat, self.zDim),
                          (B, self.zDim, self.flat), (B * r2, self.enc_ch[-1], 1)):
            need = max(need, L.uad_dense_workspace_bytes(M, K, N))
        self.ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        self.ws_bytes = need

    def _st(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def set_inputs(self, x):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, np.float32))
        self.x.copy_(x.reshape(self.x.shape), non_blocking=True)

    def encode(self, mask=None, keep=1.0):
        """x -> z_enc = tanh(dropout(Dense(flatten(conv1x1(encoder(x))))))   (fanogan.py:15-29)"""
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self.ws.data_ptr(), self.ws_bytes
        B = self.B
        h, s, cin = self.x, self.S, 1
        for i, co in enumerate(self.enc_ch):
            pre, bnn = f'Encoder/enc_conv2D_{i}', f'Encoder/{_bn(i)}'
            call('uad_conv2d_fwd', ptr(h), ptr(fp.p(pre + '/kernel')), ptr(fp.p(pre + '/bias')), ptr(fp.p(bnn + '/gamma')),
                 ptr(fp.p(bnn + '/beta')), None, ptr(self.enc_a[i]), B, s, s, cin, co, KSIZE, ACT_LEAKY, LRELU_ALPHA, BN_C, mm, ws,
                 wsb, st)
            h, s, cin = self.enc_a[i], s // 2, co
        r2 = self.res * self.res
        call('uad_dense_fwd', ptr(h), ptr(fp.p('Encoder/conv2d/kernel')), ptr(fp.p('Encoder/conv2d/bias')), None, 1.0, None, None,
             ptr(self.zb), None, B * r2, cin, self.cb, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        call('uad_dense_fwd', ptr(self.zb), ptr(fp.p('Encoder/dense/kernel')), ptr(fp.p('Encoder/dense/bias')), ptr(mask), keep,
             None, None, ptr(self.z_pre), ptr(self.z_enc), B, self.flat, self.zDim, ACT_TANH, 0.0, 1.0, ws, wsb, st)
        return self.z_enc

    def generate(self, z, mask=None, keep=1.0):
        """z -> sigmoid(G(z))   (fanogan.py:33-46)"""
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self.ws.data_ptr(), self.ws_bytes
        B, r = self.B, self.res
        r2 = r * r
        ctop = self.enc_ch[-1]
        call('uad_dense_fwd', ptr(z), ptr(fp.p('Generator/dense_1/kernel')), ptr(fp.p('Generator/dense_1/bias')), ptr(mask), keep,
             None, None, ptr(self.d), None, B, self.zDim, self.flat, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        call('uad_dense_fwd', ptr(self.d), ptr(fp.p('Generator/conv2d_1/kernel')), ptr(fp.p('Generator/conv2d_1/bias')), None, 1.0,
             None, None, ptr(self.zr), None, B * r2, self.cb, ctop, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        ln = 0
        call('uad_layernorm_hw_fwd', ptr(self.zr), ptr(fp.p(f'Generator/{_ln(ln)}/gamma')), ptr(fp.p(f'Generator/{_ln(ln)}/beta')),
             ptr(self.ar), B, r2, ctop, LN_EPS, ACT_RELU, 0.0, ws, wsb, st)
        ln += 1
        h, s, cin = self.ar, r, ctop
        for i, co in enumerate(self.dec_ch):
            pre = f'Generator/dec_Conv2DT_{i}'
            call('uad_convT2d_fwd', ptr(h), ptr(fp.p(pre + '/kernel')), ptr(fp.p(pre + '/bias')), None, None, ptr(self.gen_z[i]), None,
                 B, s, s, cin, co, KSIZE, ACT_NONE, 0.0, 1.0, mm, ws, wsb, st)
            s *= 2
            call('uad_layernorm_hw_fwd', ptr(self.gen_z[i]), ptr(fp.p(f'Generator/{_ln(ln)}/gamma')),
                 ptr(fp.p(f'Generator/{_ln(ln)}/beta')), ptr(self.gen_a[i]), B, s * s, co, LN_EPS, ACT_LEAKY, LRELU_ALPHA, ws, wsb, st)
            ln += 1
            h, cin = self.gen_a[i], co
        # final 1x1 conv (Cin -> 1), then sigmoid (fanogan.py:41)
        call('uad_final1x1_l1_fwd', ptr(h), ptr(fp.p('Generator/dec_Conv2D_final/kernel')), ptr(fp.p('Generator/dec_Conv2D_final/bias')),
             ptr(self.x), ptr(self.g_pre), None, None, B, self.S * self.S, cin, ws, wsb, st)
        call('uad_activation', ptr(self.g_pre), ptr(self.x_enc), self.g_pre.numel(), ACT_SIGMOID, 0.0, st)
        return self.x_enc

    def discriminate(self, x_dev):
        """x -> (features [B,r,r,128], critic [B,r,r,1])   (fanogan.py:50-58; Dense(1) acts on the channel axis)"""
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self.ws.data_ptr(), self.ws_bytes
        B = self.B
        ln = self.n + 1
        h, s, cin = x_dev, self.S, 1
        for i, co in enumerate(self.enc_ch):
            pre = f'Discriminator/enc_conv2D_{i}'
            call('uad_conv2d_fwd', ptr(h), ptr(fp.p(pre + '/kernel')), ptr(fp.p(pre + '/bias')), None, None, ptr(self.dis_z[i]), None, B,
                 s, s, cin, co, KSIZE, ACT_NONE, 0.0, 1.0, mm, ws, wsb, st)
            s //= 2
            call('uad_layernorm_hw_fwd', ptr(self.dis_z[i]), ptr(fp.p(f'Discriminator/{_ln(ln)}/gamma')),
                 ptr(fp.p(f'Discriminator/{_ln(ln)}/beta')), ptr(self.dis_a[i]), B, s * s, co, LN_EPS, ACT_LEAKY, LRELU_ALPHA, ws, wsb,
                 st)
            ln += 1
            h, cin = self.dis_a[i], co
        r2 = self.res * self.res
        call('uad_dense_fwd', ptr(h), ptr(fp.p('Discriminator/dense_2/kernel')), ptr(fp.p('Discriminator/dense_2/bias')), None, 1.0,
             None, None, ptr(self.d_out), None, B * r2, cin, 1, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        return h, self.d_out

    def reconstruct(self):
        """x_enc = sigmoid(G(E(x))) with dropout off (trainers/fAnoGAN.py:220-239)."""
        return self.generate(self.encode())
