"""Device executor of the f-AnoGAN forward paths (reference models/fanogan.py:11-84): Encoder -> z_enc (tanh),
Generator -> x_enc = sigmoid(G(z)) with LayerNormalization([1,2]) blocks, Discriminator feature stack + Dense(1).

Inference (reconstruct / scoring, trainers/fAnoGAN.py:220-239) and the three train ops of fAnoGAN.py:50-77:
``step_gen`` (optim_gen), ``step_disc`` (optim_dis, WGAN-GP) and ``step_enc`` (optim_enc, izi_f).  The gradient penalty's
second-order term is evaluated without a tape: with u = dGP/d(ddx) fixed, grad_theta GP equals the ordinary reverse pass of
the directional derivative of sum(D(x_hat)) along u, so the critic runs forward -> reverse to x_hat -> tangent forward
(uad_layernorm_hw_jvp) -> joint reverse (uad_layernorm_hw_bwd2); every conv in those passes is the same fwd / dgrad / wgrad
kernel the autoencoders use."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from . import abi
from .abi import (ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, OP_CONV_DGRAD, OP_CONV_FWD, OP_CONV_WGRAD, OP_CONVT_DGRAD,
                  OP_CONVT_FWD, OP_CONVT_WGRAD, call, ptr)
from .engine import BN_C, KSIZE, LRELU_ALPHA, FlatParams, glorot_init, graph_capture, stack_plan

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


class _CriticPass:
    """Activations of one pass through the Discriminator stack (fanogan.py:54-82 runs it on x_, x, x_hat and x_enc)."""

    def __init__(self, eng):
        B, s = eng.B, eng.S
        self.z, self.a, self.stats = [], [], []
        for co in eng.enc_ch:
            s //= 2
            self.z.append(eng._new(B, s, s, co))
            self.a.append(eng._new(B, s, s, co))
            self.stats.append(eng._new(2 * B * co))
        self.d = eng._new(B, eng.res, eng.res, 1)


class FanoganEngine:
    SC = dict(disc_fake=0, disc_real=1, gp=2, loss_img=3, loss_fts=4, reconstructionLoss=5)
    # what a sibling graph on the same stacks overrides (anovaegan_engine.AnoVaeGanEngine): TF names of the Dense layers (the
    # tf.layers name counter runs over the whole graph) and the generator's output non-linearity
    GEN_DENSE = 'Generator/dense_1'
    DISC_DENSE = 'Discriminator/dense_2'
    FINAL_ACT = ACT_SIGMOID

    @staticmethod
    def _param_specs(S, C, zDim, res):
        return param_specs(S, C, zDim, res)

    def __init__(self, S, C=1, zDim=128, res=8, batch=8, device='cuda:0', math_mode=abi.MATH_TC_3XTF32, seed=1, kappa=1.0,
                 scale=10.0):
        if C != 1:
            raise NotImplementedError('numChannels == 1 only (all reference datasets are single-channel)')
        abi.lib()
        self.S, self.C, self.zDim, self.res, self.B = S, C, zDim, res, batch
        self.kappa, self.scale = float(kappa), float(scale)
        self.device = torch.device(device)
        self.math_mode = math_mode
        self.n, self.enc_ch, self.dec_ch = stack_plan(S, res)
        self.cb = self.enc_ch[-1] // 8
        self.flat = res * res * self.cb
        self.specs = self._param_specs(S, C, zDim, res)
        self.fp = FlatParams(self.specs, self.device)
        init = glorot_init(self.specs, seed)
        self.fp.load(init)
        self.seed = int(seed)
        self._train_ready = False
        self._alloc()

    def _new(self, *shape):
        return torch.empty(*shape, dtype=torch.float32, device=self.device)

    def _alloc(self):
        B, S, r = self.B, self.S, self.res
        ctop = self.enc_ch[-1]
        self.x = self._new(B, S, S, 1)
        self.enc_z, self.enc_a = [], []
        s = S
        for co in self.enc_ch:
            s //= 2
            self.enc_z.append(None)                         # allocated by enable_training()
            self.enc_a.append(self._new(B, s, s, co))
        self.zb = self._new(B, r, r, self.cb)
        self.z_pre = self._new(B, self.zDim)
        self.z_enc = self._new(B, self.zDim)
        self.d = self._new(B, self.flat)
        self.zr = self._new(B, r, r, ctop)
        self.ar = self._new(B, r, r, ctop)
        self.g_stats_top = self._new(2 * B * ctop)
        self.gen_z, self.gen_a, self.gen_stats = [], [], []
        s = r
        for co in self.dec_ch:
            s *= 2
            self.gen_z.append(self._new(B, s, s, co))
            self.gen_a.append(self._new(B, s, s, co))
            self.gen_stats.append(self._new(2 * B * co))
        self.g_pre = self._new(B, S, S, 1)
        self.x_enc = self._new(B, S, S, 1)                  # sigmoid(G(.)) of the last generate() call
        self.pass0 = _CriticPass(self)
        self.dis_z, self.dis_a, self.d_out = self.pass0.z, self.pass0.a, self.pass0.d
        L = abi.lib()
        mm = self.math_mode
        need = 1 << 20
        s, cin = S, 1
        for co in self.enc_ch:
            for op in (OP_CONV_FWD, OP_CONV_DGRAD, OP_CONV_WGRAD):
                need = max(need, L.uad_conv_workspace_bytes(op, B, s, s, cin, co, KSIZE, mm))
            need = max(need, L.uad_layernorm_hw_train_workspace_bytes(B, (s // 2) ** 2, co))
            need = max(need, L.uad_rowreduce_workspace_bytes(B * (s // 2) ** 2, co))
            s //= 2
            cin = co
        need = max(need, L.uad_layernorm_hw_train_workspace_bytes(B, s * s, cin))
        for co in self.dec_ch:
            for op in (OP_CONVT_FWD, OP_CONVT_DGRAD, OP_CONVT_WGRAD):
                need = max(need, L.uad_conv_workspace_bytes(op, B, s, s, cin, co, KSIZE, mm))
            need = max(need, L.uad_layernorm_hw_train_workspace_bytes(B, (2 * s) ** 2, co))
            s *= 2
            cin = co
        r2 = r * r
        for (M, K, N) in ((B * r2, ctop, self.cb), (B * r2, self.cb, ctop), (B, self.flat, self.zDim),
                          (B, self.zDim, self.flat), (B * r2, ctop, 1)):
            need = max(need, L.uad_dense_workspace_bytes(M, K, N))
        need = max(need, 4 * 148 * 129 * 4, L.uad_reduce_workspace_bytes(), B * S * 4)
        self.ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        self.ws_bytes = need

    # ------------------------------------------------------------------ training state
    def enable_training(self):
        if self._train_ready:
            return
        B, S, r = self.B, self.S, self.res
        ctop = self.enc_ch[-1]
        for i, a in enumerate(self.enc_a):
            self.enc_z[i] = torch.empty_like(a)
        self.enc_g = [torch.empty_like(a) for a in self.enc_a]           # gradient scratch per encoder level
        self.pass1 = _CriticPass(self)
        self.dis_zd = [torch.empty_like(t) for t in self.pass0.z]        # tangent pre-norm / activations of the GP pass
        self.dis_hd = [torch.empty_like(t) for t in self.pass0.z]
        self.dis_js = [torch.empty_like(t) for t in self.pass0.stats]
        self.dis_g1 = [torch.empty_like(t) for t in self.pass0.z]        # gradient scratch per critic level (two chains)
        self.dis_g2 = [torch.empty_like(t) for t in self.pass0.z]
        self.gen_g = [torch.empty_like(t) for t in self.gen_z]
        self.dzr = self._new(B, r, r, ctop)
        self.dd = self._new(B, self.flat)
        self.dzb = self._new(B, r, r, self.cb)
        self.dz_lat = self._new(B, self.zDim)
        self.x_gen = self._new(B, S, S, 1)
        self.x_hat = self._new(B, S, S, 1)
        self.ddx = self._new(B, S, S, 1)
        self.u = self._new(B, S, S, 1)
        self.dxi = self._new(B, S, S, 1)
        self.dxi2 = self._new(B, S, S, 1)
        self.ones = self._new(B, r, r, 1)
        self.z_in = self._new(B, self.zDim)
        self.alpha = self._new(B)
        self.mask_enc = self._new(B, self.zDim)
        self.mask_gen = self._new(B, self.flat)
        self.l1 = self._new(B, S, S, 1)
        self.rec = self._new(B)
        self.sc = torch.zeros(16, dtype=torch.float32, device=self.device)
        self.steps = {k: torch.zeros(1, dtype=torch.int64, device=self.device) for k in ('Encoder', 'Generator', 'Discriminator')}
        self.t = {k: 0 for k in self.steps}
        self.rng_ctr = torch.zeros(1, dtype=torch.int64, device=self.device)   # Philox offset lives on the device (graph-capturable)
        self._graphs, self._warm, self._graph_scopes = {}, {}, {}
        self._train_ready = True

    def _st(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def set_inputs(self, x):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, np.float32))
        self.x.copy_(x.reshape(self.x.shape), non_blocking=True)

    def _wsp(self):
        return self.ws.data_ptr(), self.ws_bytes

    # ------------------------------------------------------------------ forward passes
    def encode(self, mask=None, keep=1.0):
        """x -> z_enc = tanh(dropout(Dense(flatten(conv1x1(encoder(x))))))   (fanogan.py:15-29)"""
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self._wsp()
        B = self.B
        h, s, cin = self.x, self.S, 1
        for i, co in enumerate(self.enc_ch):
            pre, bnn = f'Encoder/enc_conv2D_{i}', f'Encoder/{_bn(i)}'
            call('uad_conv2d_fwd', ptr(h), ptr(fp.p(pre + '/kernel')), ptr(fp.p(pre + '/bias')), ptr(fp.p(bnn + '/gamma')),
                 ptr(fp.p(bnn + '/beta')), ptr(self.enc_z[i]), ptr(self.enc_a[i]), B, s, s, cin, co, KSIZE, ACT_LEAKY, LRELU_ALPHA,
                 BN_C, mm, ws, wsb, st)
            h, s, cin = self.enc_a[i], s // 2, co
        r2 = self.res * self.res
        call('uad_dense_fwd', ptr(h), ptr(fp.p('Encoder/conv2d/kernel')), ptr(fp.p('Encoder/conv2d/bias')), None, 1.0, None, None,
             ptr(self.zb), None, B * r2, cin, self.cb, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        call('uad_dense_fwd', ptr(self.zb), ptr(fp.p('Encoder/dense/kernel')), ptr(fp.p('Encoder/dense/bias')), ptr(mask), keep,
             None, None, ptr(self.z_pre), ptr(self.z_enc), B, self.flat, self.zDim, ACT_TANH, 0.0, 1.0, ws, wsb, st)
        return self.z_enc

    def generate(self, z, mask=None, keep=1.0, out=None):
        """z -> sigmoid(G(z))   (fanogan.py:33-46); LayerNorm statistics are kept for the backward passes."""
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self._wsp()
        B, r = self.B, self.res
        r2 = r * r
        ctop = self.enc_ch[-1]
        out = self.x_enc if out is None else out
        self._g_in, self._g_mask, self._g_keep = z, mask, keep
        call('uad_dense_fwd', ptr(z), ptr(fp.p(self.GEN_DENSE + '/kernel')), ptr(fp.p(self.GEN_DENSE + '/bias')), ptr(mask), keep,
             None, None, ptr(self.d), None, B, self.zDim, self.flat, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        call('uad_dense_fwd', ptr(self.d), ptr(fp.p('Generator/conv2d_1/kernel')), ptr(fp.p('Generator/conv2d_1/bias')), None, 1.0,
             None, None, ptr(self.zr), None, B * r2, self.cb, ctop, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        ln = 0
        call('uad_layernorm_hw_fwd_train', ptr(self.zr), ptr(fp.p(f'Generator/{_ln(ln)}/gamma')),
             ptr(fp.p(f'Generator/{_ln(ln)}/beta')), ptr(self.ar), ptr(self.g_stats_top), B, r2, ctop, LN_EPS, ACT_RELU, 0.0, ws, wsb,
             st)
        ln += 1
        h, s, cin = self.ar, r, ctop
        for i, co in enumerate(self.dec_ch):
            pre = f'Generator/dec_Conv2DT_{i}'
            call('uad_convT2d_fwd', ptr(h), ptr(fp.p(pre + '/kernel')), ptr(fp.p(pre + '/bias')), None, None, ptr(self.gen_z[i]), None,
                 B, s, s, cin, co, KSIZE, ACT_NONE, 0.0, 1.0, mm, ws, wsb, st)
            s *= 2
            call('uad_layernorm_hw_fwd_train', ptr(self.gen_z[i]), ptr(fp.p(f'Generator/{_ln(ln)}/gamma')),
                 ptr(fp.p(f'Generator/{_ln(ln)}/beta')), ptr(self.gen_a[i]), ptr(self.gen_stats[i]), B, s * s, co, LN_EPS, ACT_LEAKY,
                 LRELU_ALPHA, ws, wsb, st)
            ln += 1
            h, cin = self.gen_a[i], co
        # final 1x1 conv (Cin -> 1), then sigmoid (fanogan.py:41)
        head = self.g_pre if self.FINAL_ACT != ACT_NONE else out
        call('uad_final1x1_l1_fwd', ptr(h), ptr(fp.p('Generator/dec_Conv2D_final/kernel')), ptr(fp.p('Generator/dec_Conv2D_final/bias')),
             ptr(self.x), ptr(head), None, None, B, self.S * self.S, cin, ws, wsb, st)
        if self.FINAL_ACT != ACT_NONE:
            call('uad_activation', ptr(self.g_pre), ptr(out), self.g_pre.numel(), self.FINAL_ACT, 0.0, st)
        return out

    def _critic_forward(self, cp, x_dev, critic=True):
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self._wsp()
        B = self.B
        ln = self.n + 1
        h, s, cin = x_dev, self.S, 1
        for i, co in enumerate(self.enc_ch):
            pre = f'Discriminator/enc_conv2D_{i}'
            call('uad_conv2d_fwd', ptr(h), ptr(fp.p(pre + '/kernel')), ptr(fp.p(pre + '/bias')), None, None, ptr(cp.z[i]), None, B,
                 s, s, cin, co, KSIZE, ACT_NONE, 0.0, 1.0, mm, ws, wsb, st)
            s //= 2
            call('uad_layernorm_hw_fwd_train', ptr(cp.z[i]), ptr(fp.p(f'Discriminator/{_ln(ln)}/gamma')),
                 ptr(fp.p(f'Discriminator/{_ln(ln)}/beta')), ptr(cp.a[i]), ptr(cp.stats[i]), B, s * s, co, LN_EPS, ACT_LEAKY,
                 LRELU_ALPHA, ws, wsb, st)
            ln += 1
            h, cin = cp.a[i], co
        if critic:
            r2 = self.res * self.res
            call('uad_dense_fwd', ptr(h), ptr(fp.p(self.DISC_DENSE + '/kernel')), ptr(fp.p(self.DISC_DENSE + '/bias')), None,
                 1.0, None, None, ptr(cp.d), None, B * r2, cin, 1, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        return h, cp.d

    def discriminate(self, x_dev):
        """x -> (features [B,r,r,128], critic [B,r,r,1])   (fanogan.py:50-58; Dense(1) acts on the channel axis)"""
        return self._critic_forward(self.pass0, x_dev)

    def reconstruct(self):
        """x_enc = sigmoid(G(E(x))) with dropout off (trainers/fAnoGAN.py:220-239)."""
        return self.generate(self.encode())

    # ------------------------------------------------------------------ backward passes
    def _critic_top(self, cp, coef, params):
        """Gradient of coef*sum(critic) w.r.t. the feature map (-> dis_g1[-1]); Dense(1) parameter gradients if asked."""
        fp, st = self.fp, self._st()
        ws, wsb = self._wsp()
        r2 = self.res * self.res
        ctop = self.enc_ch[-1]
        call('uad_fill', ptr(self.ones), float(coef), self.ones.numel(), st)
        call('uad_dense_bwd', ptr(cp.a[-1]), ptr(fp.p(self.DISC_DENSE + '/kernel')), ptr(self.ones), None, 1.0,
             ptr(self.dis_g1[-1]), ptr(fp.g(self.DISC_DENSE + '/kernel')) if params else None,
             ptr(fp.g(self.DISC_DENSE + '/bias')) if params else None, self.B * r2, ctop, 1, 1, ws, wsb, st)
        return self.dis_g1[-1]

    def _critic_backward(self, cp, x_in, params, dx_out):
        """Reverse pass of one critic pass from the gradient held in dis_g1[-1] (w.r.t. the features).  params: accumulate
        kernel / LayerNorm gradients.  Conv biases feed a LayerNorm over (H,W), which removes any per-channel constant, so
        their exact gradient is zero and is left at zero.  dx_out: buffer for the gradient w.r.t. the input image."""
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self._wsp()
        B = self.B
        n = len(self.enc_ch)
        sizes = [self.S >> k for k in range(n + 1)]
        for i in reversed(range(n)):
            co = self.enc_ch[i]
            cin = 1 if i == 0 else self.enc_ch[i - 1]
            s_in, s_out = sizes[i], sizes[i + 1]
            ln = self.n + 1 + i
            g = self.dis_g1[i]
            gam, bet = f'Discriminator/{_ln(ln)}/gamma', f'Discriminator/{_ln(ln)}/beta'
            call('uad_layernorm_hw_bwd', ptr(g), ptr(cp.z[i]), ptr(cp.stats[i]), ptr(fp.p(gam)), ptr(fp.p(bet)), ptr(g),
                 ptr(fp.g(gam)) if params else None, ptr(fp.g(bet)) if params else None, B, s_out * s_out, co, ACT_LEAKY,
                 LRELU_ALPHA, 1, ws, wsb, st)
            pre = f'Discriminator/enc_conv2D_{i}'
            src = x_in if i == 0 else cp.a[i - 1]
            if params:
                call('uad_conv2d_wgrad', ptr(src), ptr(g), ptr(fp.g(pre + '/kernel')), B, s_in, s_in, cin, co, KSIZE, 1, mm, ws, wsb,
                     st)
            dst = dx_out if i == 0 else self.dis_g1[i - 1]
            if dst is not None:
                call('uad_conv2d_dgrad', ptr(g), ptr(fp.p(pre + '/kernel')), ptr(dst), B, s_in, s_in, cin, co, KSIZE, mm, ws, wsb, st)

    def _critic_gp(self, cp, x_hat):
        """Gradient penalty on an already-forwarded pass: ddx, gp scalar (sc[2]) and, accumulated into the critic's gradient
        slots, d gp / d theta  (trainers/fAnoGAN.py:55-58)."""
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self._wsp()
        B, S = self.B, self.S
        n = len(self.enc_ch)
        sizes = [S >> k for k in range(n + 1)]
        r2 = self.res * self.res
        ctop = self.enc_ch[-1]
        wd = fp.p(self.DISC_DENSE + '/kernel')
        # (1) ddx = d sum(d_hat) / d x_hat
        self._critic_top(cp, 1.0, params=False)
        self._critic_backward(cp, x_hat, params=False, dx_out=self.ddx)
        # (2) gp and its seed u = d gp / d ddx
        call('uad_gradient_penalty', ptr(self.ddx), B, S, S * self.C, self.scale, ptr(self.u), self.sc[2:].data_ptr(), ws, wsb, st)
        # (3) tangent forward along u
        hd, s, cin = self.u, S, 1
        for i, co in enumerate(self.enc_ch):
            pre = f'Discriminator/enc_conv2D_{i}'
            ln = self.n + 1 + i
            call('uad_conv2d_fwd', ptr(hd), ptr(fp.p(pre + '/kernel')), None, None, None, ptr(self.dis_zd[i]), None, B, s, s, cin, co,
                 KSIZE, ACT_NONE, 0.0, 1.0, mm, ws, wsb, st)
            s //= 2
            call('uad_layernorm_hw_jvp', ptr(self.dis_zd[i]), ptr(cp.z[i]), ptr(cp.stats[i]), ptr(fp.p(f'Discriminator/{_ln(ln)}/gamma')),
                 ptr(fp.p(f'Discriminator/{_ln(ln)}/beta')), ptr(self.dis_hd[i]), ptr(self.dis_js[i]), B, s * s, co, ACT_LEAKY,
                 LRELU_ALPHA, ws, wsb, st)
            hd, cin = self.dis_hd[i], co
        # (4) joint reverse of s_dot = sum(hd_top . w_d): adjoint of hd_top = w_d, of h_top = 0; d/dw_d = sum hd_top
        call('uad_fill', ptr(self.ones), 1.0, self.ones.numel(), st)
        call('uad_dense_bwd', ptr(self.dis_hd[-1]), ptr(wd), ptr(self.ones), None, 1.0, ptr(self.dis_g1[-1]),
             ptr(fp.g(self.DISC_DENSE + '/kernel')), None, B * r2, ctop, 1, 1, ws, wsb, st)
        have_dh = False
        for i in reversed(range(n)):
            co = self.enc_ch[i]
            cin = 1 if i == 0 else self.enc_ch[i - 1]
            s_in, s_out = sizes[i], sizes[i + 1]
            ln = self.n + 1 + i
            gam, bet = f'Discriminator/{_ln(ln)}/gamma', f'Discriminator/{_ln(ln)}/beta'
            g1, g2 = self.dis_g1[i], self.dis_g2[i]
            call('uad_layernorm_hw_bwd2', ptr(g1), ptr(g2) if have_dh else None, ptr(cp.z[i]), ptr(self.dis_zd[i]), ptr(cp.stats[i]),
                 ptr(self.dis_js[i]), ptr(fp.p(gam)), ptr(fp.p(bet)), ptr(g1), ptr(g2), ptr(fp.g(gam)), ptr(fp.g(bet)), B,
                 s_out * s_out, co, ACT_LEAKY, LRELU_ALPHA, 1, ws, wsb, st)
            pre = f'Discriminator/enc_conv2D_{i}'
            src_t = self.u if i == 0 else self.dis_hd[i - 1]
            src_p = x_hat if i == 0 else cp.a[i - 1]
            call('uad_conv2d_wgrad', ptr(src_t), ptr(g1), ptr(fp.g(pre + '/kernel')), B, s_in, s_in, cin, co, KSIZE, 1, mm, ws, wsb, st)
            call('uad_conv2d_wgrad', ptr(src_p), ptr(g2), ptr(fp.g(pre + '/kernel')), B, s_in, s_in, cin, co, KSIZE, 1, mm, ws, wsb, st)
            if i > 0:
                call('uad_conv2d_dgrad', ptr(g1), ptr(fp.p(pre + '/kernel')), ptr(self.dis_g1[i - 1]), B, s_in, s_in, cin, co, KSIZE,
                     mm, ws, wsb, st)
                call('uad_conv2d_dgrad', ptr(g2), ptr(fp.p(pre + '/kernel')), ptr(self.dis_g2[i - 1]), B, s_in, s_in, cin, co, KSIZE,
                     mm, ws, wsb, st)
                have_dh = True

    def _generator_backward(self, dx_out, params, dz_out=None, head=True):
        """Reverse pass of the last generate() call from d/d(output).  params: Generator gradients (overwritten);
        dz_out: gradient w.r.t. the latent input.  head=False: the caller has already taken the final 1x1 conv's backward
        (gradient w.r.t. the last block's activation in gen_g[-1], its parameter gradients written)."""
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self._wsp()
        B, r = self.B, self.res
        r2 = r * r
        ctop = self.enc_ch[-1]
        nd = len(self.dec_ch)
        G = (lambda name: ptr(fp.g(name))) if params else (lambda name: None)
        if head:
            seed = dx_out
            if self.FINAL_ACT != ACT_NONE:
                call('uad_activation_bwd', ptr(dx_out), ptr(self.g_pre), ptr(self.dxi2), self.g_pre.numel(), self.FINAL_ACT, 0.0, st)
                seed = self.dxi2
            call('uad_final1x1_bwd', ptr(self.gen_a[-1]), ptr(fp.p('Generator/dec_Conv2D_final/kernel')), ptr(seed), ptr(self.gen_g[-1]),
                 G('Generator/dec_Conv2D_final/kernel'), G('Generator/dec_Conv2D_final/bias'), B, self.S * self.S, self.dec_ch[-1], 0,
                 ws, wsb, st)
        s = self.S
        for i in reversed(range(nd)):
            co = self.dec_ch[i]
            ci = ctop if i == 0 else self.dec_ch[i - 1]
            ln = i + 1
            gam, bet = f'Generator/{_ln(ln)}/gamma', f'Generator/{_ln(ln)}/beta'
            g = self.gen_g[i]
            call('uad_layernorm_hw_bwd', ptr(g), ptr(self.gen_z[i]), ptr(self.gen_stats[i]), ptr(fp.p(gam)), ptr(fp.p(bet)), ptr(g),
                 G(gam), G(bet), B, s * s, co, ACT_LEAKY, LRELU_ALPHA, 0, ws, wsb, st)
            pre = f'Generator/dec_Conv2DT_{i}'
            src = self.ar if i == 0 else self.gen_a[i - 1]
            dst = self.dzr if i == 0 else self.gen_g[i - 1]
            s //= 2
            if params:
                call('uad_convT2d_wgrad', ptr(src), ptr(g), ptr(fp.g(pre + '/kernel')), B, s, s, ci, co, KSIZE, 0, mm, ws, wsb, st)
            call('uad_convT2d_dgrad', ptr(g), ptr(fp.p(pre + '/kernel')), ptr(dst), B, s, s, ci, co, KSIZE, mm, ws, wsb, st)
        gam, bet = f'Generator/{_ln(0)}/gamma', f'Generator/{_ln(0)}/beta'
        call('uad_layernorm_hw_bwd', ptr(self.dzr), ptr(self.zr), ptr(self.g_stats_top), ptr(fp.p(gam)), ptr(fp.p(bet)),
             ptr(self.dzr), G(gam), G(bet), B, r2, ctop, ACT_RELU, 0.0, 0, ws, wsb, st)
        call('uad_dense_bwd', ptr(self.d), ptr(fp.p('Generator/conv2d_1/kernel')), ptr(self.dzr), None, 1.0, ptr(self.dd),
             G('Generator/conv2d_1/kernel'), None, B * r2, self.cb, ctop, 0, ws, wsb, st)
        call('uad_dense_bwd', ptr(self._g_in), ptr(fp.p(self.GEN_DENSE + '/kernel')), ptr(self.dd), ptr(self._g_mask), self._g_keep,
             ptr(dz_out), G(self.GEN_DENSE + '/kernel'), G(self.GEN_DENSE + '/bias'), B, self.zDim, self.flat, 0, ws, wsb, st)

    def _encoder_backward(self, dz_enc):
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self._wsp()
        B, r = self.B, self.res
        r2 = r * r
        ctop = self.enc_ch[-1]
        call('uad_activation_bwd', ptr(dz_enc), ptr(self.z_pre), ptr(dz_enc), dz_enc.numel(), ACT_TANH, 0.0, st)
        call('uad_dense_bwd', ptr(self.zb), ptr(fp.p('Encoder/dense/kernel')), ptr(dz_enc), ptr(self._e_mask), self._e_keep,
             ptr(self.dzb), ptr(fp.g('Encoder/dense/kernel')), ptr(fp.g('Encoder/dense/bias')), B, self.flat, self.zDim, 0, ws, wsb, st)
        self._encoder_stack_backward()

    def _encoder_stack_backward(self):
        """From dzb (gradient w.r.t. the 1x1 bottleneck conv's output) down through the frozen-BN conv blocks."""
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self._wsp()
        B, r = self.B, self.res
        r2 = r * r
        ctop = self.enc_ch[-1]
        call('uad_dense_bwd', ptr(self.enc_a[-1]), ptr(fp.p('Encoder/conv2d/kernel')), ptr(self.dzb), None, 1.0, ptr(self.enc_g[-1]),
             ptr(fp.g('Encoder/conv2d/kernel')), ptr(fp.g('Encoder/conv2d/bias')), B * r2, ctop, self.cb, 0, ws, wsb, st)
        n = len(self.enc_ch)
        sizes = [self.S >> k for k in range(n + 1)]
        for i in reversed(range(n)):
            co = self.enc_ch[i]
            cin = 1 if i == 0 else self.enc_ch[i - 1]
            s_in, s_out = sizes[i], sizes[i + 1]
            pre, bnn = f'Encoder/enc_conv2D_{i}', f'Encoder/{_bn(i)}'
            g = self.enc_g[i]
            call('uad_act_bn_bwd', ptr(g), ptr(self.enc_z[i]), ptr(fp.p(bnn + '/gamma')), ptr(fp.p(bnn + '/beta')), ptr(g),
                 ptr(fp.g(bnn + '/gamma')), ptr(fp.g(bnn + '/beta')), ptr(fp.g(pre + '/bias')), B * s_out * s_out, co, ACT_LEAKY,
                 LRELU_ALPHA, BN_C, 0, ws, wsb, st)
            src = self.x if i == 0 else self.enc_a[i - 1]
            call('uad_conv2d_wgrad', ptr(src), ptr(g), ptr(fp.g(pre + '/kernel')), B, s_in, s_in, cin, co, KSIZE, 0, mm, ws, wsb, st)
            if i > 0:
                call('uad_conv2d_dgrad', ptr(g), ptr(fp.p(pre + '/kernel')), ptr(self.enc_g[i - 1]), B, s_in, s_in, cin, co, KSIZE, mm,
                     ws, wsb, st)

    # ------------------------------------------------------------------ noise, optimiser
    # Each draw uses its own Philox sub-stream (id << 40) plus the device-resident counter, which advances once per train op:
    # nothing about the noise is baked into a captured CUDA graph.
    def draw_masks(self, rate, enc=False, gen=False):
        st = self._st()
        ctr = self.rng_ctr.data_ptr()
        if enc:
            call('uad_dropout_mask', ptr(self.mask_enc), self.mask_enc.numel(), float(rate), self.seed, 1 << 40, ctr, st)
        if gen:
            call('uad_dropout_mask', ptr(self.mask_gen), self.mask_gen.numel(), float(rate), self.seed, 2 << 40, ctr, st)

    def draw_alpha(self):
        call('uad_uniform', ptr(self.alpha), self.alpha.numel(), self.seed, 3 << 40, self.rng_ctr.data_ptr(), self._st())

    def _advance_rng(self):
        call('uad_counter_add', self.rng_ctr.data_ptr(), 1 << 20, self._st())

    def _run(self, name, key, body, use_graph):
        """Issue ``body`` (the kernel sequence of one train op): eagerly, or - with use_graph - eagerly once (warm-up), then
        captured into a CUDA graph and replayed (a critic step is ~320 launches; at the 16-slice per-GPU shard of
        BASELINE config 5 the op is launch-bound without this)."""
        if not use_graph:
            body()
            return
        k = (name, key)
        g = self._graphs.get(k)
        if g is None and self._warm.get(name) == key:
            t_save = dict(self.t)
            g = torch.cuda.CUDAGraph()
            with graph_capture(g):
                body()
            self.t = t_save                       # capture does not execute: undo the host-side step bookkeeping
            self._graphs[k] = g
        if g is not None:
            g.replay()
            for scope in self._graph_scopes.get(k, ()):
                self.t[scope] += 1
            return
        body()
        self._warm[name] = key

    def set_latent(self, z):
        if isinstance(z, np.ndarray):
            z = torch.from_numpy(np.ascontiguousarray(z, np.float32))
        self.z_in.copy_(z.reshape(self.z_in.shape), non_blocking=True)

    def _zero_grads(self, scope):
        lo, hi = self.fp.subset_ranges(scope + '/')
        call('uad_fill', ptr(self.fp.grads[lo:]), 0.0, hi - lo, self._st())

    def _update_in_graph(self, allreduce):
        """The optimiser step is part of the (capturable) train-op body unless it needs a host-side collective: single GPU, or
        the fused peer-memory kernel (enable_peer_optimizer)."""
        return allreduce is None or getattr(self, 'peer', None) is not None

    def enable_peer_optimizer(self):
        """Data parallel: `all-reduce of the scope's gradient slice + Adam` becomes ONE kernel over NVLink peer memory on that slice
        (csrc/uad_peer.cu, dist.PeerOptimizer).  Call once after the parameter broadcast, before the first train op."""
        from . import dist as udist
        if getattr(self, 'peer', None) is None and udist.world_size() > 1:
            self.peer = udist.PeerOptimizer(self.fp, self.device)
            self._graphs, self._warm = {}, {}
        return getattr(self, 'peer', None)

    def _adam(self, scope, lr, allreduce, world):
        """tf.train.AdamOptimizer(lr, beta1=0.5, beta2=0.9) on the scope's contiguous slice (fAnoGAN.py:71-77)."""
        fp, st = self.fp, self._st()
        lo, hi = fp.subset_ranges(scope + '/')
        if allreduce is not None and world > 1 and getattr(self, 'peer', None) is not None:
            self.t[scope] += 1
            call('uad_counter_add', self.steps[scope].data_ptr(), 1, st)
            self.peer.step(fp.m, fp.v, float(lr), 0.5, 0.9, 1e-8, 1.0 / world, self.steps[scope], st, lo=lo, hi=hi)
            return
        if allreduce is not None and world > 1:
            allreduce(fp.grads[lo:hi])
        self.t[scope] += 1
        call('uad_counter_add', self.steps[scope].data_ptr(), 1, st)
        call('uad_adam_tf_step', ptr(fp.params[lo:]), ptr(fp.grads[lo:]), ptr(fp.m[lo:]), ptr(fp.v[lo:]), hi - lo, float(lr), 0.5, 0.9,
             1e-8, 1.0 / world, self.steps[scope].data_ptr(), st)

    def _scalars(self, names):
        host = self.sc.cpu().numpy()
        return {k: float(host[self.SC[k]]) for k in names}

    def _mask_args(self, rate, dropout):
        on = bool(dropout) and rate > 0
        return on, (1.0 / (1.0 - rate) if on else 1.0)

    # ------------------------------------------------------------------ the three train ops
    def step_gen(self, lr, dropout_rate=0.0, dropout=True, parity_noise=False, allreduce=None, world=1, apply=True,
                 use_graph=False):
        """optim_gen: minimise gen_loss = -mean(D(G(z))) over the Generator variables (fAnoGAN.py:52,76,99-112)."""
        self.enable_training()
        on, keep = self._mask_args(dropout_rate, dropout)

        def body():
            st = self._st()
            ws, wsb = self._wsp()
            if on and not parity_noise:
                self.draw_masks(dropout_rate, gen=True)
                self._advance_rng()
            self.generate(self.z_in, self.mask_gen if on else None, keep, out=self.x_gen)
            _, d = self._critic_forward(self.pass0, self.x_gen)
            nd = d.numel()
            call('uad_sum_scaled', ptr(d), nd, 1.0 / nd, self.sc[0:].data_ptr(), ws, wsb, st)
            self._critic_top(self.pass0, -1.0 / nd, params=False)
            self._critic_backward(self.pass0, self.x_gen, params=False, dx_out=self.dxi)
            self._zero_grads('Generator')
            self._generator_backward(self.dxi, params=True)
            if apply and self._update_in_graph(allreduce):
                self._adam('Generator', lr, allreduce, world)

        key = (float(lr), float(dropout_rate), bool(dropout), bool(parity_noise), allreduce is None, world, bool(apply))
        self._graph_scopes[('gen', key)] = ('Generator',) if (apply and self._update_in_graph(allreduce)) else ()
        self._run('gen', key, body, use_graph)
        if apply and not self._update_in_graph(allreduce):
            self._adam('Generator', lr, allreduce, world)
        s = self._scalars(['disc_fake'])
        return {'gen_loss': -s['disc_fake'], 'disc_fake': s['disc_fake']}

    def step_disc(self, lr, dropout_rate=0.0, dropout=True, parity_noise=False, allreduce=None, world=1, apply=True,
                  use_graph=False):
        """optim_dis: minimise mean(D(x_)) - mean(D(x)) + gp over the Discriminator variables (fAnoGAN.py:50-58,75,114-129)."""
        self.enable_training()
        on, keep = self._mask_args(dropout_rate, dropout)

        def body():
            st = self._st()
            ws, wsb = self._wsp()
            if not parity_noise:
                if on:
                    self.draw_masks(dropout_rate, gen=True)
                self.draw_alpha()
                self._advance_rng()
            self.generate(self.z_in, self.mask_gen if on else None, keep, out=self.x_gen)
            self._zero_grads('Discriminator')
            _, d_f = self._critic_forward(self.pass0, self.x_gen)
            nd = d_f.numel()
            call('uad_sum_scaled', ptr(d_f), nd, 1.0 / nd, self.sc[0:].data_ptr(), ws, wsb, st)
            self._critic_top(self.pass0, 1.0 / nd, params=True)
            self._critic_backward(self.pass0, self.x_gen, params=True, dx_out=None)
            _, d_r = self._critic_forward(self.pass0, self.x)
            call('uad_sum_scaled', ptr(d_r), nd, 1.0 / nd, self.sc[1:].data_ptr(), ws, wsb, st)
            self._critic_top(self.pass0, -1.0 / nd, params=True)
            self._critic_backward(self.pass0, self.x, params=True, dx_out=None)
            call('uad_interpolate', ptr(self.x), ptr(self.x_gen), ptr(self.alpha), ptr(self.x_hat), self.B, self.S * self.S * self.C, st)
            self._critic_forward(self.pass0, self.x_hat, critic=False)
            self._critic_gp(self.pass0, self.x_hat)
            if apply and self._update_in_graph(allreduce):
                self._adam('Discriminator', lr, allreduce, world)

        key = (float(lr), float(dropout_rate), bool(dropout), bool(parity_noise), allreduce is None, world, bool(apply), self.scale)
        self._graph_scopes[('disc', key)] = ('Discriminator',) if (apply and self._update_in_graph(allreduce)) else ()
        self._run('disc', key, body, use_graph)
        if apply and not self._update_in_graph(allreduce):
            self._adam('Discriminator', lr, allreduce, world)
        s = self._scalars(['disc_fake', 'disc_real', 'gp'])
        s['disc_loss'] = s['disc_fake'] - s['disc_real'] + s['gp']
        return s

    def step_enc(self, lr, dropout_rate=0.0, dropout=True, parity_noise=False, allreduce=None, world=1, apply=True, train=True,
                 use_graph=False):
        """optim_enc: minimise mean((x-x_enc)^2) + kappa*mean((f(x_enc)-f(x))^2) over the Encoder variables
        (fAnoGAN.py:60-66,77,150-166).  train=False evaluates the losses only (validation loop, fAnoGAN.py:181-199)."""
        self.enable_training()
        on, keep = self._mask_args(dropout_rate, dropout)
        self._e_mask, self._e_keep = (self.mask_enc if on else None), keep
        kappa = self.kappa

        def body():
            st = self._st()
            ws, wsb = self._wsp()
            if on and not parity_noise:
                self.draw_masks(dropout_rate, enc=True, gen=True)
                self._advance_rng()
            z_enc = self.encode(self._e_mask, keep)
            x_enc = self.generate(z_enc, self.mask_gen if on else None, keep, out=self.x_enc)
            f_real, _ = self._critic_forward(self.pass1, self.x, critic=False)
            f_enc, _ = self._critic_forward(self.pass0, x_enc, critic=False)
            nx, nf = x_enc.numel(), f_enc.numel()
            # d enc_loss / d f_enc -> dis_g1[-1];  d loss_img / d x_enc -> dxi2 (added after the critic's input gradient)
            call('uad_mse', ptr(f_enc), ptr(f_real), nf, 2.0 * kappa / nf, ptr(self.dis_g1[-1]), 1.0 / nf, self.sc[4:].data_ptr(),
                 ws, wsb, st)
            call('uad_mse', ptr(x_enc), ptr(self.x), nx, 2.0 / nx, ptr(self.u), 1.0 / nx, self.sc[3:].data_ptr(), ws, wsb, st)
            call('uad_l1_map', ptr(self.x), ptr(x_enc), ptr(self.l1), ptr(self.rec), self.B, self.S * self.S * self.C, st)
            call('uad_sum_scaled', ptr(self.rec), self.B, 1.0 / self.B, self.sc[5:].data_ptr(), ws, wsb, st)
            if train:
                self._critic_backward(self.pass0, x_enc, params=False, dx_out=self.dxi)
                call('uad_axpby', 1.0, ptr(self.u), 1.0, ptr(self.dxi), nx, st)
                self._generator_backward(self.dxi, params=False, dz_out=self.dz_lat)
                self._encoder_backward(self.dz_lat)
                if apply and self._update_in_graph(allreduce):
                    self._adam('Encoder', lr, allreduce, world)

        key = (float(lr), float(dropout_rate), bool(dropout), bool(parity_noise), allreduce is None, world, bool(apply), bool(train),
               kappa)
        self._graph_scopes[('enc', key)] = ('Encoder',) if (train and apply and self._update_in_graph(allreduce)) else ()
        self._run('enc', key, body, use_graph)
        if train and apply and not self._update_in_graph(allreduce):
            self._adam('Encoder', lr, allreduce, world)
        s = self._scalars(['loss_img', 'loss_fts', 'reconstructionLoss'])
        s['enc_loss'] = s['loss_img'] + self.kappa * s['loss_fts']
        s['loss'] = s['reconstructionLoss']
        return s

    def wgan_scalars(self, dropout_rate=0.0, dropout=True, use_graph=False):
        """disc_real / disc_fake / gen_loss / disc_loss on the current x, z (the reference fetches **self.losses in the encoder
        phase, fAnoGAN.py:157,190; none of them depends on the Encoder, so evaluating them after its update is equivalent)."""
        return self.step_disc(0.0, dropout_rate, dropout, apply=False, use_graph=use_graph)
