"""Device executor of the AnoVAEGAN graph (reference models/anovaegan.py:10-83) and its three train ops
(trainers/AnoVAEGAN.py:50-83) - a composition of kernels the f-AnoGAN and VAE engines already run:

    Encoder    unified encoder (frozen BN) -> 1x1 conv -> Dense mu / Dense log-sigma (+ dropout) -> z = mu + eps * exp(ls)
    Generator  Dense (+ dropout) -> 1x1 conv -> unified decoder with LayerNormalization([1,2]); NO output non-linearity
    Discriminator  unified encoder with LayerNormalization + Dense(1) on the channel axis, on out / x / x_hat

    step_vae   optim_vae: enc_loss = mean_b sum|x - out| + kl_weight * mean_b kl    over Encoder + Generator
    step_gen   optim_gen: gen_loss = -mean(D(out))                                  over Generator
    step_disc  optim_dis: WGAN-GP critic loss (tape-free second-order term, see fanogan_engine)   over Discriminator
Each op is its own tf.train.AdamOptimizer(lr, 0.5, 0.9): the Generator slice owns a second pair of Adam slots for optim_gen.

STATUS: written after round 1's GPU budget was spent - checked on CPU only (oracle/anovaegan_cpu.py restates the graph; the
host bookkeeping is unit-tested); tests/test_gpu_anovaegan.py is opt-in (UAD_UNVERIFIED=1) until its first hardware run."""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import abi
from .abi import ACT_LEAKY, ACT_NONE, call, ptr
from .engine import BN_C, KSIZE, LRELU_ALPHA, stack_plan
from .fanogan_engine import FanoganEngine, _bn, _ln


def param_specs(S, C=1, zDim=128, res=8):
    """TF variable names -> shapes, scope-contiguous (Encoder | Generator | Discriminator); Dense / Conv2D name counters run over
    the whole graph: Encoder/{conv2d, dense, dense_1}, Generator/{conv2d_1, dense_2}, Discriminator/dense_3."""
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
    for name in ('Encoder/dense', 'Encoder/dense_1'):
        sp[name + '/kernel'] = (flat, zDim)
        sp[name + '/bias'] = (zDim,)
    sp['Generator/conv2d_1/kernel'] = (1, 1, cb, cin)
    sp['Generator/conv2d_1/bias'] = (cin,)
    sp['Generator/dense_2/kernel'] = (zDim, flat)
    sp['Generator/dense_2/bias'] = (flat,)
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
    sp['Discriminator/dense_3/kernel'] = (cin, 1)
    sp['Discriminator/dense_3/bias'] = (1,)
    return sp


# train op -> (scopes it updates, whose Adam slots it uses); optim_gen's slots for the Generator are separate from optim_vae's
OPS = OrderedDict(vae=('Encoder', 'Generator'), gen=('Generator',), disc=('Discriminator',))


class AnoVaeGanEngine(FanoganEngine):
    SC = dict(disc_fake=0, disc_real=1, gp=2, loss_img=3, loss_fts=4, reconstructionLoss=5, kl=6)
    GEN_DENSE = 'Generator/dense_2'
    DISC_DENSE = 'Discriminator/dense_3'
    FINAL_ACT = ACT_NONE

    def __init__(self, S, C=1, zDim=128, res=8, batch=8, device='cuda:0', math_mode=abi.MATH_TC_3XTF32, seed=1, kl_weight=1.0,
                 scale=10.0):
        super().__init__(S, C, zDim, res, batch, device, math_mode, seed, kappa=1.0, scale=scale)
        self.kl_weight = float(kl_weight)
        B = self.B
        self.mu, self.ls, self.sigma, self.zv = (self._new(B, zDim) for _ in range(4))
        self.eps = torch.zeros(B, zDim, dtype=torch.float32, device=self.device)
        self.kl = self._new(B)

    @staticmethod
    def _param_specs(S, C, zDim, res):
        return param_specs(S, C, zDim, res)

    def op_range(self, op):
        """[lo, hi) of the flat buffer a train op updates (its scopes are adjacent in the layout)."""
        lo = min(self.fp.subset_ranges(s + '/')[0] for s in OPS[op])
        hi = max(self.fp.subset_ranges(s + '/')[1] for s in OPS[op])
        return lo, hi

    def enable_training(self):
        if self._train_ready:
            return
        super().enable_training()
        B = self.B
        self.mask_mu = self._new(B, self.zDim)
        self.mask_ls = self._new(B, self.zDim)
        self.dmu, self.dls = self._new(B, self.zDim), self._new(B, self.zDim)
        self.dzb2 = self._new(B, self.res, self.res, self.cb)
        lo, hi = self.op_range('gen')
        self.m_gen = torch.zeros(hi - lo, dtype=torch.float32, device=self.device)     # optim_gen's own Adam slots
        self.v_gen = torch.zeros(hi - lo, dtype=torch.float32, device=self.device)
        self.steps = {k: torch.zeros(1, dtype=torch.int64, device=self.device) for k in OPS}
        self.t = {k: 0 for k in OPS}

    def set_noise(self, eps):
        """Parity aid: the N(0,1) draw of anovaegan.py:35 supplied by the caller (used with parity_noise=True)."""
        if not isinstance(eps, torch.Tensor):
            import numpy as np
            eps = torch.from_numpy(np.ascontiguousarray(eps, np.float32))
        self.eps.copy_(eps.reshape(self.eps.shape), non_blocking=True)

    # ------------------------------------------------------------------ forward
    def encode_vae(self, mask_mu=None, mask_ls=None, keep=1.0):
        """x -> z_mu, z_log_sigma (dropout on both, anovaegan.py:32-33), z_sigma, z_vae = z_mu + eps * z_sigma, kl per sample."""
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
        call('uad_dense_fwd', ptr(self.zb), ptr(fp.p('Encoder/dense/kernel')), ptr(fp.p('Encoder/dense/bias')), ptr(mask_mu), keep,
             None, None, ptr(self.mu), None, B, self.flat, self.zDim, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        call('uad_dense_fwd', ptr(self.zb), ptr(fp.p('Encoder/dense_1/kernel')), ptr(fp.p('Encoder/dense_1/bias')), ptr(mask_ls), keep,
             None, None, ptr(self.ls), None, B, self.flat, self.zDim, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        call('uad_reparam_kl_fwd', ptr(self.mu), ptr(self.ls), ptr(self.eps), ptr(self.sigma), ptr(self.zv), ptr(self.kl), B, self.zDim, st)
        self._e_masks, self._e_keep = (mask_mu, mask_ls), keep
        return self.zv

    def reconstruct(self, masks=(None, None, None), keep=1.0, out=None):
        """out = G(z_vae(E(x))); the caller refreshes eps (the graph's tf.random_normal is live at inference too)."""
        return self.generate(self.encode_vae(masks[0], masks[1], keep), masks[2], keep, out=out)

    def _forward_out(self, on, keep):
        m = (self.mask_mu, self.mask_ls, self.mask_gen) if on else (None, None, None)
        return self.reconstruct(m, keep, out=self.x_gen)

    # ------------------------------------------------------------------ noise, optimiser
    def draw_noise(self, rate, on, alpha=False):
        st = self._st()
        ctr = self.rng_ctr.data_ptr()
        call('uad_randn', ptr(self.eps), self.eps.numel(), self.seed, 4 << 40, ctr, st)
        if on:
            call('uad_dropout_mask', ptr(self.mask_mu), self.mask_mu.numel(), float(rate), self.seed, 5 << 40, ctr, st)
            call('uad_dropout_mask', ptr(self.mask_ls), self.mask_ls.numel(), float(rate), self.seed, 6 << 40, ctr, st)
            call('uad_dropout_mask', ptr(self.mask_gen), self.mask_gen.numel(), float(rate), self.seed, 2 << 40, ctr, st)
        if alpha:
            self.draw_alpha()
        self._advance_rng()

    def _zero_op_grads(self, op):
        lo, hi = self.op_range(op)
        call('uad_fill', ptr(self.fp.grads[lo:]), 0.0, hi - lo, self._st())

    def _adam_op(self, op, lr, allreduce, world):
        fp, st = self.fp, self._st()
        lo, hi = self.op_range(op)
        if allreduce is not None and world > 1:
            allreduce(fp.grads[lo:hi])
        self.t[op] += 1
        call('uad_counter_add', self.steps[op].data_ptr(), 1, st)
        m, v = (self.m_gen, self.v_gen) if op == 'gen' else (fp.m[lo:], fp.v[lo:])
        call('uad_adam_tf_step', ptr(fp.params[lo:]), ptr(fp.grads[lo:]), ptr(m), ptr(v), hi - lo, float(lr), 0.5, 0.9, 1e-8, 1.0 / world,
             self.steps[op].data_ptr(), st)

    def _launch(self, op, key, body, use_graph, apply, lr, allreduce, world):
        in_graph = apply and allreduce is None
        self._graph_scopes[(op, key)] = (op,) if in_graph else ()
        self._run(op, key, body, use_graph)
        if apply and allreduce is not None:
            self._adam_op(op, lr, allreduce, world)

    # ------------------------------------------------------------------ the three train ops
    def step_vae(self, lr, dropout_rate=0.0, dropout=True, parity_noise=False, allreduce=None, world=1, apply=True, train=True,
                 use_graph=False):
        """optim_vae (AnoVAEGAN.py:58-71,83,97-110); train=False evaluates the fetches of the validation loop (:166-186)."""
        self.enable_training()
        on, keep = self._mask_args(dropout_rate, dropout)
        klw = self.kl_weight

        def body():
            fp, st = self.fp, self._st()
            ws, wsb = self._wsp()
            B, HW = self.B, self.S * self.S
            if not parity_noise:
                self.draw_noise(dropout_rate, on)
            out = self._forward_out(on, keep)
            call('uad_l1_map', ptr(self.x), ptr(out), ptr(self.l1), ptr(self.rec), B, HW * self.C, st)
            call('uad_sum_scaled', ptr(self.rec), B, 1.0 / B, self.sc[5:].data_ptr(), ws, wsb, st)
            call('uad_sum_scaled', ptr(self.kl), B, 1.0 / B, self.sc[6:].data_ptr(), ws, wsb, st)
            if not train:
                return
            self._zero_op_grads('vae')
            # d enc_loss / d out = sign(out - x) / B, folded into the final 1x1 conv's backward
            call('uad_final1x1_l1_bwd', ptr(self.gen_a[-1]), ptr(fp.p('Generator/dec_Conv2D_final/kernel')), ptr(self.x), ptr(out),
                 1.0 / B, ptr(self.gen_g[-1]), ptr(fp.g('Generator/dec_Conv2D_final/kernel')),
                 ptr(fp.g('Generator/dec_Conv2D_final/bias')), B, HW, self.dec_ch[-1], 0, ws, wsb, st)
            self._generator_backward(None, params=True, dz_out=self.dz_lat, head=False)
            call('uad_reparam_kl_bwd', ptr(self.mu), ptr(self.ls), ptr(self.eps), ptr(self.dz_lat), klw / B, ptr(self.dmu), ptr(self.dls),
                 B, self.zDim, st)
            mm_, ml_ = self._e_masks
            call('uad_dense_bwd', ptr(self.zb), ptr(fp.p('Encoder/dense/kernel')), ptr(self.dmu), ptr(mm_), self._e_keep, ptr(self.dzb),
                 ptr(fp.g('Encoder/dense/kernel')), ptr(fp.g('Encoder/dense/bias')), B, self.flat, self.zDim, 0, ws, wsb, st)
            call('uad_dense_bwd', ptr(self.zb), ptr(fp.p('Encoder/dense_1/kernel')), ptr(self.dls), ptr(ml_), self._e_keep, ptr(self.dzb2),
                 ptr(fp.g('Encoder/dense_1/kernel')), ptr(fp.g('Encoder/dense_1/bias')), B, self.flat, self.zDim, 0, ws, wsb, st)
            call('uad_axpby', 1.0, ptr(self.dzb2), 1.0, ptr(self.dzb), self.dzb.numel(), st)
            self._encoder_stack_backward()
            if apply and allreduce is None:
                self._adam_op('vae', lr, None, world)

        key = (float(lr), float(dropout_rate), bool(dropout), bool(parity_noise), allreduce is None, world, bool(apply), bool(train), klw)
        self._launch('vae', key, body, use_graph, apply and train, lr, allreduce, world)
        s = self._scalars(['reconstructionLoss', 'kl'])
        s['loss'] = s['reconstructionLoss']
        s['enc_loss'] = s['reconstructionLoss'] + klw * s['kl']
        return s

    def step_gen(self, lr, dropout_rate=0.0, dropout=True, parity_noise=False, allreduce=None, world=1, apply=True, use_graph=False):
        """optim_gen (AnoVAEGAN.py:50,70,82,112-122): -mean(D(out)) w.r.t. the Generator variables; the Encoder only feeds z."""
        self.enable_training()
        on, keep = self._mask_args(dropout_rate, dropout)

        def body():
            st = self._st()
            ws, wsb = self._wsp()
            if not parity_noise:
                self.draw_noise(dropout_rate, on)
            out = self._forward_out(on, keep)
            _, d = self._critic_forward(self.pass0, out)
            nd = d.numel()
            call('uad_sum_scaled', ptr(d), nd, 1.0 / nd, self.sc[0:].data_ptr(), ws, wsb, st)
            self._critic_top(self.pass0, -1.0 / nd, params=False)
            self._critic_backward(self.pass0, out, params=False, dx_out=self.dxi)
            self._zero_op_grads('gen')
            self._generator_backward(self.dxi, params=True)
            if apply and allreduce is None:
                self._adam_op('gen', lr, None, world)

        key = (float(lr), float(dropout_rate), bool(dropout), bool(parity_noise), allreduce is None, world, bool(apply))
        self._launch('gen', key, body, use_graph, apply, lr, allreduce, world)
        s = self._scalars(['disc_fake'])
        return {'gen_loss': -s['disc_fake'], 'disc_fake': s['disc_fake']}

    def step_disc(self, lr, dropout_rate=0.0, dropout=True, parity_noise=False, allreduce=None, world=1, apply=True, use_graph=False):
        """optim_dis (AnoVAEGAN.py:50-57,81,124-148): WGAN-GP critic step with fake = out, x_hat = x + alpha * (out - x)."""
        self.enable_training()
        on, keep = self._mask_args(dropout_rate, dropout)

        def body():
            st = self._st()
            ws, wsb = self._wsp()
            if not parity_noise:
                self.draw_noise(dropout_rate, on, alpha=True)
            out = self._forward_out(on, keep)
            self._zero_op_grads('disc')
            _, d_f = self._critic_forward(self.pass0, out)
            nd = d_f.numel()
            call('uad_sum_scaled', ptr(d_f), nd, 1.0 / nd, self.sc[0:].data_ptr(), ws, wsb, st)
            self._critic_top(self.pass0, 1.0 / nd, params=True)
            self._critic_backward(self.pass0, out, params=True, dx_out=None)
            _, d_r = self._critic_forward(self.pass0, self.x)
            call('uad_sum_scaled', ptr(d_r), nd, 1.0 / nd, self.sc[1:].data_ptr(), ws, wsb, st)
            self._critic_top(self.pass0, -1.0 / nd, params=True)
            self._critic_backward(self.pass0, self.x, params=True, dx_out=None)
            call('uad_interpolate', ptr(self.x), ptr(out), ptr(self.alpha), ptr(self.x_hat), self.B, self.S * self.S * self.C, st)
            self._critic_forward(self.pass0, self.x_hat, critic=False)
            self._critic_gp(self.pass0, self.x_hat)
            if apply and allreduce is None:
                self._adam_op('disc', lr, None, world)

        key = (float(lr), float(dropout_rate), bool(dropout), bool(parity_noise), allreduce is None, world, bool(apply), self.scale)
        self._launch('disc', key, body, use_graph, apply, lr, allreduce, world)
        s = self._scalars(['disc_fake', 'disc_real', 'gp'])
        s['disc_loss'] = s['disc_fake'] - s['disc_real'] + s['gp']
        return s

    def step_enc(self, *a, **k):
        raise NotImplementedError('AnoVAEGAN has no izi_f encoder phase (that is f-AnoGAN); see step_vae')
