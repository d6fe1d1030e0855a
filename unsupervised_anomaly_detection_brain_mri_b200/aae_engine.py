"""Device executor of the adversarial autoencoder (reference models/adversarial_autoencoder.py:10-73) and its three train ops
(trainers/AAE.py:41-69): the dense-AE engine (engine.ConvAutoencoderEngine, arch ``adversarial_autoencoder``) plus the latent
critic Dense(50, leaky_relu 0.2) -> Dense(50, leaky_relu 0.2) -> Dense(1) as `uad_dense_fwd` / `uad_dense_bwd` calls.

    step_ae    optim_ae : loss = mean_b mean_hwc (x - x_hat)^2            -> Encoder | Bottleneck | Decoder (the critic gets no gradient)
    step_disc  optim_dis: mean(D(z_)) - mean(D(z)) + gradient penalty     -> Discriminator
    step_gen   optim_gen: -mean(D(z_))                                    -> Encoder scope ONLY (AAE.py:64), own Adam slots
The gradient penalty mean((|d sum D(z_hat) / d z_hat|_2 - 1)^2 * scale) needs d/d theta of a gradient; as in fanogan_engine it is
taken without a tape: with u = d gp / d(ddz) fixed, grad_theta gp = grad_theta <J(theta) u, 1>, i.e. one tangent forward of the
MLP along u (Dense without bias, then * leaky'(pre-activation)) and its reverse pass.  The leaky ReLU is piecewise linear, so
the primal activations enter only through their sign patterns.

``constrained=True`` is the constrained adversarial AE (reference models/constrained_adversarial_autoencoder.py,
trainers/ConstrainedAAE.py:44-72): the constrained AE's two-pass graph and loss mean_b(L2 + rho * Rec_z) for optim_ae (the Dropout
calls on dec_dense(z_) and on z_rec carry no flag there: identity), a 100-50-1 critic, and an optim_gen whose 'Encoder' scope also
holds the 1x1 bottleneck conv and the latent Dense (the reference builds them inside that scope, :13-29) - so step_gen updates
Encoder/* + Bottleneck/conv2d + Bottleneck/dense (this repo's canonical names; two Adam ranges, one step counter).

STATUS: written after round 1's GPU budget was spent; the call sequences are verified on CPU against oracle/aae_cpu.py through
the ABI emulator (tests/test_engine_emulated.py); tests/test_gpu_aae.py is opt-in (UAD_UNVERIFIED=1) until its first hardware run."""
from __future__ import annotations

import numpy as np
import torch

from . import abi
from .abi import ACT_LEAKY, ACT_NONE, call, ptr
from .engine import AAE, BN_C, CAAE, CRITIC_WIDTHS, KSIZE, LRELU_ALPHA, ConvAutoencoderEngine, FlatParams, _bn, glorot_init, graph_capture, param_specs

CRITIC_ALPHA = 0.2          # tf.nn.leaky_relu default (adversarial_autoencoder.py:4,45-46)


class _CriticPass:
    """Pre-activations / activations of one pass through the latent critic."""

    def __init__(self, eng):
        self.pre = [eng._new(eng.B, w) for w in eng.widths[:-1]]
        self.act = [eng._new(eng.B, w) for w in eng.widths[:-1]]
        self.d = eng._new(eng.B, eng.widths[-1])


class AdversarialAEEngine(ConvAutoencoderEngine):
    SC = dict(reconstructionLoss=0, L2=4, disc_fake=8, disc_real=9, gp=10)
    OPS = ('ae', 'disc', 'gen')

    def __init__(self, S, C=1, zDim=128, res=8, batch=8, device='cuda:0', math_mode=abi.MATH_TC_3XTF32, seed=1, scale=10.0,
                 share_params=None, constrained=False, rho=1.0):
        super().__init__(CAAE if constrained else AAE, S, C, zDim, res, batch, device, math_mode, seed, share_params=share_params)
        self.constrained = bool(constrained)
        if constrained:
            self.rho = float(rho)
        self.widths = CRITIC_WIDTHS[self.arch]
        self.scale = float(scale)
        B = self.B
        self.scalars = torch.zeros(16, dtype=torch.float32, device=self.device)
        self.z_real = self._new(B, zDim)                 # the fed z ~ N(0,1) (AAE.get_feed_dict)
        self.epsilon = self._new(B)                      # tf.random_uniform of adversarial_autoencoder.py:64 (stored NEGATED, see step_disc)
        self.z_hat = self._new(B, zDim)
        self.cp = _CriticPass(self)
        self.tan = [self._new(B, w) for w in self.widths[:-1]]     # tangent activations of the gradient-penalty pass
        self.tpre = [self._new(B, w) for w in self.widths[:-1]]
        self.gh = [self._new(B, w) for w in self.widths[:-1]]      # gradient scratch per hidden layer
        self.dd = self._new(B, 1)
        self.ddz = self._new(B, zDim)
        self.u = self._new(B, zDim)
        self.dz_lat = self._new(B, zDim)
        fp = self.fp
        self.rng = {op: fp.subset_ranges(p) for op, p in (('disc', 'Discriminator/'), ('gen', 'Encoder/'))}
        self.rng['ae'] = (0, self.rng['disc'][0])        # Encoder | Bottleneck | Decoder precede the critic in the layout
        self.rngs = {op: [r] for op, r in self.rng.items()}                # the slices an op's Adam touches (ascending)
        if constrained:                                  # optim_gen: Encoder/* + the 1x1 bottleneck conv + the latent Dense
            self.rngs['gen'] += [fp.subset_ranges('Bottleneck/conv2d/'), fp.subset_ranges('Bottleneck/dense/')]
            self.rng['gen'] = (self.rngs['gen'][0][0], self.rngs['gen'][-1][1])
        lo, hi = self.rng['gen']
        self.m_gen = torch.zeros(hi - lo, dtype=torch.float32, device=self.device)    # optim_gen's own Adam slots
        self.v_gen = torch.zeros(hi - lo, dtype=torch.float32, device=self.device)
        self.op_steps = {op: torch.zeros(1, dtype=torch.int64, device=self.device) for op in self.OPS}
        self.op_t = {op: 0 for op in self.OPS}
        need = max(self.ws_bytes, max(abi.lib().uad_dense_workspace_bytes(B, k, n) for k, n in self._critic_dims()),
                   abi.lib().uad_reduce_workspace_bytes(), B * 4)
        if need > self.ws_bytes:
            self.ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self.ws_bytes = need
        self._graphs, self._warmed = {}, {}

    def _critic_dims(self):
        k, dims = self.zDim, []
        for w in self.widths:
            dims.append((k, w))
            k = w
        return dims

    def _critic_names(self):
        return [f'Discriminator/dense_{2 + j}' for j in range(len(self.widths))]

    def set_latent(self, z):
        if isinstance(z, np.ndarray):
            z = torch.from_numpy(np.ascontiguousarray(z, np.float32))
        self.z_real.copy_(z.reshape(self.z_real.shape), non_blocking=True)

    # ------------------------------------------------------------------ encoder -> z_ only (what optim_dis / optim_gen evaluate)
    def encode_latent(self, training=True):
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self._wsargs()
        B, br = self.B, self.br[0]
        h, s, cin = br.x, self.S, 1
        for i, co in enumerate(self.enc_ch):
            pre, bnn = f'Encoder/enc_conv2D_{i}', f'Encoder/{_bn(i)}'
            call('uad_conv2d_fwd', ptr(h), ptr(fp.p(pre + '/kernel')), ptr(fp.p(pre + '/bias')), ptr(fp.p(bnn + '/gamma')),
                 ptr(fp.p(bnn + '/beta')), ptr(br.enc_z[i]) if training and self.keep_preact else None, ptr(br.enc_a[i]), B, s, s, cin, co,
                 KSIZE, ACT_LEAKY, LRELU_ALPHA, BN_C, mm, ws, wsb, st)
            h, s, cin = br.enc_a[i], s // 2, co
        r2 = self.res * self.res
        call('uad_dense_fwd', ptr(h), ptr(fp.p('Bottleneck/conv2d/kernel')), ptr(fp.p('Bottleneck/conv2d/bias')), None, 1.0, None, None,
             ptr(br.zb), None, B * r2, cin, self.cb, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        call('uad_dense_fwd', ptr(br.zb), ptr(fp.p('Bottleneck/dense/kernel')), ptr(fp.p('Bottleneck/dense/bias')), ptr(br.masks['mu']),
             self._keep, None, None, ptr(br.mu), None, B, self.flat, self.zDim, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        return br.mu

    # ------------------------------------------------------------------ the latent critic
    def critic_forward(self, z_dev):
        fp, st = self.fp, self._st()
        ws, wsb = self._wsargs()
        cp, h = self.cp, z_dev
        names, dims = self._critic_names(), self._critic_dims()
        for j, (name, (k, n)) in enumerate(zip(names, dims)):
            last = j == len(names) - 1
            call('uad_dense_fwd', ptr(h), ptr(fp.p(name + '/kernel')), ptr(fp.p(name + '/bias')), None, 1.0, None, None,
                 ptr(cp.d if last else cp.pre[j]), None if last else ptr(cp.act[j]), self.B, k, n, ACT_NONE if last else ACT_LEAKY,
                 CRITIC_ALPHA, 1.0, ws, wsb, st)
            h = None if last else cp.act[j]
        return cp.d

    def critic_backward(self, z_in, coef, params, dz_out):
        """Reverse pass of the last critic_forward (whose input was z_in) from d/dD = coef per sample.  params: accumulate the
        critic's parameter gradients; dz_out (nullable): gradient w.r.t. the critic's input."""
        fp, st = self.fp, self._st()
        ws, wsb = self._wsargs()
        cp, B = self.cp, self.B
        names, dims = self._critic_names(), self._critic_dims()
        G = (lambda n: ptr(fp.g(n))) if params else (lambda n: None)
        call('uad_fill', ptr(self.dd), float(coef), B, st)
        g = self.dd
        for j in reversed(range(len(names))):
            k, n = dims[j]
            x_in = z_in if j == 0 else cp.act[j - 1]
            dx = dz_out if j == 0 else self.gh[j - 1]
            call('uad_dense_bwd', ptr(x_in), ptr(fp.p(names[j] + '/kernel')), ptr(g), None, 1.0, ptr(dx), G(names[j] + '/kernel'),
                 G(names[j] + '/bias'), B, k, n, 1, ws, wsb, st)
            if j > 0:
                call('uad_activation_bwd', ptr(dx), ptr(cp.pre[j - 1]), ptr(dx), dx.numel(), ACT_LEAKY, CRITIC_ALPHA, st)
                g = dx

    def critic_gp(self, z_hat):
        """Gradient penalty on the already-forwarded z_hat pass: gp -> scalars[10]; d gp / d theta accumulated into the critic's
        kernel gradients (its biases do not enter d D / d z)."""
        fp, st = self.fp, self._st()
        ws, wsb = self._wsargs()
        cp, B = self.cp, self.B
        names, dims = self._critic_names(), self._critic_dims()
        self.critic_backward(z_hat, 1.0, params=False, dz_out=self.ddz)                    # ddz = d sum(D) / d z_hat
        call('uad_gradient_penalty', ptr(self.ddz), B, self.zDim, 1, self.scale, ptr(self.u), self.scalars[10:].data_ptr(), ws, wsb, st)
        h = self.u                                                                         # tangent forward along u
        for j in range(len(names) - 1):
            k, n = dims[j]
            call('uad_dense_fwd', ptr(h), ptr(fp.p(names[j] + '/kernel')), None, None, 1.0, None, None, ptr(self.tpre[j]), None, B, k, n,
                 ACT_NONE, 0.0, 1.0, ws, wsb, st)
            call('uad_activation_bwd', ptr(self.tpre[j]), ptr(cp.pre[j]), ptr(self.tan[j]), self.tan[j].numel(), ACT_LEAKY, CRITIC_ALPHA, st)
            h = self.tan[j]
        call('uad_fill', ptr(self.dd), 1.0, B, st)                                          # reverse of s = sum(tan_last . W_last)
        g = self.dd
        for j in reversed(range(len(names))):
            k, n = dims[j]
            x_in = self.u if j == 0 else self.tan[j - 1]
            dx = None if j == 0 else self.gh[j - 1]
            call('uad_dense_bwd', ptr(x_in), ptr(fp.p(names[j] + '/kernel')), ptr(g), None, 1.0, ptr(dx), ptr(fp.g(names[j] + '/kernel')),
                 None, B, k, n, 1, ws, wsb, st)
            if j > 0:
                call('uad_activation_bwd', ptr(dx), ptr(cp.pre[j - 1]), ptr(dx), dx.numel(), ACT_LEAKY, CRITIC_ALPHA, st)
                g = dx

    # ------------------------------------------------------------------ noise / optimiser / launch
    def draw_epsilon(self):
        st = self._st()
        call('uad_uniform', ptr(self.epsilon), self.B, self.rng_seed, 7 << 40, self.rng_ctr.data_ptr(), st)
        call('uad_axpby', 0.0, ptr(self.epsilon), -1.0, ptr(self.epsilon), self.B, st)     # keep -epsilon (see step_disc)

    def set_epsilon(self, eps):
        self.epsilon.copy_(-torch.as_tensor(np.asarray(eps, np.float32)).reshape(-1))

    def _zero(self, op):
        lo, hi = self.rng[op]
        call('uad_fill', ptr(self.fp.grads[lo:]), 0.0, hi - lo, self._st())

    def _adam(self, op, lr, allreduce, world):
        fp, st = self.fp, self._st()
        base = self.rng[op][0]
        self.op_t[op] += 1
        call('uad_counter_add', self.op_steps[op].data_ptr(), 1, st)
        for lo, hi in self.rngs[op]:
            if allreduce is not None and world > 1:
                allreduce(fp.grads[lo:hi])
            m, v = (self.m_gen[lo - base:], self.v_gen[lo - base:]) if op == 'gen' else (fp.m[lo:], fp.v[lo:])
            call('uad_adam_tf_step', ptr(fp.params[lo:]), ptr(fp.grads[lo:]), ptr(m), ptr(v), hi - lo, float(lr), 0.5, 0.9, 1e-8, 1.0 / world,
                 self.op_steps[op].data_ptr(), st)

    def _launch(self, op, key, body, use_graph, in_graph_adam):
        """Eager, or (use_graph) eager once, then captured into a CUDA graph and replayed - as fanogan_engine._run."""
        if not use_graph:
            body()
            return
        k = (op, key)
        g = self._graphs.get(k)
        if g is None and self._warmed.get(op) == key:
            t_save = dict(self.op_t)
            g = torch.cuda.CUDAGraph()
            with graph_capture(g):
                body()
            self.op_t = t_save
            self._graphs[k] = g
        if g is not None:
            g.replay()
            if in_graph_adam:
                self.op_t[op] += 1
            return
        body()
        self._warmed[op] = key

    def _noise(self, dropout, rate, parity_noise, epsilon=False):
        self._keep = 1.0 / (1.0 - rate) if (dropout and rate > 0) else 1.0
        if not parity_noise:
            if epsilon:
                self.draw_epsilon()
            self.draw_noise(dropout, rate)                  # masks z / dec; advances the Philox counter

    def _sc(self, names):
        host = self.scalars.cpu().numpy()
        return {k: float(host[self.SC[k]]) for k in names}

    # ------------------------------------------------------------------ the three train ops
    def step_ae(self, lr, dropout_rate=0.0, dropout=True, parity_noise=False, allreduce=None, world=1, apply=True, train=True,
                use_graph=False):
        """optim_ae (AAE.py:55-57,67); train=False: the reconstruction fetches of the validation loop."""
        rate = dropout_rate if dropout else 0.0

        def body():
            self._noise(dropout, rate, parity_noise)
            self.forward(training=train, dropout_rate=rate)
            if train:
                if self.constrained:
                    self.backward_constrained()
                else:
                    self.backward_from_gxhat()
                if apply and allreduce is None:
                    self._adam('ae', lr, None, world)

        key = (float(lr), rate, bool(parity_noise), allreduce is None, world, bool(apply), bool(train))
        self._launch('ae', key, body, use_graph, train and apply and allreduce is None)
        if train and apply and allreduce is not None:
            self._adam('ae', lr, allreduce, world)
        s = self._sc(['reconstructionLoss', 'L2'])
        s['loss'] = s['L2']
        if self.constrained:
            s['Rec_z'] = float(self.scalars[5])
            s['loss'] = s['L2'] + self.rho * s['Rec_z']
        return s

    def step_disc(self, lr, dropout_rate=0.0, dropout=True, parity_noise=False, allreduce=None, world=1, apply=True, use_graph=False):
        """optim_dis (AAE.py:41-52,66).  z_hat = z + epsilon * (z - z_) as the reference writes it, i.e. x + a * (x_gen - x) of
        uad_interpolate with a = -epsilon (the buffer holds the negated draw)."""
        rate = dropout_rate if dropout else 0.0

        def body():
            st = self._st()
            ws, wsb = self._wsargs()
            B = self.B
            self._noise(dropout, rate, parity_noise, epsilon=True)
            z_ = self.encode_latent()
            self._zero('disc')
            d = self.critic_forward(z_)
            call('uad_sum_scaled', ptr(d), B, 1.0 / B, self.scalars[8:].data_ptr(), ws, wsb, st)
            self.critic_backward(z_, 1.0 / B, params=True, dz_out=None)
            d = self.critic_forward(self.z_real)
            call('uad_sum_scaled', ptr(d), B, 1.0 / B, self.scalars[9:].data_ptr(), ws, wsb, st)
            self.critic_backward(self.z_real, -1.0 / B, params=True, dz_out=None)
            call('uad_interpolate', ptr(self.z_real), ptr(z_), ptr(self.epsilon), ptr(self.z_hat), B, self.zDim, st)
            self.critic_forward(self.z_hat)
            self.critic_gp(self.z_hat)
            if apply and allreduce is None:
                self._adam('disc', lr, None, world)

        key = (float(lr), rate, bool(parity_noise), allreduce is None, world, bool(apply), self.scale)
        self._launch('disc', key, body, use_graph, apply and allreduce is None)
        if apply and allreduce is not None:
            self._adam('disc', lr, allreduce, world)
        s = self._sc(['disc_fake', 'disc_real', 'gp'])
        s['gen_loss'] = -s['disc_fake']
        s['disc_loss_without_grad'] = s['disc_fake'] - s['disc_real']
        s['disc_loss'] = s['disc_loss_without_grad'] + s['gp']
        return s

    def step_gen(self, lr, dropout_rate=0.0, dropout=True, parity_noise=False, allreduce=None, world=1, apply=True, use_graph=False):
        """optim_gen (AAE.py:43,64,68): -mean(D(z_)) through the critic, the latent Dense (+ Dropout) and the 1x1 bottleneck conv
        into the conv encoder; only the Encoder scope is updated (the Bottleneck gradients this pass forms are discarded)."""
        rate = dropout_rate if dropout else 0.0

        def body():
            fp, st = self.fp, self._st()
            ws, wsb = self._wsargs()
            B, br, sm = self.B, self.br[0], self.small
            self._noise(dropout, rate, parity_noise)
            z_ = self.encode_latent()
            d = self.critic_forward(z_)
            call('uad_sum_scaled', ptr(d), B, 1.0 / B, self.scalars[8:].data_ptr(), ws, wsb, st)
            self.critic_backward(z_, -1.0 / B, params=False, dz_out=self.dz_lat)
            g, gn = self.gbuf
            G = (lambda n: ptr(fp.g(n))) if self.constrained else (lambda n: None)      # constrained: these two layers are updated too
            call('uad_dense_bwd', ptr(br.zb), ptr(fp.p('Bottleneck/dense/kernel')), ptr(self.dz_lat), ptr(br.masks['mu']), self._keep,
                 ptr(sm['dflat']), G('Bottleneck/dense/kernel'), G('Bottleneck/dense/bias'), B, self.flat, self.zDim, 0, ws, wsb, st)
            call('uad_dense_bwd', ptr(br.enc_a[-1]), ptr(fp.p('Bottleneck/conv2d/kernel')), ptr(sm['dflat']), None, 1.0, ptr(g),
                 G('Bottleneck/conv2d/kernel'), G('Bottleneck/conv2d/bias'), B * self.res * self.res, self.enc_ch[-1], self.cb, 0, ws, wsb, st)
            self._encoder_backward(br, g, gn, 0, None)
            if apply and allreduce is None:
                self._adam('gen', lr, None, world)

        key = (float(lr), rate, bool(parity_noise), allreduce is None, world, bool(apply))
        self._launch('gen', key, body, use_graph, apply and allreduce is None)
        if apply and allreduce is not None:
            self._adam('gen', lr, allreduce, world)
        s = self._sc(['disc_fake'])
        return {'gen_loss': -s['disc_fake'], 'disc_fake': s['disc_fake']}


def make_params(S, C=1, zDim=128, res=8, device='cuda:0', seed=1, constrained=False):
    """A FlatParams for the (constrained) AAE variable set (what share_params expects), Glorot-initialised."""
    specs = param_specs(CAAE if constrained else AAE, S, C, zDim, res)
    fp = FlatParams(specs, torch.device(device))
    fp.load(glorot_init(specs, seed))
    return fp
