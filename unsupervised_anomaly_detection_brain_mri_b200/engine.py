"""Device-side executor of the conv encoder->decoder hot path (AE / VAE / ceVAE graphs).

This is what replaces ``sess.run`` of the reference (trainers/AE.py:83, VAE.py:96, ceVAE.py:107): one object that owns
the flat parameter / gradient / Adam buffers and the activation buffers in HBM and walks the layer list, issuing one
C-ABI call (``libuad_b200.so``) per fused block.  PyTorch supplies device memory and streams only - no torch math op
is on this path.

HBM layout
  * params / grads / adam-m / adam-v: four flat fp32 buffers, one slot per TF variable (256-byte aligned slots),
    so the optimiser is ONE fused kernel and data-parallel training is ONE all-reduce.
  * activations: NHWC fp32, per conv block the pre-BN output ``z`` (kept for backward) and the activation ``a``
    (consumed by the next block); gradients ping-pong between two max-sized buffers.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from . import abi
from .abi import (ACT_LEAKY, ACT_NONE, ACT_RELU, OP_CONV_DGRAD, OP_CONV_FWD, OP_CONV_WGRAD, OP_CONVT_DGRAD,
                  OP_CONVT_FWD, OP_CONVT_WGRAD, call, ptr)

BN_EPS = 1e-3           # tf.layers.BatchNormalization default epsilon
BN_C = 1.0 / math.sqrt(1.0 + BN_EPS)   # frozen BN: moving_mean = 0, moving_var = 1
LRELU_ALPHA = 0.3       # tf.keras.layers.LeakyReLU() default (models/customlayers.py:23,36)
KSIZE = 5

AE = 'autoencoder'
VAE = 'variational_autoencoder'
CEVAE = 'context_encoder_variational_autoencoder'
AES = 'autoencoder_spatial'      # encoder -> Dropout -> decoder, no dense bottleneck (reference models/autoencoder_spatial.py)
CAE = 'constrained_autoencoder'  # dense AE whose reconstruction is re-encoded: z_rec = Enc(x_hat) (models/constrained_autoencoder.py)
AAE = 'adversarial_autoencoder'  # dense AE (both bottleneck Dropouts honour the flag, MSE loss) + latent MLP critic (aae_engine.py)
# constrained AE + the latent critic.  The reference scopes its layers differently there (Encoder/{conv2d, dense, dense_1},
# Decoder/conv2d_1: models/constrained_adversarial_autoencoder.py:13-36); the engine keeps ONE canonical naming for the shared
# bottleneck layers (Bottleneck/...), what differs in behaviour - which variables optim_gen updates - is handled by aae_engine.
CAAE = 'constrained_adversarial_autoencoder'
# Gaussian-mixture VAE (models/gaussian_mixture_variational_autoencoder.py): four Dense heads (w_mu, w_log_sigma, z_mu, z_log_sigma),
# z and w reparameterised with exp(0.5 * log-variance), p(z|w,c) heads on w and the mixture latent block (uad_gmvae_latent_*)
GMVAE = 'gaussian_mixture_variational_autoencoder'
# spatial GMVAE (models/gaussian_mixture_variational_autoencoder_spatial.py): the spatial AE's conv stacks (no Dropout; the decoder
# runs on the encoder output itself), 1x1-conv latent heads on the spatial code and the mixture block at every spatial position
GMVAES = 'gaussian_mixture_variational_autoencoder_spatial'
ARCHS = (AE, VAE, CEVAE, AES, CAE, AAE, CAAE, GMVAE, GMVAES)
SPATIAL = (AES, GMVAES)          # graphs without the dense bottleneck
GMVAES_MID = 64                  # filters of p_z_wc/1x1convlayer (model :37)
# Dense widths of the latent critic (models/adversarial_autoencoder.py:44-48, constrained_adversarial_autoencoder.py:52-56)
CRITIC_WIDTHS = {AAE: (50, 50, 1), CAAE: (100, 50, 1)}
AAE_CRITIC = CRITIC_WIDTHS[AAE]


import contextlib
import gc
import os


@contextlib.contextmanager
def graph_capture(graph):
    """torch.cuda.graph with two guards.  (1) The cyclic GC is paused: a collection that runs inside the capture window can
    finalise objects of earlier work (pinned staging buffers, events, other engines) whose destructors issue CUDA calls that
    are illegal while a stream captures - observed as cudaErrorStreamCaptureInvalidated in the middle of a long test
    session.  (2) capture_error_mode='thread_local': only this thread's calls are checked, so a data-loader or logging
    thread of the host application cannot invalidate the capture either."""
    was = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph, capture_error_mode='thread_local'):
            yield
    finally:
        if was:
            gc.enable()


def stack_plan(S, res=8):
    """Channel plan of build_unified_encoder/decoder (reference models/customlayers.py:16-38)."""
    n = int(math.log(S, 2) - math.log(float(res), 2))
    return n, [int(min(128, 32 * 2 ** i)) for i in range(n)], [int(max(32, 128 / 2 ** i)) for i in range(n)]


def _bn(k):
    return 'batch_normalization' if k == 0 else f'batch_normalization_{k}'


def param_specs(arch, S, C=1, zDim=128, res=8, dim_w=1, dim_c=9):
    """TF variable names -> shapes, in creation order."""
    assert arch in ARCHS, arch
    n, enc, dec = stack_plan(S, res)
    sp = OrderedDict()
    cin, bn = C, 0
    for i, co in enumerate(enc):
        sp[f'Encoder/enc_conv2D_{i}/kernel'] = (KSIZE, KSIZE, cin, co)
        sp[f'Encoder/enc_conv2D_{i}/bias'] = (co,)
        sp[f'Encoder/{_bn(bn)}/gamma'] = (co,)
        sp[f'Encoder/{_bn(bn)}/beta'] = (co,)
        bn += 1
        cin = co
    cb = cin // 8
    if arch not in SPATIAL:
        sp['Bottleneck/conv2d/kernel'] = (1, 1, cin, cb)
        sp['Bottleneck/conv2d/bias'] = (cb,)
        sp['Bottleneck/conv2d_1/kernel'] = (1, 1, cb, cin)
        sp['Bottleneck/conv2d_1/bias'] = (cin,)
        flat = res * res * cb
        heads = 1 if arch in (AE, CAE, AAE, CAAE) else (4 if arch == GMVAE else 2)
        for h in range(heads):
            nm = 'dense' if h == 0 else f'dense_{h}'
            width = dim_w if (arch == GMVAE and h < 2) else zDim         # GMVAE: w_mu, w_log_sigma, z_mu, z_log_sigma
            sp[f'Bottleneck/{nm}/kernel'] = (flat, width)
            sp[f'Bottleneck/{nm}/bias'] = (width,)
        sp[f'Bottleneck/dense_{heads}/kernel'] = (zDim, flat)
        sp[f'Bottleneck/dense_{heads}/bias'] = (flat,)
    sp[f'Decoder/{_bn(bn)}/gamma'] = (cin,)
    sp[f'Decoder/{_bn(bn)}/beta'] = (cin,)
    bn += 1
    for i, co in enumerate(dec):
        sp[f'Decoder/dec_Conv2DT_{i}/kernel'] = (KSIZE, KSIZE, co, cin)
        sp[f'Decoder/dec_Conv2DT_{i}/bias'] = (co,)
        sp[f'Decoder/{_bn(bn)}/gamma'] = (co,)
        sp[f'Decoder/{_bn(bn)}/beta'] = (co,)
        bn += 1
        cin = co
    sp['Decoder/dec_Conv2D_final/kernel'] = (1, 1, cin, C)
    sp['Decoder/dec_Conv2D_final/bias'] = (C,)
    if arch == GMVAES:               # the explicitly named 1x1-conv heads (model :15-39) and the trainable 0.1 bias
        ctop = enc[-1]
        for nm, ci, co in (('q_wz_x/w_mu', ctop, dim_w), ('q_wz_x/w_log_sigma', ctop, dim_w), ('q_wz_x/z_mu', ctop, zDim),
                           ('q_wz_x/z_log_sigma', ctop, zDim), ('p_z_wc/1x1convlayer', dim_w, GMVAES_MID),
                           ('p_z_wc/z_wc_mu', GMVAES_MID, zDim * dim_c), ('p_z_wc/z_wc_log_sigma', GMVAES_MID, zDim * dim_c)):
            sp[nm + '/kernel'] = (1, 1, ci, co)
            sp[nm + '/bias'] = (co,)
        sp['Variable'] = (zDim * dim_c,)
    if arch == GMVAE:                # p(z|w,c): the two un-scoped Dense layers on w and the trainable 0.1 bias (model :46-51)
        for nm in ('dense_5', 'dense_6'):
            sp[nm + '/kernel'] = (dim_w, zDim * dim_c)
            sp[nm + '/bias'] = (zDim * dim_c,)
        sp['Variable'] = (zDim * dim_c,)
    if arch in CRITIC_WIDTHS:        # the tf.layers Dense counter runs on: Bottleneck/{dense, dense_1}, Discriminator/dense_{2,3,4}
        k = zDim
        for j, width in enumerate(CRITIC_WIDTHS[arch]):
            sp[f'Discriminator/dense_{2 + j}/kernel'] = (k, width)
            sp[f'Discriminator/dense_{2 + j}/bias'] = (width,)
            k = width
    return sp


def glorot_init(specs, seed=1):
    """Glorot-uniform kernels, zero biases, gamma=1, beta=0 - the TF defaults the reference relies on."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shape in specs.items():
        if name.endswith('/kernel'):
            if len(shape) == 4:
                kh, kw, a, b = shape
                fan_in, fan_out = kh * kw * a, kh * kw * b
            else:
                fan_in, fan_out = shape
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            out[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        elif name.endswith('/gamma'):
            out[name] = np.ones(shape, np.float32)
        elif name == 'Variable':                         # GMVAE: tf.constant(0.1) bias of z_wc_log_sigma_inv
            out[name] = np.full(shape, 0.1, np.float32)
        else:
            out[name] = np.zeros(shape, np.float32)
    return out


class FlatParams:
    """Flat fp32 parameter / gradient / Adam-moment buffers with named views (TF variable names)."""
    ALIGN = 64   # floats (256 B) per slot boundary: float4 / TMA friendly

    def __init__(self, specs, device):
        self.specs = OrderedDict(specs)
        self.offsets = OrderedDict()
        off = 0
        for name, shape in self.specs.items():
            self.offsets[name] = off
            off += -(-int(np.prod(shape)) // self.ALIGN) * self.ALIGN
        self.numel = off
        self.n_params = sum(int(np.prod(s)) for s in self.specs.values())
        self.params = torch.zeros(off, dtype=torch.float32, device=device)
        self.grads = torch.zeros(off, dtype=torch.float32, device=device)
        self.m = torch.zeros(off, dtype=torch.float32, device=device)
        self.v = torch.zeros(off, dtype=torch.float32, device=device)

    def _view(self, buf, name):
        o = self.offsets[name]
        n = int(np.prod(self.specs[name]))
        return buf[o:o + n]

    def p(self, name):
        return self._view(self.params, name)

    def g(self, name):
        return self._view(self.grads, name)

    def load(self, values, buf=None):
        buf = self.params if buf is None else buf
        host = torch.zeros(self.numel, dtype=torch.float32)
        for name in self.specs:
            v = np.asarray(values[name], np.float32).reshape(-1)
            assert v.size == int(np.prod(self.specs[name])), name
            host[self.offsets[name]:self.offsets[name] + v.size] = torch.from_numpy(v)
        buf.copy_(host)

    def to_numpy(self, buf=None):
        buf = self.params if buf is None else buf
        host = buf.detach().cpu().numpy()
        return OrderedDict((n, host[self.offsets[n]:self.offsets[n] + int(np.prod(s))].reshape(s).copy())
                           for n, s in self.specs.items())

    def subset_ranges(self, prefix):
        """Contiguous [lo, hi) range of slots whose names start with ``prefix`` (scope-contiguous layout)."""
        names = [n for n in self.specs if n.startswith(prefix)]
        lo = self.offsets[names[0]]
        last = names[-1]
        hi = self.offsets[last] + -(-int(np.prod(self.specs[last])) // self.ALIGN) * self.ALIGN
        return lo, hi


class _Branch:
    """Activation storage of one pass through the shared layers (ceVAE runs two)."""
    pass


class ConvAutoencoderEngine:
    """Forward / backward / Adam of the AE, VAE and ceVAE graphs on one GPU.

    Restates (as fused device calls) reference models/autoencoder.py:9-40, variational_autoencoder.py:9-47,
    context_encoder_variational_autoencoder.py:9-59 and the loss graphs of trainers/AE.py:28-29, VAE.py:36-42,
    ceVAE.py:38-51.
    """

    def __init__(self, arch, S, C=1, zDim=128, res=8, batch=8, device='cuda:0', math_mode=abi.MATH_FP32_SIMT, seed=1,
                 rng_seed=0x5eed, share_params=None, keep_preact=False, dim_w=1, dim_c=9, c_lambda=1.0):
        if arch not in ARCHS:
            raise ValueError(f'unsupported architecture {arch!r}; supported: {ARCHS}')
        if C != 1:
            raise NotImplementedError('the fused final-1x1+L1 path supports numChannels == 1 (all reference datasets are '
                                      'single-channel: dataloaders/BRAINWEB.py:351-352)')
        abi.lib()   # fail loudly if the CUDA library is missing
        self.arch, self.S, self.C, self.zDim, self.res, self.B = arch, S, C, zDim, res, batch
        self.device = torch.device(device)
        self.math_mode = math_mode
        self.n, self.enc_ch, self.dec_ch = stack_plan(S, res)
        self.cb = self.enc_ch[-1] // 8
        self.flat = res * res * self.cb
        self.dim_w, self.dim_c, self.c_lambda = int(dim_w), int(dim_c), float(c_lambda)      # GMVAE only
        self.specs = param_specs(arch, S, C, zDim, res, dim_w, dim_c)
        if share_params is not None:      # e.g. an evaluation engine with another batch size on the same weights
            self.fp = share_params
        else:
            self.fp = FlatParams(self.specs, self.device)
            self.fp.load(glorot_init(self.specs, seed))
        self.t = 0
        self.probes = None
        self.graph = None
        self.rng_seed = rng_seed
        # keep_preact=False (default): conv blocks never write their pre-BN tensor z; the backward recovers what it needs
        # from the activation a (UAD_ACT_FROM_OUTPUT, LeakyReLU is invertible) - half the forward HBM write volume and
        # ~1 GB less activation memory at 256x256, B=64.  keep_preact=True restores the z-based backward.
        self.keep_preact = bool(keep_preact)
        self._alloc()

    # ------------------------------------------------------------------ buffers
    def _new(self, *shape):
        return torch.empty(*shape, dtype=torch.float32, device=self.device)

    def _alloc_branch(self, encoder_only=False, x=None):
        B, S = self.B, self.S
        br = _Branch()
        br.x = self._new(B, S, S, 1) if x is None else x
        br.target = None                                 # reconstruction target when it is not the input (context encoder, set_target)
        br.enc_z, br.enc_a = [], []
        s = S
        for co in self.enc_ch:
            s //= 2
            br.enc_z.append(self._new(B, s, s, co) if self.keep_preact else None)
            br.enc_a.append(self._new(B, s, s, co))
        r = self.res
        br.zb = self._new(B, r, r, self.cb)              # bottleneck 1x1 output == flatten input [B, flat]
        br.mu = self._new(B, self.zDim)                  # AE: z
        br.ls = self._new(B, self.zDim)
        br.sigma = self._new(B, self.zDim)
        br.zv = self._new(B, self.zDim)
        br.kl = self._new(B)
        br.masks = {k: None for k in ('mu', 'ls', 'dec', 'sp')}
        br.mask_bufs = {'mu': self._new(B, self.zDim)}
        if encoder_only:                                 # the re-encoding pass of the constrained AE: x_hat -> z_rec
            return br
        br.d = self._new(B, self.flat)                   # dec_dense output (post-dropout)
        br.zr = self._new(B, r, r, self.enc_ch[-1])      # conv2d_1 output (pre decoder-BN)
        br.ar = self._new(B, r, r, self.enc_ch[-1])      # after decoder BN + ReLU
        br.dec_z, br.dec_a = [], []
        s = r
        for co in self.dec_ch:
            s *= 2
            br.dec_z.append(self._new(B, s, s, co) if self.keep_preact else None)
            br.dec_a.append(self._new(B, s, s, co))
        br.xhat = self._new(B, S, S, 1)
        br.l1 = self._new(B, S, S, 1)
        br.rec = self._new(B)
        br.eps = self._new(B, self.zDim)
        br.mask_bufs.update(ls=self._new(B, self.zDim), dec=self._new(B, self.flat))
        if self.arch == AES:
            br.mask_bufs['sp'] = self._new(B, r, r, self.enc_ch[-1])      # Dropout on the spatial code z [B,res,res,C]
        if self.arch == GMVAES:
            R, dw, dz, n, ct = B * r * r, self.dim_w, self.zDim, self.zDim * self.dim_c, self.enc_ch[-1]
            for k, width in (('w_mu', dw), ('w_ls', dw), ('w_lsh', dw), ('w_sigma', dw), ('w_s', dw), ('eps_w', dw), ('dws', dw), ('dwmu', dw),
                             ('dwlsh', dw), ('mu', dz), ('ls', dz), ('z_lsh', dz), ('sigma', dz), ('zv', dz), ('eps', dz), ('gzmu', dz),
                             ('gzls', dz), ('gzs', dz), ('dmu', dz), ('dlsh', dz), ('mid_pre', GMVAES_MID), ('mid', GMVAES_MID),
                             ('dmid', GMVAES_MID), ('dmid2', GMVAES_MID), ('Mz', n), ('S0', n), ('Sz', n), ('dM', n), ('dS', n), ('dh', ct)):
                setattr(br, k, self._new(R, width))
            br.kl_w, br.kl_unused, br.con, br.closs = (self._new(R) for _ in range(4))
            br.pc = self._new(R, self.dim_c)
            br.ones_n = torch.ones(n, dtype=torch.float32, device=self.device)
        if self.arch == GMVAE:
            dw, n = self.dim_w, self.zDim * self.dim_c
            for k in ('w_mu', 'w_ls', 'w_lsh', 'w_sigma', 'w_s', 'eps_w', 'dws', 'dws2', 'dwmu', 'dwlsh'):
                setattr(br, k, self._new(B, dw))
            for k in ('z_lsh', 'gzmu', 'gzls', 'gzs', 'dlsh'):
                setattr(br, k, self._new(B, self.zDim))
            for k in ('Mz', 'S0', 'Sz', 'dM', 'dS'):
                setattr(br, k, self._new(B, n))
            br.kl_w, br.kl_unused, br.con, br.closs = (self._new(B) for _ in range(4))
            br.pc = self._new(B, self.dim_c)
            br.ones_n = torch.ones(n, dtype=torch.float32, device=self.device)
            br.masks.update(wmu=None, wls=None)
            br.mask_bufs.update(wmu=self._new(B, dw), wls=self._new(B, dw))
        return br

    def _alloc(self):
        B, S = self.B, self.S
        self.br = [self._alloc_branch()]
        if self.arch == CEVAE:
            self.br.append(self._alloc_branch())
        if self.arch in (CAE, CAAE):
            self.br.append(self._alloc_branch(encoder_only=True, x=self.br[0].xhat))
            self.gxhat = self._new(B, S, S, 1)           # d loss / d x_hat: MSE term + the re-encoding pass
            self.dzrec = self._new(B, self.zDim)
            self.rho = 1.0                               # trainers/ConstrainedAE.py:16
        if self.arch == AAE:
            self.gxhat = self._new(B, S, S, 1)           # d loss / d x_hat of the MSE loss (trainers/AAE.py:55-57)
        big = B * S * S * max(32, self.dec_ch[-1])
        self.gbuf = [self._new(big), self._new(big)]
        self.gx = self._new(B, S, S, 1)                  # d loss / d x (ceVAE anomaly)
        self.anomaly = self._new(B, S, S, 1)
        self.small = {k: self._new(B, n) for k, n in (('dd', self.flat), ('dzv', self.zDim), ('dmu', self.zDim),
                                                      ('dls', self.zDim), ('dflat', self.flat), ('dflat2', self.flat))}
        self.scalars = torch.zeros(8, dtype=torch.float32, device=self.device)
        self.lr_dev = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.rng_ctr = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        # workspace: max over every op this engine issues
        L = abi.lib()
        need = 1 << 20
        s, cin = S, 1
        for co in self.enc_ch:
            for op in (OP_CONV_FWD, OP_CONV_DGRAD, OP_CONV_WGRAD):
                need = max(need, L.uad_conv_workspace_bytes(op, B, s, s, cin, co, KSIZE, self.math_mode))
            need = max(need, L.uad_rowreduce_workspace_bytes(B * (s // 2) ** 2, co))
            s //= 2
            cin = co
        for co in self.dec_ch:
            for op in (OP_CONVT_FWD, OP_CONVT_DGRAD, OP_CONVT_WGRAD):
                need = max(need, L.uad_conv_workspace_bytes(op, B, s, s, cin, co, KSIZE, self.math_mode))
            need = max(need, L.uad_rowreduce_workspace_bytes(B * (2 * s) ** 2, co))
            s *= 2
            cin = co
        r2 = self.res * self.res
        for (M, K, N) in ((B * r2, self.enc_ch[-1], self.cb), (B * r2, self.cb, self.enc_ch[-1]), (B, self.flat, self.zDim),
                          (B, self.zDim, self.flat)):
            need = max(need, L.uad_dense_workspace_bytes(M, K, N))
        need = max(need, L.uad_tv_restore_workspace_bytes(B, S, S))
        self.ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        self.ws_bytes = need
        # filter gradients run on a side stream beside the input-gradient chain (backward()): their own workspace
        self.side_wgrad = os.environ.get('UAD_SIDE_WGRAD', '1') != '0' and torch.device(self.device).type == 'cuda'
        self.ws2 = torch.empty(need, dtype=torch.uint8, device=self.device) if self.side_wgrad else None
        self._side = None
        self._side_last = None
        # data parallel, opt-in (UAD_DP_BUCKETS=1): the decoder's gradient bucket is all-reduced while the encoder's backward pass
        # runs (train_step).  Bit-identical to the single all-reduce; measured on 2 x B200 it does not pay (the 8.8 MB collective
        # takes ~25 us there: c2 4.215 vs 4.211 ms, c4 1.743 vs 1.735 ms per step), so the default stays one all-reduce behind the replay
        self.dp_buckets = os.environ.get('UAD_DP_BUCKETS', '0') != '0' and torch.device(self.device).type == 'cuda'
        self._bucket_async = None
        self._bucket_work = None
        self._bucket_lo = None
        self.peer = None                 # dist.PeerOptimizer (enable_peer_optimizer): fused reduce-scatter + Adam + all-gather

    # ------------------------------------------------------------------ helpers
    def _op(self, label, fname, *args):
        """Issue one ABI call; with ``self.probes`` set (eager mode only) bracket it with CUDA events for per-kernel timing."""
        if self.probes is None:
            return call(fname, *args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call(fname, *args)
        e1.record()
        self.probes.setdefault(f'{label}:{fname[4:]}', []).append((e0, e1))

    def _op_side(self, label, fname, *args):
        """Issue a filter-gradient call (its last three arguments are ws, ws_bytes, stream) on the side stream with the side
        workspace: it reads what the main stream has produced so far (fork event) and runs beside the input-gradient kernel of
        the same layer, filling the SMs that kernel's tail leaves idle.  Returns the event that marks its completion; the
        caller makes the main stream wait for it before a buffer the call reads is overwritten (``_wait_side``).  With
        per-kernel probes on, or UAD_SIDE_WGRAD=0, the call runs in line."""
        if not self.side_wgrad or self.probes is not None:
            self._op(label, fname, *args)
            return None
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        fork = torch.cuda.Event()
        fork.record(main)
        self._side.wait_event(fork)
        call(fname, *args[:-3], self.ws2.data_ptr(), self.ws_bytes, self._side.cuda_stream)
        done = torch.cuda.Event()
        done.record(self._side)
        self._side_last = done
        return done

    def _wait_side(self, ev):
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)

    def _launch_decoder_bucket(self):
        """Called by backward() once every Decoder/ gradient of the step has been issued (the decoder's filter gradients on
        the side stream, its BN / bias gradients on the main stream): start the sum-all-reduce of that contiguous slice of the
        flat gradient buffer behind both, without blocking either - it travels over NVLink while the bottleneck and encoder
        backward run.  train_step reduces the rest and waits for this bucket before Adam."""
        if self._bucket_async is None or self.probes is not None:
            return
        lo, hi = self.fp.subset_ranges('Decoder/')
        if hi != self.fp.numel:
            return
        main = torch.cuda.current_stream(self.device)
        if self._side is not None:
            fork = torch.cuda.Event()
            fork.record(main)
            self._side.wait_event(fork)
            with torch.cuda.stream(self._side):
                self._bucket_work = self._bucket_async(self.fp.grads[lo:hi])
        else:
            self._bucket_work = self._bucket_async(self.fp.grads[lo:hi])
        self._bucket_lo = lo if self._bucket_work is not None else None

    def _finish_allreduce(self, allreduce):
        """The gradient all-reduce of a step: everything, or - when the decoder bucket is already in flight - the rest."""
        if self._bucket_work is not None:
            allreduce(self.fp.grads[:self._bucket_lo])
            self._bucket_work.wait()
            self._bucket_work = None
        else:
            allreduce(self.fp.grads)

    def _join_side(self):
        """Main stream waits for everything issued on the side stream (end of a backward pass; required before a capture ends)."""
        if self._side_last is not None:
            torch.cuda.current_stream(self.device).wait_event(self._side_last)
            self._side_last = None

    def probe_times_ms(self):
        torch.cuda.synchronize(self.device)
        return {k: [a.elapsed_time(b) for a, b in v] for k, v in (self.probes or {}).items()}

    def _st(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _wsargs(self):
        return self.ws.data_ptr(), self.ws_bytes

    def set_inputs(self, x, x_ce=None):
        """Stage a batch (numpy NHWC or a device tensor) into the static input buffers (the feed_dict of sess.run)."""
        for br, v in zip(self.br, (x, x_ce)):
            if v is None:
                continue
            if isinstance(v, np.ndarray):
                v = torch.from_numpy(np.ascontiguousarray(v, np.float32))
            br.x.copy_(v.reshape(br.x.shape), non_blocking=True)

    def set_target(self, x):
        """Reconstruction target of branch 0 when it differs from the input: the context-encoder trainer feeds the masked
        batch and scores against the plain one (reference trainers/CE.py:21,34,87-88).  None: back to target == input."""
        br = self.br[0]
        if x is None:
            if br.target is not None:
                br.target, self.graph = None, None       # a captured step holds the old target pointer
            return
        if br.target is None:
            br.target, self.graph = self._new(*br.x.shape), None
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, np.float32))
        br.target.copy_(x.reshape(br.target.shape), non_blocking=True)

    def set_noise(self, eps=None, masks=None, masks_ce=None):
        """Parity mode: caller-supplied eps / dropout masks ({0,1} arrays keyed 'mu','ls','dec'; AE uses 'mu' for z)."""
        if eps is not None:
            self.br[0].eps.copy_(torch.as_tensor(eps, dtype=torch.float32).reshape(self.br[0].eps.shape))
        for br, ms in zip(self.br, (masks, masks_ce)):
            if ms is None:
                continue
            for k, m in ms.items():
                br.mask_bufs[k].copy_(torch.as_tensor(m, dtype=torch.float32).reshape(br.mask_bufs[k].shape))
                br.masks[k] = br.mask_bufs[k]

    def draw_noise(self, dropout, rate):
        """Perf mode: fresh eps / masks from the in-library Philox streams (tf.random_normal / Dropout are always live)."""
        st = self._st()
        ctr = self.rng_ctr.data_ptr()
        nb = 0
        if self.arch == GMVAES:          # no Dropout in this graph; eps_z, eps_w per spatial position
            br = self.br[0]
            call('uad_randn', ptr(br.eps), br.eps.numel(), self.rng_seed, 0 << 40, ctr, st)
            call('uad_randn', ptr(br.eps_w), br.eps_w.numel(), self.rng_seed, 1 << 40, ctr, st)
            br.masks['sp'] = None
            call('uad_counter_add', ctr, 1 << 20, st)
            return
        if self.arch == AES:
            br = self.br[0]
            if dropout and rate > 0:
                call('uad_dropout_mask', ptr(br.mask_bufs['sp']), br.mask_bufs['sp'].numel(), float(rate), self.rng_seed, 1 << 40, ctr, st)
                br.masks['sp'] = br.mask_bufs['sp']
            else:
                br.masks['sp'] = None
            call('uad_counter_add', ctr, 1 << 20, st)
            return
        if self.arch in (CAE, AAE, CAAE):                # Dropout calls: z, dec_dense(z) [, z_rec] (each its own draw)
            on = bool(dropout) and rate > 0
            # constrained_adversarial_autoencoder.py:35,48 call Dropout on dec_dense(z_) and on z_rec WITHOUT the flag: identity
            sites = ((self.br[0], 'mu'),) + (((self.br[0], 'dec'),) if self.arch != CAAE else ()) + \
                    (((self.br[1], 'mu'),) if self.arch == CAE else ())
            for sid, (br, k) in enumerate(sites):
                if on:
                    call('uad_dropout_mask', ptr(br.mask_bufs[k]), br.mask_bufs[k].numel(), float(rate), self.rng_seed,
                         (sid + 1) << 40, ctr, st)
                br.masks[k] = br.mask_bufs[k] if on else None
            call('uad_counter_add', ctr, 1 << 20, st)
            return
        if self.arch == GMVAE:           # eps_z, eps_w; Dropout with the flag on w_mu, w_log_sigma, z_mu, dec_dense(z) (z_log_sigma: none)
            br = self.br[0]
            on = bool(dropout) and rate > 0
            call('uad_randn', ptr(br.eps), br.eps.numel(), self.rng_seed, 0 << 40, ctr, st)
            call('uad_randn', ptr(br.eps_w), br.eps_w.numel(), self.rng_seed, 1 << 40, ctr, st)
            for sid, k in enumerate(('wmu', 'wls', 'mu', 'dec')):
                if on:
                    call('uad_dropout_mask', ptr(br.mask_bufs[k]), br.mask_bufs[k].numel(), float(rate), self.rng_seed, (sid + 2) << 40,
                         ctr, st)
                br.masks[k] = br.mask_bufs[k] if on else None
            br.masks['ls'] = None
            call('uad_counter_add', ctr, 1 << 20, st)
            return
        for bi, br in enumerate(self.br):
            if self.arch != AE and bi == 0:
                call('uad_randn', ptr(br.eps), br.eps.numel(), self.rng_seed, nb << 40, ctr, st)
            nb += 1
            for k in ('mu', 'ls', 'dec'):
                if dropout and rate > 0 and not (self.arch == AE and k != 'mu') and not (bi == 1 and k == 'ls'):
                    call('uad_dropout_mask', ptr(br.mask_bufs[k]), br.mask_bufs[k].numel(), float(rate), self.rng_seed,
                         nb << 40, ctr, st)
                    br.masks[k] = br.mask_bufs[k]
                else:
                    br.masks[k] = None
                nb += 1
        call('uad_counter_add', ctr, 1 << 20, st)

    # ------------------------------------------------------------------ forward
    def forward(self, training=True, dropout_rate=0.0, branches=None, need_l1=True):
        """One pass x -> x_hat (+ per-sample rec / kl).  ``training`` keeps the pre-BN tensors for backward."""
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self._wsargs()
        keep = 1.0 / (1.0 - dropout_rate) if dropout_rate > 0 else 1.0
        B = self.B
        for bi, br in enumerate(self.br if branches is None else [self.br[i] for i in branches]):
            is_ce = br is not self.br[0]
            h, s, cin = br.x, self.S, 1
            for i, co in enumerate(self.enc_ch):
                pre = f'Encoder/enc_conv2D_{i}'
                bnn = f'Encoder/{_bn(i)}'
                self._op(pre.split('/')[-1], 'uad_conv2d_fwd', ptr(h), ptr(fp.p(pre + '/kernel')), ptr(fp.p(pre + '/bias')),
                     ptr(fp.p(bnn + '/gamma')), ptr(fp.p(bnn + '/beta')), ptr(br.enc_z[i]) if training and self.keep_preact else None,
                     ptr(br.enc_a[i]), B, s, s, cin, co, KSIZE, ACT_LEAKY, LRELU_ALPHA, BN_C, mm, ws, wsb, st)
                h, s, cin = br.enc_a[i], s // 2, co
            r2 = self.res * self.res
            m = br.masks
            if self.arch in SPATIAL:
                # autoencoder_spatial.py:16-23: z = Dropout(encoder(x)); decoder = BN -> ReLU -> ...
                dbn = f'Decoder/{_bn(self.n)}'
                self._op('bneck01', 'uad_mask_bn_act_fwd', ptr(h), ptr(m['sp']), keep, ptr(fp.p(dbn + '/gamma')), ptr(fp.p(dbn + '/beta')),
                         BN_C, ACT_RELU, 0.0, ptr(br.zr), ptr(br.ar), B * r2, cin, st)
                if self.arch == GMVAES:
                    self._gmvaes_latent_fwd(br, h)
            else:
                self._op('bneck01', 'uad_dense_fwd', ptr(h), ptr(fp.p('Bottleneck/conv2d/kernel')), ptr(fp.p('Bottleneck/conv2d/bias')), None,
                   1.0, None, None, ptr(br.zb), None, B * r2, cin, self.cb, ACT_NONE, 0.0, 1.0, ws, wsb, st)
            if self.arch in SPATIAL:
                pass
            elif self.arch in (AE, CAE, AAE, CAAE):
                # autoencoder.py:29: dropout on z honours the flag; :30 dropout on dec_dense(z) has no flag -> identity.
                # constrained_autoencoder.py:29-30, adversarial_autoencoder.py:30-31: BOTH dropout calls honour the flag.
                self._op('bneck02', 'uad_dense_fwd', ptr(br.zb), ptr(fp.p('Bottleneck/dense/kernel')), ptr(fp.p('Bottleneck/dense/bias')),
                     ptr(m['mu']), keep, None, None, ptr(br.mu), None, B, self.flat, self.zDim, ACT_NONE, 0.0, 1.0, ws, wsb, st)
                zsrc, dd_name, dec_mask = br.mu, 'Bottleneck/dense_1', (m['dec'] if self.arch in (CAE, AAE, CAAE) else None)
                if is_ce:                                # constrained AE, re-encoding pass: z_rec is all that is needed
                    continue
            elif self.arch == GMVAE:
                zsrc, dd_name, dec_mask = self._gmvae_bottleneck_fwd(br, keep)
            else:
                self._op('bneck03', 'uad_dense_fwd', ptr(br.zb), ptr(fp.p('Bottleneck/dense/kernel')), ptr(fp.p('Bottleneck/dense/bias')),
                     ptr(m['mu']), keep, None, None, ptr(br.mu), None, B, self.flat, self.zDim, ACT_NONE, 0.0, 1.0, ws, wsb, st)
                if not is_ce:
                    self._op('bneck04', 'uad_dense_fwd', ptr(br.zb), ptr(fp.p('Bottleneck/dense_1/kernel')),
                         ptr(fp.p('Bottleneck/dense_1/bias')), ptr(m['ls']), keep, None, None, ptr(br.ls), None, B,
                         self.flat, self.zDim, ACT_NONE, 0.0, 1.0, ws, wsb, st)
                    self._op('bneck05', 'uad_reparam_kl_fwd', ptr(br.mu), ptr(br.ls), ptr(br.eps), ptr(br.sigma), ptr(br.zv), ptr(br.kl),
                         B, self.zDim, st)
                    zsrc = br.zv
                else:
                    zsrc = br.mu      # ce branch decodes z_mu_ce without sampling (ceVAE model :37,43)
                dd_name, dec_mask = 'Bottleneck/dense_2', m['dec']
            if self.arch not in SPATIAL:
                self._op('bneck06', 'uad_dense_fwd', ptr(zsrc), ptr(fp.p(dd_name + '/kernel')), ptr(fp.p(dd_name + '/bias')), ptr(dec_mask),
                   keep, None, None, ptr(br.d), None, B, self.zDim, self.flat, ACT_NONE, 0.0, 1.0, ws, wsb, st)
                dbn = f'Decoder/{_bn(self.n)}'
                self._op('bneck07', 'uad_dense_fwd', ptr(br.d), ptr(fp.p('Bottleneck/conv2d_1/kernel')), ptr(fp.p('Bottleneck/conv2d_1/bias')),
                   None, 1.0, ptr(fp.p(dbn + '/gamma')), ptr(fp.p(dbn + '/beta')), ptr(br.zr) if training else None,
                   ptr(br.ar), B * r2, self.cb, cin, ACT_RELU, 0.0, BN_C, ws, wsb, st)
            h, s = br.ar, self.res
            fused_head = False
            for i, co in enumerate(self.dec_ch):
                pre = f'Decoder/dec_Conv2DT_{i}'
                bnn = f'Decoder/{_bn(self.n + 1 + i)}'
                keep_z = training and self.keep_preact
                # the last block + dec_Conv2D_final (customlayers.py:34-37) as ONE launch where the library offers it: x_hat comes out
                # of the block's epilogue and the 537 MB re-read of the block output by the 1x1 conv disappears
                if i == len(self.dec_ch) - 1 and not keep_z and self.C == 1 and \
                        abi.lib().uad_convT2d_fwd_head_supported(B, s, s, cin, co, KSIZE, mm):
                    self._op(pre.split('/')[-1], 'uad_convT2d_fwd_head', ptr(h), ptr(fp.p(pre + '/kernel')), ptr(fp.p(pre + '/bias')),
                         ptr(fp.p(bnn + '/gamma')), ptr(fp.p(bnn + '/beta')), ptr(br.dec_a[i]),
                         ptr(fp.p('Decoder/dec_Conv2D_final/kernel')), ptr(fp.p('Decoder/dec_Conv2D_final/bias')), ptr(br.xhat),
                         B, s, s, cin, co, KSIZE, ACT_LEAKY, LRELU_ALPHA, BN_C, mm, ws, wsb, st)
                    fused_head = True
                else:
                    self._op(pre.split('/')[-1], 'uad_convT2d_fwd', ptr(h), ptr(fp.p(pre + '/kernel')), ptr(fp.p(pre + '/bias')),
                         ptr(fp.p(bnn + '/gamma')), ptr(fp.p(bnn + '/beta')), ptr(br.dec_z[i]) if keep_z else None,
                         ptr(br.dec_a[i]), B, s, s, cin, co, KSIZE, ACT_LEAKY, LRELU_ALPHA, BN_C, mm, ws, wsb, st)
                h, s, cin = br.dec_a[i], s * 2, co
            if fused_head:
                self._op('dec_Conv2D_final', 'uad_l1_map', ptr(br.x if br.target is None else br.target), ptr(br.xhat),
                     ptr(br.l1) if need_l1 else None, ptr(br.rec), B, self.S * self.S, st)
            else:
                self._op('dec_Conv2D_final', 'uad_final1x1_l1_fwd', ptr(h), ptr(fp.p('Decoder/dec_Conv2D_final/kernel')),
                     ptr(fp.p('Decoder/dec_Conv2D_final/bias')), ptr(br.x if br.target is None else br.target), ptr(br.xhat),
                     ptr(br.l1) if need_l1 else None,
                     ptr(br.rec), B, self.S * self.S, cin, ws, wsb, st)
        # loss scalars (trainers/VAE.py:40-42; ceVAE.py:44-49): out = [mean rec, mean kl, mean(rec+kl)] per branch
        b0 = self.br[0]
        self._op('bneck08', 'uad_loss_scalars', ptr(b0.rec), ptr(b0.kl) if self.arch not in (AE, AES, CAE, AAE, CAAE, GMVAE, GMVAES) else None, ptr(self.scalars), B, st)
        if self.arch in (CAE, CAAE) and (branches is None or 1 in branches):
            # trainers/ConstrainedAE.py:37-43: L2 = mean_hwc (x - x_hat)^2, Rec_z = mean_j (z - z_rec)^2 (per sample);
            # loss = mean_b(L2 + rho * Rec_z).  The same calls leave d loss/d x_hat and d loss/d z_rec (d/dz = -d/dz_rec).
            nx, nz = b0.x.numel(), B * self.zDim
            self._op('bneck09', 'uad_mse', ptr(b0.xhat), ptr(b0.x), nx, 2.0 / nx, ptr(self.gxhat), 1.0 / nx, self.scalars[4:].data_ptr(),
                     ws, wsb, st)
            self._op('bneck09', 'uad_mse', ptr(self.br[1].mu), ptr(b0.mu), nz, 2.0 * self.rho / nz, ptr(self.dzrec), 1.0 / nz,
                     self.scalars[5:].data_ptr(), ws, wsb, st)
        if self.arch == CEVAE and (branches is None or 1 in branches):
            self._op('bneck09', 'uad_loss_scalars', ptr(self.br[1].rec), None, ptr(self.scalars[3:]), B, st)
        if self.arch == AAE:         # trainers/AAE.py:55-57: loss = mean_b mean_hwc (x - x_hat)^2; leaves d loss / d x_hat in gxhat
            nx = b0.x.numel()
            self._op('bneck09', 'uad_mse', ptr(b0.xhat), ptr(b0.x), nx, 2.0 / nx, ptr(self.gxhat), 1.0 / nx, self.scalars[4:].data_ptr(),
                     ws, wsb, st)

    # ------------------------------------------------------------------ GMVAE bottleneck (models/gaussian_mixture_variational_autoencoder.py:21-71)
    def _gmvae_bottleneck_fwd(self, br, keep):
        """zb -> w_mu, w_log_sigma, z_mu, z_log_sigma -> w, z (std = exp(0.5 * log-variance): uad_reparam_kl_fwd on the halved
        log-variance, whose KL output for w IS w_prior_loss) -> p(z|w,c) heads -> responsibilities, conditional-prior and
        cluster-prior losses per sample.  Returns what the shared decoder entry needs (source, Dense name, mask)."""
        fp, st = self.fp, self._st()
        ws, wsb = self._wsargs()
        B, dz, dw, n, m = self.B, self.zDim, self.dim_w, self.zDim * self.dim_c, br.masks
        for name, out, mask, width in (('dense', br.w_mu, m['wmu'], dw), ('dense_1', br.w_ls, m['wls'], dw), ('dense_2', br.mu, m['mu'], dz),
                                       ('dense_3', br.ls, None, dz)):
            self._op('bneck03', 'uad_dense_fwd', ptr(br.zb), ptr(fp.p(f'Bottleneck/{name}/kernel')), ptr(fp.p(f'Bottleneck/{name}/bias')),
                     ptr(mask), keep, None, None, ptr(out), None, B, self.flat, width, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        self._op('bneck05', 'uad_axpby', 0.5, ptr(br.w_ls), 0.0, ptr(br.w_lsh), B * dw, st)
        self._op('bneck05', 'uad_axpby', 0.5, ptr(br.ls), 0.0, ptr(br.z_lsh), B * dz, st)
        self._op('bneck05', 'uad_reparam_kl_fwd', ptr(br.w_mu), ptr(br.w_lsh), ptr(br.eps_w), ptr(br.w_sigma), ptr(br.w_s), ptr(br.kl_w), B, dw, st)
        self._op('bneck05', 'uad_reparam_kl_fwd', ptr(br.mu), ptr(br.z_lsh), ptr(br.eps), ptr(br.sigma), ptr(br.zv), ptr(br.kl_unused), B, dz, st)
        self._op('bneck05', 'uad_dense_fwd', ptr(br.w_s), ptr(fp.p('dense_5/kernel')), ptr(fp.p('dense_5/bias')), None, 1.0, None, None,
                 ptr(br.Mz), None, B, dw, n, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        # z_wc_log_sigma_inv = Dense(w) + Variable: the extra trainable bias rides the affine slot (gamma = 1, beta = Variable)
        self._op('bneck05', 'uad_dense_fwd', ptr(br.w_s), ptr(fp.p('dense_6/kernel')), ptr(fp.p('dense_6/bias')), None, 1.0, ptr(br.ones_n),
                 ptr(fp.p('Variable')), ptr(br.S0), ptr(br.Sz), B, dw, n, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        self._op('bneck05', 'uad_gmvae_latent_fwd', ptr(br.mu), ptr(br.ls), ptr(br.zv), ptr(br.Mz), ptr(br.Sz), ptr(br.pc), ptr(br.con),
                 ptr(br.closs), B, dz, self.dim_c, self.c_lambda, st)
        for buf, slot in ((br.con, 5), (br.kl_w, 6), (br.closs, 7)):     # batch means: conditional_prior / w_prior / c_prior loss
            self._op('bneck08', 'uad_sum_scaled', ptr(buf), B, 1.0 / B, self.scalars[slot:].data_ptr(), ws, wsb, st)
        return br.zv, 'Bottleneck/dense_4', m['dec']

    def _gmvae_bottleneck_bwd(self, br, params, scale, acc):
        """From sm['dd'] (gradient w.r.t. dec_dense's post-dropout output) to sm['dflat'] (w.r.t. the flattened 1x1-conv output):
        the reconstruction path through z plus scale * d(con + w_loss + c_loss)_b.  params=False forms no weight gradient
        (restoration, tf.gradients(..., x)); scale = 1/B for the batch-mean training loss, 1 for the per-sample sums of GMVAE.py:89-90."""
        fp, st = self.fp, self._st()
        ws, wsb = self._wsargs()
        B, dz, dw, n, m, sm, keep = self.B, self.zDim, self.dim_w, self.zDim * self.dim_c, br.masks, self.small, self._keep
        G = (lambda name: ptr(fp.g(name))) if params else (lambda name: None)
        self._op('bneck13', 'uad_dense_bwd', ptr(br.zv), ptr(fp.p('Bottleneck/dense_4/kernel')), ptr(sm['dd']), ptr(m['dec']), keep,
                 ptr(sm['dzv']), G('Bottleneck/dense_4/kernel'), G('Bottleneck/dense_4/bias'), B, dz, self.flat, acc, ws, wsb, st)
        self._op('bneck14', 'uad_gmvae_latent_bwd', ptr(br.mu), ptr(br.ls), ptr(br.zv), ptr(br.Mz), ptr(br.Sz), float(scale), ptr(br.gzmu),
                 ptr(br.gzls), ptr(br.gzs), ptr(br.dM), ptr(br.dS), B, dz, self.dim_c, self.c_lambda, st)
        self._op('bneck14', 'uad_axpby', 1.0, ptr(br.gzs), 1.0, ptr(sm['dzv']), B * dz, st)                  # d/dz_sampled, both paths
        self._op('bneck14', 'uad_reparam_kl_bwd', ptr(br.mu), ptr(br.z_lsh), ptr(br.eps), ptr(sm['dzv']), 0.0, ptr(sm['dmu']), ptr(br.dlsh),
                 B, dz, st)
        self._op('bneck14', 'uad_axpby', 1.0, ptr(br.gzmu), 1.0, ptr(sm['dmu']), B * dz, st)                 # d/dz_mu
        self._op('bneck14', 'uad_axpby', 0.5, ptr(br.dlsh), 1.0, ptr(br.gzls), B * dz, st)                   # d/dz_log_sigma (in gzls)
        # p(z|w,c) heads -> d/dw_sampled; the 0.1 bias Variable shares dense_6's bias gradient
        self._op('bneck14', 'uad_dense_bwd', ptr(br.w_s), ptr(fp.p('dense_5/kernel')), ptr(br.dM), None, 1.0, ptr(br.dws),
                 G('dense_5/kernel'), G('dense_5/bias'), B, dw, n, acc, ws, wsb, st)
        self._op('bneck14', 'uad_dense_bwd', ptr(br.w_s), ptr(fp.p('dense_6/kernel')), ptr(br.dS), None, 1.0, ptr(br.dws2),
                 G('dense_6/kernel'), G('dense_6/bias'), B, dw, n, acc, ws, wsb, st)
        if params:
            self._op('bneck14', 'uad_axpby', 1.0, ptr(fp.g('dense_6/bias')), 0.0, ptr(fp.g('Variable')), n, st)
        self._op('bneck14', 'uad_axpby', 1.0, ptr(br.dws2), 1.0, ptr(br.dws), B * dw, st)
        self._op('bneck14', 'uad_reparam_kl_bwd', ptr(br.w_mu), ptr(br.w_lsh), ptr(br.eps_w), ptr(br.dws), float(scale), ptr(br.dwmu),
                 ptr(br.dwlsh), B, dw, st)
        self._op('bneck14', 'uad_axpby', 0.5, ptr(br.dwlsh), 0.0, ptr(br.dwlsh), B * dw, st)                 # d/dw_log_sigma
        # the four heads back to the flattened bottleneck
        first = True
        for name, g_in, mask, width in (('dense_2', sm['dmu'], m['mu'], dz), ('dense_3', br.gzls, None, dz), ('dense', br.dwmu, m['wmu'], dw),
                                        ('dense_1', br.dwlsh, m['wls'], dw)):
            dst = sm['dflat'] if first else sm['dflat2']
            self._op('bneck15', 'uad_dense_bwd', ptr(br.zb), ptr(fp.p(f'Bottleneck/{name}/kernel')), ptr(g_in), ptr(mask), keep, ptr(dst),
                     G(f'Bottleneck/{name}/kernel'), G(f'Bottleneck/{name}/bias'), B, self.flat, width, acc, ws, wsb, st)
            if not first:
                self._op('bneck17', 'uad_axpby', 1.0, ptr(sm['dflat2']), 1.0, ptr(sm['dflat']), B * self.flat, st)
            first = False

    # ------------------------------------------------------------------ spatial GMVAE latent heads (models/..._spatial.py:15-39,58-63)
    def _gmvaes_latent_fwd(self, br, h):
        """1x1-conv heads on the encoder output h [B*r*r, C] (a 1x1 conv is a Dense over the rows), w / z per position, the
        64-filter ReLU layer and the two p(z|w,c) heads on w, the mixture block with one 'sample' per spatial position; the three
        prior terms are summed over the positions and averaged over the batch (trainers/GMVAE_spatial.py:77-91)."""
        fp, st = self.fp, self._st()
        ws, wsb = self._wsargs()
        B, dz, dw, n, C = self.B, self.zDim, self.dim_w, self.zDim * self.dim_c, self.enc_ch[-1]
        R = B * self.res * self.res
        for name, out, width in (('q_wz_x/w_mu', br.w_mu, dw), ('q_wz_x/w_log_sigma', br.w_ls, dw), ('q_wz_x/z_mu', br.mu, dz),
                                 ('q_wz_x/z_log_sigma', br.ls, dz)):
            self._op('bneck03', 'uad_dense_fwd', ptr(h), ptr(fp.p(name + '/kernel')), ptr(fp.p(name + '/bias')), None, 1.0, None, None,
                     ptr(out), None, R, C, width, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        self._op('bneck05', 'uad_axpby', 0.5, ptr(br.w_ls), 0.0, ptr(br.w_lsh), R * dw, st)
        self._op('bneck05', 'uad_axpby', 0.5, ptr(br.ls), 0.0, ptr(br.z_lsh), R * dz, st)
        self._op('bneck05', 'uad_reparam_kl_fwd', ptr(br.w_mu), ptr(br.w_lsh), ptr(br.eps_w), ptr(br.w_sigma), ptr(br.w_s), ptr(br.kl_w), R, dw, st)
        self._op('bneck05', 'uad_reparam_kl_fwd', ptr(br.mu), ptr(br.z_lsh), ptr(br.eps), ptr(br.sigma), ptr(br.zv), ptr(br.kl_unused), R, dz, st)
        self._op('bneck05', 'uad_dense_fwd', ptr(br.w_s), ptr(fp.p('p_z_wc/1x1convlayer/kernel')), ptr(fp.p('p_z_wc/1x1convlayer/bias')), None,
                 1.0, None, None, ptr(br.mid_pre), ptr(br.mid), R, dw, GMVAES_MID, ACT_RELU, 0.0, 1.0, ws, wsb, st)
        self._op('bneck05', 'uad_dense_fwd', ptr(br.mid), ptr(fp.p('p_z_wc/z_wc_mu/kernel')), ptr(fp.p('p_z_wc/z_wc_mu/bias')), None, 1.0,
                 None, None, ptr(br.Mz), None, R, GMVAES_MID, n, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        self._op('bneck05', 'uad_dense_fwd', ptr(br.mid), ptr(fp.p('p_z_wc/z_wc_log_sigma/kernel')), ptr(fp.p('p_z_wc/z_wc_log_sigma/bias')),
                 None, 1.0, ptr(br.ones_n), ptr(fp.p('Variable')), ptr(br.S0), ptr(br.Sz), R, GMVAES_MID, n, ACT_NONE, 0.0, 1.0, ws, wsb, st)
        self._op('bneck05', 'uad_gmvae_latent_fwd', ptr(br.mu), ptr(br.ls), ptr(br.zv), ptr(br.Mz), ptr(br.Sz), ptr(br.pc), ptr(br.con),
                 ptr(br.closs), R, dz, self.dim_c, self.c_lambda, st)
        for buf, slot in ((br.con, 5), (br.kl_w, 6), (br.closs, 7)):
            self._op('bneck08', 'uad_sum_scaled', ptr(buf), R, 1.0 / B, self.scalars[slot:].data_ptr(), ws, wsb, st)

    def _gmvaes_latent_bwd(self, br, g, params, scale, acc):
        """ADDS scale * d(con + w_loss + c_loss) / d(encoder output) to g [B*r*r, C] (which already holds the decoder's path)."""
        fp, st = self.fp, self._st()
        ws, wsb = self._wsargs()
        B, dz, dw, n, C = self.B, self.zDim, self.dim_w, self.zDim * self.dim_c, self.enc_ch[-1]
        R = B * self.res * self.res
        G = (lambda name: ptr(fp.g(name))) if params else (lambda name: None)
        self._op('bneck14', 'uad_gmvae_latent_bwd', ptr(br.mu), ptr(br.ls), ptr(br.zv), ptr(br.Mz), ptr(br.Sz), float(scale), ptr(br.gzmu),
                 ptr(br.gzls), ptr(br.gzs), ptr(br.dM), ptr(br.dS), R, dz, self.dim_c, self.c_lambda, st)
        self._op('bneck14', 'uad_reparam_kl_bwd', ptr(br.mu), ptr(br.z_lsh), ptr(br.eps), ptr(br.gzs), 0.0, ptr(br.dmu), ptr(br.dlsh), R, dz, st)
        self._op('bneck14', 'uad_axpby', 1.0, ptr(br.gzmu), 1.0, ptr(br.dmu), R * dz, st)                    # d/dz_mu
        self._op('bneck14', 'uad_axpby', 0.5, ptr(br.dlsh), 1.0, ptr(br.gzls), R * dz, st)                   # d/dz_log_sigma (in gzls)
        self._op('bneck14', 'uad_dense_bwd', ptr(br.mid), ptr(fp.p('p_z_wc/z_wc_mu/kernel')), ptr(br.dM), None, 1.0, ptr(br.dmid),
                 G('p_z_wc/z_wc_mu/kernel'), G('p_z_wc/z_wc_mu/bias'), R, GMVAES_MID, n, acc, ws, wsb, st)
        self._op('bneck14', 'uad_dense_bwd', ptr(br.mid), ptr(fp.p('p_z_wc/z_wc_log_sigma/kernel')), ptr(br.dS), None, 1.0, ptr(br.dmid2),
                 G('p_z_wc/z_wc_log_sigma/kernel'), G('p_z_wc/z_wc_log_sigma/bias'), R, GMVAES_MID, n, acc, ws, wsb, st)
        if params:
            self._op('bneck14', 'uad_axpby', 1.0, ptr(fp.g('p_z_wc/z_wc_log_sigma/bias')), 0.0, ptr(fp.g('Variable')), n, st)
        self._op('bneck14', 'uad_axpby', 1.0, ptr(br.dmid2), 1.0, ptr(br.dmid), R * GMVAES_MID, st)
        self._op('bneck14', 'uad_activation_bwd', ptr(br.dmid), ptr(br.mid_pre), ptr(br.dmid), R * GMVAES_MID, ACT_RELU, 0.0, st)
        self._op('bneck14', 'uad_dense_bwd', ptr(br.w_s), ptr(fp.p('p_z_wc/1x1convlayer/kernel')), ptr(br.dmid), None, 1.0, ptr(br.dws),
                 G('p_z_wc/1x1convlayer/kernel'), G('p_z_wc/1x1convlayer/bias'), R, dw, GMVAES_MID, acc, ws, wsb, st)
        self._op('bneck14', 'uad_reparam_kl_bwd', ptr(br.w_mu), ptr(br.w_lsh), ptr(br.eps_w), ptr(br.dws), float(scale), ptr(br.dwmu),
                 ptr(br.dwlsh), R, dw, st)
        self._op('bneck14', 'uad_axpby', 0.5, ptr(br.dwlsh), 0.0, ptr(br.dwlsh), R * dw, st)                 # d/dw_log_sigma
        h = br.enc_a[-1]
        for name, g_in, width in (('q_wz_x/z_mu', br.dmu, dz), ('q_wz_x/z_log_sigma', br.gzls, dz), ('q_wz_x/w_mu', br.dwmu, dw),
                                  ('q_wz_x/w_log_sigma', br.dwlsh, dw)):
            self._op('bneck15', 'uad_dense_bwd', ptr(h), ptr(fp.p(name + '/kernel')), ptr(g_in), None, 1.0, ptr(br.dh), G(name + '/kernel'),
                     G(name + '/bias'), R, C, width, acc, ws, wsb, st)
            self._op('bneck17', 'uad_axpby', 1.0, ptr(br.dh), 1.0, ptr(g), R * C, st)

    # ------------------------------------------------------------------ backward
    def backward(self, want_input_grad=False):
        """tf.gradients of losses['loss'] w.r.t. every trainable variable into the flat gradient buffer.

        loss = mean_b(rec + kl) (+ mean_b rec_ce for ceVAE).  With ``want_input_grad`` the x-branch also produces
        d loss_vae / d x (the ceVAE 'anomaly' term, trainers/ceVAE.py:51)."""
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self._wsargs()
        B = self.B
        scale = 1.0 / B
        r2 = self.res * self.res
        kp = self.keep_preact
        act_blk = ACT_LEAKY if kp else (ACT_LEAKY | abi.ACT_FROM_OUTPUT)
        for bi, br in enumerate(self.br):
            acc = 1 if bi > 0 else 0
            is_ce = bi > 0
            g, gn = self.gbuf
            cin = self.dec_ch[-1]
            last = self.n - 1
            lpre = f'Decoder/dec_Conv2DT_{last}'
            lbn = f'Decoder/{_bn(self.n + 1 + last)}'
            # fused: final 1x1 + L1 backward AND the BN/LeakyReLU backward of the last transposed-conv block
            self._op('dec_Conv2D_final', 'uad_final1x1_l1_bwd_fused', ptr(br.dec_z[last] if kp else br.dec_a[last]), ptr(fp.p(lbn + '/gamma')),
                     ptr(fp.p(lbn + '/beta')), ptr(fp.p('Decoder/dec_Conv2D_final/kernel')), ptr(br.x if br.target is None else br.target),
                     ptr(br.xhat), scale,
                     ptr(g), ptr(fp.g(lbn + '/gamma')), ptr(fp.g(lbn + '/beta')), ptr(fp.g(lpre + '/bias')),
                     ptr(fp.g('Decoder/dec_Conv2D_final/kernel')), ptr(fp.g('Decoder/dec_Conv2D_final/bias')), B,
                     self.S * self.S, cin, act_blk, LRELU_ALPHA, BN_C, acc, ws, wsb, st)
            s = self.S
            prev_wg = None           # completion of the previous layer's filter gradient (it reads the buffer the next dgrad overwrites)
            for i in reversed(range(self.n)):
                co = self.dec_ch[i]
                ci = self.dec_ch[i - 1] if i > 0 else self.enc_ch[-1]
                pre = f'Decoder/dec_Conv2DT_{i}'
                bnn = f'Decoder/{_bn(self.n + 1 + i)}'
                if i != self.n - 1:      # the last block's BN/activation backward is fused into the final-1x1 backward above
                  self._op(pre.split('/')[-1], 'uad_act_bn_bwd', ptr(g), ptr(br.dec_z[i] if kp else br.dec_a[i]), ptr(fp.p(bnn + '/gamma')), ptr(fp.p(bnn + '/beta')), ptr(g),
                     ptr(fp.g(bnn + '/gamma')), ptr(fp.g(bnn + '/beta')), ptr(fp.g(pre + '/bias')), B * s * s, co, act_blk,
                     LRELU_ALPHA, BN_C, acc, ws, wsb, st)
                xin = br.dec_a[i - 1] if i > 0 else br.ar
                wg = self._op_side(pre.split('/')[-1], 'uad_convT2d_wgrad', ptr(xin), ptr(g), ptr(fp.g(pre + '/kernel')), B, s // 2, s // 2, ci, co, KSIZE,
                     acc, mm, ws, wsb, st)
                self._wait_side(prev_wg)
                self._op(pre.split('/')[-1], 'uad_convT2d_dgrad', ptr(g), ptr(fp.p(pre + '/kernel')), ptr(gn), B, s // 2, s // 2, ci, co, KSIZE, mm,
                     ws, wsb, st)
                prev_wg = wg
                g, gn = gn, g
                s //= 2
            # decoder-entry BN + ReLU on the 1x1 conv output, then the 1x1 conv (as a dense over B*res*res rows)
            ctop = self.enc_ch[-1]
            dbn = f'Decoder/{_bn(self.n)}'
            self._op('dec_entry_bn', 'uad_act_bn_bwd', ptr(g), ptr(br.zr), ptr(fp.p(dbn + '/gamma')), ptr(fp.p(dbn + '/beta')), ptr(g),
                 ptr(fp.g(dbn + '/gamma')), ptr(fp.g(dbn + '/beta')), ptr(fp.g('Bottleneck/conv2d_1/bias')) if self.arch not in SPATIAL else None,
                 B * r2, ctop, ACT_RELU, 0.0, BN_C, acc, ws, wsb, st)
            if bi == len(self.br) - 1:
                self._launch_decoder_bucket()
            sm = self.small
            m = br.masks
            # keep factor is stored with the mask application: masks carry {0,1}, scale passed explicitly
            keep = self._keep
            if self.arch in SPATIAL:
                if m['sp'] is not None:
                    self._op('bneck10', 'uad_mask_scale', ptr(g), ptr(m['sp']), keep, ptr(g), B * r2 * ctop, st)
                if self.arch == GMVAES:
                    self._gmvaes_latent_bwd(br, g, True, scale, acc)
            else:
                self._op('bneck10', 'uad_dense_bwd', ptr(br.d), ptr(fp.p('Bottleneck/conv2d_1/kernel')), ptr(g), None, 1.0, ptr(sm['dd']),
                   ptr(fp.g('Bottleneck/conv2d_1/kernel')), None, B * r2, self.cb, ctop, acc, ws, wsb, st)
            if self.arch in SPATIAL:
                pass
            elif self.arch == GMVAE:
                self._gmvae_bottleneck_bwd(br, True, scale, acc)
            elif self.arch == AE:
                self._op('bneck11', 'uad_dense_bwd', ptr(br.mu), ptr(fp.p('Bottleneck/dense_1/kernel')), ptr(sm['dd']), None, 1.0,
                     ptr(sm['dmu']), ptr(fp.g('Bottleneck/dense_1/kernel')), ptr(fp.g('Bottleneck/dense_1/bias')), B,
                     self.zDim, self.flat, acc, ws, wsb, st)
                self._op('bneck12', 'uad_dense_bwd', ptr(br.zb), ptr(fp.p('Bottleneck/dense/kernel')), ptr(sm['dmu']), ptr(m['mu']), keep,
                     ptr(sm['dflat']), ptr(fp.g('Bottleneck/dense/kernel')), ptr(fp.g('Bottleneck/dense/bias')), B,
                     self.flat, self.zDim, acc, ws, wsb, st)
            else:
                zsrc = br.mu if is_ce else br.zv
                self._op('bneck13', 'uad_dense_bwd', ptr(zsrc), ptr(fp.p('Bottleneck/dense_2/kernel')), ptr(sm['dd']), ptr(m['dec']), keep,
                     ptr(sm['dzv']), ptr(fp.g('Bottleneck/dense_2/kernel')), ptr(fp.g('Bottleneck/dense_2/bias')), B,
                     self.zDim, self.flat, acc, ws, wsb, st)
                if not is_ce:
                    self._op('bneck14', 'uad_reparam_kl_bwd', ptr(br.mu), ptr(br.ls), ptr(br.eps), ptr(sm['dzv']), scale, ptr(sm['dmu']),
                         ptr(sm['dls']), B, self.zDim, st)
                    dmu = sm['dmu']
                else:
                    dmu = sm['dzv']
                self._op('bneck15', 'uad_dense_bwd', ptr(br.zb), ptr(fp.p('Bottleneck/dense/kernel')), ptr(dmu), ptr(m['mu']), keep,
                     ptr(sm['dflat']), ptr(fp.g('Bottleneck/dense/kernel')), ptr(fp.g('Bottleneck/dense/bias')), B,
                     self.flat, self.zDim, acc, ws, wsb, st)
                if not is_ce:
                    self._op('bneck16', 'uad_dense_bwd', ptr(br.zb), ptr(fp.p('Bottleneck/dense_1/kernel')), ptr(sm['dls']), ptr(m['ls']),
                         keep, ptr(sm['dflat2']), ptr(fp.g('Bottleneck/dense_1/kernel')),
                         ptr(fp.g('Bottleneck/dense_1/bias')), B, self.flat, self.zDim, acc, ws, wsb, st)
                    self._op('bneck17', 'uad_axpby', 1.0, ptr(sm['dflat2']), 1.0, ptr(sm['dflat']), B * self.flat, st)
            # bottleneck 1x1 conv backward -> gradient w.r.t. the last encoder activation
            if self.arch not in SPATIAL:
                self._op('bneck18', 'uad_dense_bwd', ptr(br.enc_a[-1]), ptr(fp.p('Bottleneck/conv2d/kernel')), ptr(sm['dflat']), None, 1.0,
                   ptr(g), ptr(fp.g('Bottleneck/conv2d/kernel')), ptr(fp.g('Bottleneck/conv2d/bias')), B * r2, ctop, self.cb,
                   acc, ws, wsb, st)
            s = self.res
            for i in reversed(range(self.n)):
                co = self.enc_ch[i]
                ci = self.enc_ch[i - 1] if i > 0 else 1
                pre = f'Encoder/enc_conv2D_{i}'
                bnn = f'Encoder/{_bn(i)}'
                self._op(pre.split('/')[-1], 'uad_act_bn_bwd', ptr(g), ptr(br.enc_z[i] if kp else br.enc_a[i]), ptr(fp.p(bnn + '/gamma')), ptr(fp.p(bnn + '/beta')), ptr(g),
                     ptr(fp.g(bnn + '/gamma')), ptr(fp.g(bnn + '/beta')), ptr(fp.g(pre + '/bias')), B * s * s, co, act_blk,
                     LRELU_ALPHA, BN_C, acc, ws, wsb, st)
                xin = br.enc_a[i - 1] if i > 0 else br.x
                wg = self._op_side(pre.split('/')[-1], 'uad_conv2d_wgrad', ptr(xin), ptr(g), ptr(fp.g(pre + '/kernel')), B, 2 * s, 2 * s, ci, co, KSIZE, acc,
                     mm, ws, wsb, st)
                if i > 0:
                    self._wait_side(prev_wg)
                    self._op(pre.split('/')[-1], 'uad_conv2d_dgrad', ptr(g), ptr(fp.p(pre + '/kernel')), ptr(gn), B, 2 * s, 2 * s, ci, co, KSIZE, mm,
                         ws, wsb, st)
                    g, gn = gn, g
                elif want_input_grad and bi == 0:
                    self._op(pre.split('/')[-1], 'uad_conv2d_dgrad', ptr(g), ptr(fp.p(pre + '/kernel')), ptr(self.gx), B, 2 * s, 2 * s, ci, co,
                         KSIZE, mm, ws, wsb, st)
                prev_wg = wg
                s *= 2
            # the next branch (ceVAE) starts by overwriting the gradient buffers the last filter gradients still read
            self._join_side()

    _keep = 1.0

    # ------------------------------------------------------------------ constrained AE backward
    def _encoder_backward(self, br, g, gn, acc, dx_out):
        """Reverse of the encoder stack for one pass (BN/LeakyReLU backward, wgrad, dgrad per block); g holds the gradient
        w.r.t. the last encoder activation.  dx_out (nullable): receives d/d(input image) (first layer dgrad, Cin = 1)."""
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self._wsargs()
        B = self.B
        kp = self.keep_preact
        act_blk = ACT_LEAKY if kp else (ACT_LEAKY | abi.ACT_FROM_OUTPUT)
        s = self.res
        for i in reversed(range(self.n)):
            co = self.enc_ch[i]
            ci = self.enc_ch[i - 1] if i > 0 else 1
            pre = f'Encoder/enc_conv2D_{i}'
            bnn = f'Encoder/{_bn(i)}'
            self._op(pre.split('/')[-1], 'uad_act_bn_bwd', ptr(g), ptr(br.enc_z[i] if kp else br.enc_a[i]), ptr(fp.p(bnn + '/gamma')),
                     ptr(fp.p(bnn + '/beta')), ptr(g), ptr(fp.g(bnn + '/gamma')), ptr(fp.g(bnn + '/beta')), ptr(fp.g(pre + '/bias')),
                     B * s * s, co, act_blk, LRELU_ALPHA, BN_C, acc, ws, wsb, st)
            xin = br.enc_a[i - 1] if i > 0 else br.x
            self._op(pre.split('/')[-1], 'uad_conv2d_wgrad', ptr(xin), ptr(g), ptr(fp.g(pre + '/kernel')), B, 2 * s, 2 * s, ci, co, KSIZE,
                     acc, mm, ws, wsb, st)
            if i > 0:
                self._op(pre.split('/')[-1], 'uad_conv2d_dgrad', ptr(g), ptr(fp.p(pre + '/kernel')), ptr(gn), B, 2 * s, 2 * s, ci, co, KSIZE,
                         mm, ws, wsb, st)
                g, gn = gn, g
            elif dx_out is not None:
                self._op(pre.split('/')[-1], 'uad_conv2d_dgrad', ptr(g), ptr(fp.p(pre + '/kernel')), ptr(dx_out), B, 2 * s, 2 * s, ci, co,
                         KSIZE, mm, ws, wsb, st)
            s *= 2

    def backward_constrained(self):
        """tf.gradients of loss = mean_b(L2 + rho*Rec_z) (trainers/ConstrainedAE.py:37-43) through
        x -> Enc -> z -> Dec -> x_hat -> Enc -> z_rec (models/constrained_autoencoder.py:12-46, shared weights).
        Order: the re-encoding pass first (it yields d/d x_hat), then decoder + bottleneck + first encoder pass, whose
        Encoder / Bottleneck-dense gradients ACCUMULATE onto the re-encoding pass's."""
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self._wsargs()
        B, b0, b1, sm = self.B, self.br[0], self.br[1], self.small
        r2 = self.res * self.res
        ctop = self.enc_ch[-1]
        keep = self._keep
        kp = self.keep_preact
        act_blk = ACT_LEAKY if kp else (ACT_LEAKY | abi.ACT_FROM_OUTPUT)
        g, gn = self.gbuf
        # ---- pass 2 (x_hat -> z_rec): seed dzrec = d loss / d z_rec (left by forward)
        self._op('bneck20', 'uad_dense_bwd', ptr(b1.zb), ptr(fp.p('Bottleneck/dense/kernel')), ptr(self.dzrec), ptr(b1.masks['mu']), keep,
                 ptr(sm['dflat']), ptr(fp.g('Bottleneck/dense/kernel')), ptr(fp.g('Bottleneck/dense/bias')), B, self.flat, self.zDim,
                 0, ws, wsb, st)
        self._op('bneck21', 'uad_dense_bwd', ptr(b1.enc_a[-1]), ptr(fp.p('Bottleneck/conv2d/kernel')), ptr(sm['dflat']), None, 1.0, ptr(g),
                 ptr(fp.g('Bottleneck/conv2d/kernel')), ptr(fp.g('Bottleneck/conv2d/bias')), B * r2, ctop, self.cb, 0, ws, wsb, st)
        self._encoder_backward(b1, g, gn, 0, self.gx)
        # d loss / d x_hat = MSE term (in gxhat) + the path through the re-encoding pass (in gx)
        self._op('bneck22', 'uad_axpby', 1.0, ptr(self.gx), 1.0, ptr(self.gxhat), b0.x.numel(), st)
        self.backward_from_gxhat(acc=1, dzrec=self.dzrec)

    def backward_from_gxhat(self, acc=0, dzrec=None):
        """Decoder + bottleneck + encoder reverse pass of the dense AE whose BOTH bottleneck Dropouts honour the flag, seeded
        with d loss / d x_hat in ``gxhat``.  acc: the Encoder / Bottleneck-dense gradients accumulate onto what a preceding pass
        left (constrained AE); dzrec: an extra -dzrec on d/dz (constrained AE).  With acc=0, dzrec=None this is the whole
        backward of the adversarial AE's reconstruction loss (trainers/AAE.py:55-57,67)."""
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self._wsargs()
        B, b0, sm = self.B, self.br[0], self.small
        r2 = self.res * self.res
        ctop = self.enc_ch[-1]
        keep = self._keep
        kp = self.keep_preact
        act_blk = ACT_LEAKY if kp else (ACT_LEAKY | abi.ACT_FROM_OUTPUT)
        # ---- decoder
        g, gn = self.gbuf
        cin = self.dec_ch[-1]
        self._op('dec_Conv2D_final', 'uad_final1x1_bwd', ptr(b0.dec_a[-1]), ptr(fp.p('Decoder/dec_Conv2D_final/kernel')), ptr(self.gxhat),
                 ptr(g), ptr(fp.g('Decoder/dec_Conv2D_final/kernel')), ptr(fp.g('Decoder/dec_Conv2D_final/bias')), B, self.S * self.S,
                 cin, 0, ws, wsb, st)
        s = self.S
        for i in reversed(range(self.n)):
            co = self.dec_ch[i]
            ci = self.dec_ch[i - 1] if i > 0 else ctop
            pre = f'Decoder/dec_Conv2DT_{i}'
            bnn = f'Decoder/{_bn(self.n + 1 + i)}'
            self._op(pre.split('/')[-1], 'uad_act_bn_bwd', ptr(g), ptr(b0.dec_z[i] if kp else b0.dec_a[i]), ptr(fp.p(bnn + '/gamma')),
                     ptr(fp.p(bnn + '/beta')), ptr(g), ptr(fp.g(bnn + '/gamma')), ptr(fp.g(bnn + '/beta')), ptr(fp.g(pre + '/bias')),
                     B * s * s, co, act_blk, LRELU_ALPHA, BN_C, 0, ws, wsb, st)
            xin = b0.dec_a[i - 1] if i > 0 else b0.ar
            self._op(pre.split('/')[-1], 'uad_convT2d_wgrad', ptr(xin), ptr(g), ptr(fp.g(pre + '/kernel')), B, s // 2, s // 2, ci, co, KSIZE,
                     0, mm, ws, wsb, st)
            self._op(pre.split('/')[-1], 'uad_convT2d_dgrad', ptr(g), ptr(fp.p(pre + '/kernel')), ptr(gn), B, s // 2, s // 2, ci, co, KSIZE,
                     mm, ws, wsb, st)
            g, gn = gn, g
            s //= 2
        dbn = f'Decoder/{_bn(self.n)}'
        self._op('dec_entry_bn', 'uad_act_bn_bwd', ptr(g), ptr(b0.zr), ptr(fp.p(dbn + '/gamma')), ptr(fp.p(dbn + '/beta')), ptr(g),
                 ptr(fp.g(dbn + '/gamma')), ptr(fp.g(dbn + '/beta')), ptr(fp.g('Bottleneck/conv2d_1/bias')), B * r2, ctop, ACT_RELU, 0.0,
                 BN_C, 0, ws, wsb, st)
        # ---- bottleneck: conv2d_1, dec_dense (dense_1, Dropout honoured), then z = Dropout(dense(...)) with the extra -dzrec
        self._op('bneck10', 'uad_dense_bwd', ptr(b0.d), ptr(fp.p('Bottleneck/conv2d_1/kernel')), ptr(g), None, 1.0, ptr(sm['dd']),
                 ptr(fp.g('Bottleneck/conv2d_1/kernel')), None, B * r2, self.cb, ctop, 0, ws, wsb, st)
        self._op('bneck11', 'uad_dense_bwd', ptr(b0.mu), ptr(fp.p('Bottleneck/dense_1/kernel')), ptr(sm['dd']), ptr(b0.masks['dec']), keep,
                 ptr(sm['dmu']), ptr(fp.g('Bottleneck/dense_1/kernel')), ptr(fp.g('Bottleneck/dense_1/bias')), B, self.zDim, self.flat,
                 0, ws, wsb, st)
        if dzrec is not None:
            self._op('bneck23', 'uad_axpby', -1.0, ptr(dzrec), 1.0, ptr(sm['dmu']), B * self.zDim, st)     # d Rec_z / d z = -d/d z_rec
        self._op('bneck12', 'uad_dense_bwd', ptr(b0.zb), ptr(fp.p('Bottleneck/dense/kernel')), ptr(sm['dmu']), ptr(b0.masks['mu']), keep,
                 ptr(sm['dflat']), ptr(fp.g('Bottleneck/dense/kernel')), ptr(fp.g('Bottleneck/dense/bias')), B, self.flat, self.zDim,
                 acc, ws, wsb, st)
        self._op('bneck18', 'uad_dense_bwd', ptr(b0.enc_a[-1]), ptr(fp.p('Bottleneck/conv2d/kernel')), ptr(sm['dflat']), None, 1.0, ptr(g),
                 ptr(fp.g('Bottleneck/conv2d/kernel')), ptr(fp.g('Bottleneck/conv2d/bias')), B * r2, ctop, self.cb, acc, ws, wsb, st)
        # ---- pass 1 encoder (accumulating onto the re-encoding pass when acc = 1)
        self._encoder_backward(b0, g, gn, acc, None)

    # ------------------------------------------------------------------ gradient w.r.t. the input only (restoration)
    def backward_to_input(self, seed, kl_scale=1.0):
        """d/dx of  <seed, x_hat(x)> + kl_scale * sum_b kl_b(x)  through the x-branch: the dgrad-only chain (no weight
        gradients are formed, every parameter-gradient output is NULL).  Needs the activations of a preceding
        ``forward`` (training or not): the BN/activation backward works from the block outputs (UAD_ACT_FROM_OUTPUT).
        Result in ``self.gx``.  This is tf.gradients(..., self.x) of reference trainers/VAE_You.py:54."""
        fp, st, mm = self.fp, self._st(), self.math_mode
        ws, wsb = self._wsargs()
        B, br, sm = self.B, self.br[0], self.small
        r2 = self.res * self.res
        g, gn = self.gbuf
        FO = abi.ACT_FROM_OUTPUT
        cin = self.dec_ch[-1]
        self._op('dec_Conv2D_final', 'uad_final1x1_bwd', ptr(br.dec_a[-1]), ptr(fp.p('Decoder/dec_Conv2D_final/kernel')), ptr(seed),
                 ptr(g), None, None, B, self.S * self.S, cin, 0, ws, wsb, st)
        s = self.S
        for i in reversed(range(self.n)):
            co = self.dec_ch[i]
            ci = self.dec_ch[i - 1] if i > 0 else self.enc_ch[-1]
            pre = f'Decoder/dec_Conv2DT_{i}'
            bnn = f'Decoder/{_bn(self.n + 1 + i)}'
            self._op(pre.split('/')[-1], 'uad_act_bn_bwd', ptr(g), ptr(br.dec_a[i]), ptr(fp.p(bnn + '/gamma')), ptr(fp.p(bnn + '/beta')),
                     ptr(g), None, None, None, B * s * s, co, ACT_LEAKY | FO, LRELU_ALPHA, BN_C, 0, ws, wsb, st)
            self._op(pre.split('/')[-1], 'uad_convT2d_dgrad', ptr(g), ptr(fp.p(pre + '/kernel')), ptr(gn), B, s // 2, s // 2, ci, co,
                     KSIZE, mm, ws, wsb, st)
            g, gn = gn, g
            s //= 2
        ctop = self.enc_ch[-1]
        dbn = f'Decoder/{_bn(self.n)}'
        self._op('dec_entry_bn', 'uad_act_bn_bwd', ptr(g), ptr(br.ar), ptr(fp.p(dbn + '/gamma')), ptr(fp.p(dbn + '/beta')), ptr(g),
                 None, None, None, B * r2, ctop, ACT_RELU | FO, 0.0, BN_C, 0, ws, wsb, st)
        m, keep = br.masks, self._keep
        if self.arch in SPATIAL:
            if m['sp'] is not None:
                self._op('bneck10', 'uad_mask_scale', ptr(g), ptr(m['sp']), keep, ptr(g), B * r2 * ctop, st)
            if self.arch == GMVAES:
                self._gmvaes_latent_bwd(br, g, False, float(kl_scale), 0)
        else:
            self._op('bneck10', 'uad_dense_bwd', ptr(br.d), ptr(fp.p('Bottleneck/conv2d_1/kernel')), ptr(g), None, 1.0, ptr(sm['dd']),
                     None, None, B * r2, self.cb, ctop, 0, ws, wsb, st)
        if self.arch in SPATIAL:
            pass
        elif self.arch == GMVAE:
            self._gmvae_bottleneck_bwd(br, False, float(kl_scale), 0)
        elif self.arch == AE:
            self._op('bneck11', 'uad_dense_bwd', ptr(br.mu), ptr(fp.p('Bottleneck/dense_1/kernel')), ptr(sm['dd']), None, 1.0,
                     ptr(sm['dmu']), None, None, B, self.zDim, self.flat, 0, ws, wsb, st)
            self._op('bneck12', 'uad_dense_bwd', ptr(br.zb), ptr(fp.p('Bottleneck/dense/kernel')), ptr(sm['dmu']), ptr(m['mu']), keep,
                     ptr(sm['dflat']), None, None, B, self.flat, self.zDim, 0, ws, wsb, st)
        else:
            self._op('bneck13', 'uad_dense_bwd', ptr(br.zv), ptr(fp.p('Bottleneck/dense_2/kernel')), ptr(sm['dd']), ptr(m['dec']), keep,
                     ptr(sm['dzv']), None, None, B, self.zDim, self.flat, 0, ws, wsb, st)
            self._op('bneck14', 'uad_reparam_kl_bwd', ptr(br.mu), ptr(br.ls), ptr(br.eps), ptr(sm['dzv']), float(kl_scale),
                     ptr(sm['dmu']), ptr(sm['dls']), B, self.zDim, st)
            self._op('bneck15', 'uad_dense_bwd', ptr(br.zb), ptr(fp.p('Bottleneck/dense/kernel')), ptr(sm['dmu']), ptr(m['mu']), keep,
                     ptr(sm['dflat']), None, None, B, self.flat, self.zDim, 0, ws, wsb, st)
            self._op('bneck16', 'uad_dense_bwd', ptr(br.zb), ptr(fp.p('Bottleneck/dense_1/kernel')), ptr(sm['dls']), ptr(m['ls']), keep,
                     ptr(sm['dflat2']), None, None, B, self.flat, self.zDim, 0, ws, wsb, st)
            self._op('bneck17', 'uad_axpby', 1.0, ptr(sm['dflat2']), 1.0, ptr(sm['dflat']), B * self.flat, st)
        if self.arch not in SPATIAL:
            self._op('bneck18', 'uad_dense_bwd', ptr(br.enc_a[-1]), ptr(fp.p('Bottleneck/conv2d/kernel')), ptr(sm['dflat']), None, 1.0,
                     ptr(g), None, None, B * r2, ctop, self.cb, 0, ws, wsb, st)
        s = self.res
        for i in reversed(range(self.n)):
            co = self.enc_ch[i]
            ci = self.enc_ch[i - 1] if i > 0 else 1
            pre = f'Encoder/enc_conv2D_{i}'
            bnn = f'Encoder/{_bn(i)}'
            self._op(pre.split('/')[-1], 'uad_act_bn_bwd', ptr(g), ptr(br.enc_a[i]), ptr(fp.p(bnn + '/gamma')), ptr(fp.p(bnn + '/beta')),
                     ptr(g), None, None, None, B * s * s, co, ACT_LEAKY | FO, LRELU_ALPHA, BN_C, 0, ws, wsb, st)
            dst = gn if i > 0 else self.gx
            self._op(pre.split('/')[-1], 'uad_conv2d_dgrad', ptr(g), ptr(fp.p(pre + '/kernel')), ptr(dst), B, 2 * s, 2 * s, ci, co, KSIZE,
                     mm, ws, wsb, st)
            g, gn = gn, g
            s *= 2

    def restore_step(self, restore_lr, tv_lambda, dropout=False, dropout_rate=0.0, parity_noise=False, keep_grads=False):
        """One iteration of the MAP restoration loop of reference trainers/VAE_You.py:125-139 on the batch resident in
        ``br[0].x`` (updated IN PLACE):  x <- x - restore_lr * d/dx [ sum|x_hat-x| + kl + tv_lambda*TV(x - x_hat) ].
        The whole iteration stays on the device (the reference pays one sess.run + two host<->device image copies per step)."""
        br = self.br[0]
        rate = dropout_rate if dropout else 0.0
        self._keep = 1.0 / (1.0 - rate) if rate > 0 else 1.0
        if not parity_noise:
            self.draw_noise(dropout, rate)
        self.forward(training=False, dropout_rate=rate, branches=[0], need_l1=False)
        st = self._st()
        ws, wsb = self._wsargs()
        if not hasattr(self, 'gseed'):
            self.gseed = self._new(self.B, self.S, self.S, 1)
            self.tv = self._new(self.B)
            self.restore_grads = self._new(self.B, self.S, self.S, 1)
        self._op('restore', 'uad_tv_restore_seed', ptr(br.x), ptr(br.xhat), float(tv_lambda), ptr(self.gseed), ptr(self.tv), self.B,
                 self.S, self.S, ws, wsb, st)
        self.backward_to_input(self.gseed, kl_scale=1.0)
        self._op('restore', 'uad_restore_update', ptr(br.x), ptr(self.gx), ptr(self.gseed), float(restore_lr),
                 ptr(self.restore_grads) if keep_grads else None, br.x.numel(), st)

    def restore(self, steps, restore_lr, tv_lambda, dropout=False, dropout_rate=0.0, use_graph=True):
        """``steps`` restoration iterations; after one eager iteration the rest replay a CUDA graph of the iteration
        (the Philox offset lives on the device, so every replay draws a fresh eps as tf.random_normal does)."""
        if steps <= 0:
            return
        if not hasattr(self, 'gseed'):
            self.gseed = self._new(self.B, self.S, self.S, 1)
            self.tv = self._new(self.B)
            self.restore_grads = self._new(self.B, self.S, self.S, 1)
        key = (float(restore_lr), float(tv_lambda), bool(dropout), float(dropout_rate))
        done = 0
        if not use_graph or getattr(self, '_restore_key', None) != key:
            self.restore_step(restore_lr, tv_lambda, dropout, dropout_rate)
            done = 1
            self._restore_graph = None
            self._restore_key = key
        if not use_graph:
            for _ in range(done, steps):
                self.restore_step(restore_lr, tv_lambda, dropout, dropout_rate)
            return
        if self._restore_graph is None and steps > done:
            g = torch.cuda.CUDAGraph()
            with graph_capture(g):
                self.restore_step(restore_lr, tv_lambda, dropout, dropout_rate)
            self._restore_graph = g
        for _ in range(done, steps):
            self._restore_graph.replay()

    # ------------------------------------------------------------------ optimiser
    def adam_step(self, lr, beta1=0.5, beta2=0.999, eps=1e-8, grad_scale=1.0, lo=0, hi=None):
        """tf.train.AdamOptimizer update on (a slice of) the flat buffers (reference trainers/DLMODEL.py:112-131).
        The step counter lives on the device so the call is CUDA-graph capturable."""
        self.t += 1
        hi = self.fp.numel if hi is None else hi
        fp, st = self.fp, self._st()
        call('uad_counter_add', self.step_dev.data_ptr(), 1, st)
        call('uad_adam_tf_step', ptr(fp.params[lo:]), ptr(fp.grads[lo:]), ptr(fp.m[lo:]), ptr(fp.v[lo:]), hi - lo,
             lr, beta1, beta2, eps, grad_scale, self.step_dev.data_ptr(), st)

    def enable_peer_optimizer(self):
        """Data parallel: replace `all-reduce + Adam` by the single peer-memory kernel (dist.PeerOptimizer).  Call once, after the
        initial parameter broadcast and before the first train step."""
        from . import dist as udist
        if self.peer is None and udist.world_size() > 1:
            self.peer = udist.PeerOptimizer(self.fp, self.device)
            self.graph, self._warm = None, None
        return self.peer

    def adam_step_peer(self, lr, beta1=0.5, beta2=0.999, eps=1e-8, grad_scale=1.0):
        self.t += 1
        call('uad_counter_add', self.step_dev.data_ptr(), 1, self._st())
        self.peer.step(self.fp.m, self.fp.v, lr, beta1, beta2, eps, grad_scale, self.step_dev, self._st())

    def _reduce_and_update(self, allreduce, lr, beta1, world):
        """[gradient all-reduce] + TF-Adam: NCCL + the Adam kernel, or the fused peer-memory kernel."""
        if allreduce is not None and self.peer is not None:
            self.adam_step_peer(lr, beta1=beta1, grad_scale=1.0 / world)
            return
        if allreduce is not None:
            self._finish_allreduce(allreduce)
        self.adam_step(lr, beta1=beta1, grad_scale=1.0 / world)

    # ------------------------------------------------------------------ one train step (process(TRAIN) body)
    def _fwd_bwd(self, rate, dropout, parity_noise, want_anomaly):
        if not parity_noise:
            self.draw_noise(dropout, rate)
        self.forward(training=True, dropout_rate=rate)
        if self.arch in (CAE, CAAE):
            self.backward_constrained()
            return
        if self.arch == AAE:
            self.backward_from_gxhat()
            return
        self.backward(want_input_grad=want_anomaly)
        if want_anomaly and self.arch == CEVAE:
            self._finish_anomaly()

    def train_step(self, lr, beta1=0.5, dropout_rate=0.0, dropout=True, allreduce=None, world=1, parity_noise=False,
                   want_anomaly=False, use_graph=False):
        """Body of process(TRAIN) for one mini-batch already staged with set_inputs():
        noise -> forward -> losses -> backward -> [gradient all-reduce] -> TF-Adam.

        use_graph: the first call runs eagerly (warm-up), the second captures a CUDA graph, later calls replay it
        (the ~90 kernel launches of a step collapse into one graph launch)."""
        rate = dropout_rate if dropout else 0.0
        self._keep = 1.0 / (1.0 - rate) if rate > 0 else 1.0
        # decoder gradient bucket in flight during the encoder's backward (needs the collective's non-blocking form, dist.py)
        self._bucket_async = getattr(allreduce, 'async_', None) if (allreduce is not None and self.dp_buckets and world > 1 and
                                                                    self.peer is None) else None
        self._bucket_work = None
        key = (lr, beta1, rate, dropout, world, want_anomaly, allreduce is None)
        if use_graph and not parity_noise and self._warm == key:
            if self.graph is None:
                # data parallel: the gradient all-reduce (NCCL is stream-capturable; the communicator exists since the eager
                # warm-up step) and the Adam update CAN be part of the captured step (UAD_GRAPH_ALLREDUCE=1; bit-identical
                # results, tests/dp_equiv_worker.py).  Measured on 2 x B200 (profiles/r2_dp_2gpu.md): 4.455 ms per step inside
                # the graph, 4.425 ms with the collective issued behind the replay (1 GPU: 4.346) - the default stays outside.
                self._graph_has_update = (allreduce is None or self._bucket_async is not None or self.peer is not None or
                                          os.environ.get('UAD_GRAPH_ALLREDUCE', '0') != '0')
                t_save = self.t
                try:
                    g = torch.cuda.CUDAGraph()
                    with graph_capture(g):
                        self._fwd_bwd(rate, dropout, False, want_anomaly)
                        if self._graph_has_update:
                            self._reduce_and_update(allreduce, lr, beta1, world)
                except Exception:
                    if allreduce is None or not self._graph_has_update or self.peer is not None:
                        raise
                    torch.cuda.synchronize(self.device)           # a collective that cannot be captured here: capture without it
                    self._graph_has_update = False
                    self._bucket_async, self._bucket_work = None, None
                    g = torch.cuda.CUDAGraph()
                    with graph_capture(g):
                        self._fwd_bwd(rate, dropout, False, want_anomaly)
                self.t = t_save
                self.graph = g
            self.graph.replay()
            if not self._graph_has_update:
                allreduce(self.fp.grads)
                self.adam_step(lr, beta1=beta1, grad_scale=1.0 / world)
            else:
                self.t += 1
            return
        self.graph = None
        self._fwd_bwd(rate, dropout, parity_noise, want_anomaly)
        self._reduce_and_update(allreduce, lr, beta1, world)
        self._warm = key

    _warm = None

    def _finish_anomaly(self):
        """anomaly = L1_vae * |d loss_vae / d x| (ceVAE.py:51): add the direct term -sign(xhat - x)/B of the L1 on x."""
        b0 = self.br[0]
        st = self._st()
        # gx currently holds the encoder-path term; the L1's own dependence on x contributes -dL/dxhat.
        ws, wsb = self._wsargs()
        call('uad_l1_direct_term', ptr(b0.x), ptr(b0.xhat), 1.0 / self.B, ptr(self.gx), b0.x.numel(), st)
        call('uad_mul_abs', ptr(b0.l1), ptr(self.gx), ptr(self.anomaly), b0.x.numel(), st)

    def anomaly_per_sample(self):
        """ceVAE 'anomaly' map as ``ceVAE.reconstruct`` fetches it (reference trainers/ceVAE.py:51,119-139): the reference
        evaluates one slice per ``sess.run``, so ``loss_vae = mean_b(rec_vae + kl)`` is normalised by 1/1 - every slice of
        the stack gets the gradient of ITS OWN ``sum|x_hat - x| + kl`` (no 1/B).  Needs a preceding
        ``forward(branches=[0], need_l1=True)``; dgrad-only chain, no parameter gradients.  Result in ``self.anomaly``."""
        b0, st = self.br[0], self._st()
        ws, wsb = self._wsargs()
        if not hasattr(self, 'gseed'):
            self.gseed = self._new(self.B, self.S, self.S, 1)
            self.tv = self._new(self.B)
            self.restore_grads = self._new(self.B, self.S, self.S, 1)
        # seed = d sum|x_hat - x| / d x_hat = sign(x_hat - x)  (the restoration seed kernel with tv_lambda = 0)
        self._op('anomaly', 'uad_tv_restore_seed', ptr(b0.x), ptr(b0.xhat), 0.0, ptr(self.gseed), ptr(self.tv), self.B, self.S, self.S,
                 ws, wsb, st)
        self.backward_to_input(self.gseed, kl_scale=1.0)
        call('uad_l1_direct_term', ptr(b0.x), ptr(b0.xhat), 1.0, ptr(self.gx), b0.x.numel(), st)
        call('uad_mul_abs', ptr(b0.l1), ptr(self.gx), ptr(self.anomaly), b0.x.numel(), st)

    # ------------------------------------------------------------------ read-back
    def losses(self):
        s = self.scalars.detach().cpu().numpy()
        if self.arch in (AE, AES):
            return {'reconstructionLoss': float(s[0]), 'loss': float(s[0])}
        if self.arch in (CAE, CAAE):
            return {'reconstructionLoss': float(s[0]), 'L2': float(s[4]), 'Rec_z': float(s[5]),
                    'loss': float(s[4]) + self.rho * float(s[5])}
        if self.arch == AAE:
            return {'reconstructionLoss': float(s[0]), 'L2': float(s[4]), 'loss': float(s[4])}
        if self.arch == VAE:
            return {'reconstructionLoss': float(s[0]), 'kl': float(s[1]), 'loss': float(s[2])}
        if self.arch in (GMVAE, GMVAES):
            return {'reconstructionLoss': float(s[0]), 'mean_p_loss': float(s[0]), 'conditional_prior_loss': float(s[5]),
                    'w_prior_loss': float(s[6]), 'c_prior_loss': float(s[7]), 'loss': float(s[0]) + float(s[5]) + float(s[6]) + float(s[7])}
        return {'Rec_vae': float(s[0]), 'kl': float(s[1]), 'loss_vae': float(s[2]), 'Rec_ce': float(s[3]),
                'reconstructionLoss': 0.5 * float(s[0] + s[3]), 'loss': float(s[2] + s[3])}
