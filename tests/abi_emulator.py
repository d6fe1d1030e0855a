"""CPU emulation of the C ABI (include/uad_b200.h) for HOST-SIDE COMPOSITION tests - test infrastructure only.

What it is for: the engines (fanogan_engine, anovaegan_engine, ...) are sequences of ABI calls on caller-owned buffers.  Each
kernel is parity-tested on the GPU on its own (tests/test_gpu_ops.py, test_gpu_fanogan.py); whether a NEW sequence of calls
computes the right gradients is a property of the host code and can be checked without a GPU if every entry point is replaced
by a float64 torch-CPU function with the semantics the header documents.  `install(monkeypatch, module, ...)` swaps the
module-level `call` / `ptr` of an engine module for the emulated ones (``ptr`` then hands the tensor itself through) and
makes the engine allocate on the CPU.  The emulator is validated the other way round: the f-AnoGAN engine, whose real-kernel
runs match the oracle on the B200, must also match the oracle through the emulator (tests/test_engine_emulated.py).

The product never imports this file; nothing here is a fallback path."""
import math

import numpy as np
import torch

from oracle.tf_graph_cpu import conv2d_same_s2, conv2dT_same_s2

ACT_NONE, ACT_LEAKY, ACT_RELU, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3, 4
D = torch.float64

_registry = []          # tensors whose raw data_ptr() the engines pass (scalars, counters)


def register(*tensors):
    _registry.extend(tensors)


_CTYPES = {torch.float32: 'c_float', torch.uint8: 'c_uint8', torch.int64: 'c_int64'}


def _raw(address, n, dtype=torch.float32):
    """A tensor VIEW over n elements of host memory at a raw address (what `tensor.data_ptr()` hands the ABI when the engine
    lives on the CPU); the caller's sizes come from the call's own arguments, exactly as the kernels take them."""
    import ctypes
    buf = (getattr(ctypes, _CTYPES[dtype]) * int(n)).from_address(int(address))
    return torch.from_numpy(np.ctypeslib.as_array(buf))


def _resolve(p, n=None, dtype=torch.float32):
    """pointer argument -> flat view starting there: a tensor (ptr() patched to pass tensors through), a registered tensor's
    interior, or - given the element count - raw host memory."""
    if p is None or isinstance(p, torch.Tensor):
        return None if p is None else p.reshape(-1)
    for t in _registry:
        base, size = t.data_ptr(), t.numel() * t.element_size()
        if base <= p < base + size:
            return t.reshape(-1)[(p - base) // t.element_size():]
    if n is None:
        raise KeyError(f'pointer {p:#x} is not inside a registered tensor')
    return _raw(p, n, dtype)


def _v(t, *shape, dtype=torch.float32):
    """float64 copy of the first prod(shape) elements behind a pointer, shaped."""
    n = int(np.prod(shape))
    return _resolve(t, n, dtype)[:n].to(D).reshape(*shape)


def _w(t, value, dtype=torch.float32):
    """write value (any shape) to the first numel elements behind a pointer."""
    if t is None:
        return
    flat = _resolve(t, value.numel(), dtype)
    flat[:value.numel()].copy_(value.reshape(-1).to(flat.dtype))


def _acc(t, value, accumulate):
    if t is None:
        return
    if accumulate:
        value = value + _v(t, *value.shape)
    _w(t, value)


def _act(n, act, alpha):
    act &= 0xff
    if act == ACT_NONE:
        return n
    if act == ACT_LEAKY:
        return torch.where(n > 0, n, alpha * n)
    if act == ACT_RELU:
        return torch.where(n > 0, n, torch.zeros_like(n))
    if act == ACT_SIGMOID:
        return torch.sigmoid(n)
    if act == ACT_TANH:
        return torch.tanh(n)
    raise ValueError(act)


def _dact(n, act, alpha):
    act &= 0xff
    if act == ACT_NONE:
        return torch.ones_like(n)
    if act == ACT_LEAKY:
        return torch.where(n > 0, torch.ones_like(n), torch.full_like(n, alpha))
    if act == ACT_RELU:
        return (n > 0).to(n.dtype)
    if act == ACT_SIGMOID:
        s = torch.sigmoid(n)
        return s * (1 - s)
    if act == ACT_TANH:
        return 1 - torch.tanh(n) ** 2
    raise ValueError(act)


def _nchw(t, B, H, W, C):
    return _v(t, B, H, W, C).permute(0, 3, 1, 2)


def _affine(z_nhwc, gamma, beta, C, bn_c):
    if gamma is None:
        return z_nhwc
    return z_nhwc * (_v(gamma, C) * bn_c) + _v(beta, C)


# ------------------------------------------------------------------------------------------------ conv family
def uad_conv2d_fwd(x, w, bias, gamma, beta, z_out, a_out, B, H, W, Cin, Cout, k, act, alpha, bn_c, mm, ws, wsb, st):
    b = _v(bias, Cout) if bias is not None else torch.zeros(Cout, dtype=D)
    z = conv2d_same_s2(_nchw(x, B, H, W, Cin), _v(w, k, k, Cin, Cout), b).permute(0, 2, 3, 1)
    _w(z_out, z)
    _w(a_out, _act(_affine(z, gamma, beta, Cout, bn_c), act, alpha))


def uad_conv2d_dgrad(dz, w, dx, B, H, W, Cin, Cout, k, mm, ws, wsb, st):
    x = torch.zeros(B, Cin, H, W, dtype=D, requires_grad=True)
    y = conv2d_same_s2(x, _v(w, k, k, Cin, Cout), torch.zeros(Cout, dtype=D))
    g, = torch.autograd.grad(y, x, _nchw(dz, B, H // 2, W // 2, Cout))
    _w(dx, g.permute(0, 2, 3, 1))


def uad_conv2d_wgrad(x, dz, dw, B, H, W, Cin, Cout, k, accumulate, mm, ws, wsb, st):
    wt = torch.zeros(k, k, Cin, Cout, dtype=D, requires_grad=True)
    y = conv2d_same_s2(_nchw(x, B, H, W, Cin), wt, torch.zeros(Cout, dtype=D))
    g, = torch.autograd.grad(y, wt, _nchw(dz, B, H // 2, W // 2, Cout))
    _acc(dw, g, accumulate)


def uad_convT2d_fwd(x, w, bias, gamma, beta, z_out, a_out, B, H, W, Cin, Cout, k, act, alpha, bn_c, mm, ws, wsb, st):
    b = _v(bias, Cout) if bias is not None else torch.zeros(Cout, dtype=D)
    z = conv2dT_same_s2(_nchw(x, B, H, W, Cin), _v(w, k, k, Cout, Cin), b).permute(0, 2, 3, 1)
    _w(z_out, z)
    _w(a_out, _act(_affine(z, gamma, beta, Cout, bn_c), act, alpha))


def uad_convT2d_fwd_head_supported(B, H, W, Cin, Cout, k, mm):
    return 0          # tensor-core path only: the emulated engines take the unfused sequence (uad_convT2d_fwd + uad_final1x1_l1_fwd)


def uad_convT2d_fwd_head(x, w, bias, gamma, beta, a_out, head_w, head_b, head_out, B, H, W, Cin, Cout, k, act, alpha, bn_c, mm, ws, wsb, st):
    uad_convT2d_fwd(x, w, bias, gamma, beta, None, a_out, B, H, W, Cin, Cout, k, act, alpha, bn_c, mm, ws, wsb, st)
    _w(head_out, _v(a_out, B * 4 * H * W, Cout) @ _v(head_w, Cout) + _v(head_b, 1))


def uad_convT2d_dgrad(dz, w, dx, B, H, W, Cin, Cout, k, mm, ws, wsb, st):
    x = torch.zeros(B, Cin, H, W, dtype=D, requires_grad=True)
    y = conv2dT_same_s2(x, _v(w, k, k, Cout, Cin), torch.zeros(Cout, dtype=D))
    g, = torch.autograd.grad(y, x, _nchw(dz, B, 2 * H, 2 * W, Cout))
    _w(dx, g.permute(0, 2, 3, 1))


def uad_convT2d_wgrad(x, dz, dw, B, H, W, Cin, Cout, k, accumulate, mm, ws, wsb, st):
    wt = torch.zeros(k, k, Cout, Cin, dtype=D, requires_grad=True)
    y = conv2dT_same_s2(_nchw(x, B, H, W, Cin), wt, torch.zeros(Cout, dtype=D))
    g, = torch.autograd.grad(y, wt, _nchw(dz, B, 2 * H, 2 * W, Cout))
    _acc(dw, g, accumulate)


def uad_act_bn_bwd(da, z, gamma, beta, dz, dgamma, dbeta, dbias, rows, C, act, alpha, bn_c, accumulate, ws, wsb, st):
    zt, dat = _v(z, rows, C), _v(da, rows, C)
    g = _v(gamma, C) if gamma is not None else torch.ones(C, dtype=D)
    b = _v(beta, C) if gamma is not None else torch.zeros(C, dtype=D)
    c = bn_c if gamma is not None else 1.0
    if act & 0x100:                 # UAD_ACT_FROM_OUTPUT: `z` holds a = act(u), u = gamma*bn_c*z + beta (piecewise-linear act only)
        a = zt
        u = torch.where(a > 0, a, a / alpha) if (act & 0xff) == ACT_LEAKY else a
        zt = (u - b) / (g * c)
    else:
        u = g * c * zt + b
    du = dat * _dact(u, act, alpha)
    dzt = g * c * du
    if gamma is not None:
        _acc(dgamma, c * (du * zt).sum(0), accumulate)
        _acc(dbeta, du.sum(0), accumulate)
    _acc(dbias, dzt.sum(0), accumulate)
    _w(dz, dzt)


def uad_final1x1_l1_bwd_fused(z, gamma, beta, w, x, xhat, scale, dz, dgamma, dbeta, dbias_prev, dw, dbias, B, HW, Cin, act, alpha, bn_c,
                              accumulate, ws, wsb, st):
    """= uad_final1x1_l1_bwd followed by uad_act_bn_bwd of the preceding block, without materialising da."""
    rows = B * HW
    g, b = _v(gamma, Cin), _v(beta, Cin)
    zt = _v(z, rows, Cin)
    if act & 0x100:
        a = zt
        zt = ((torch.where(a > 0, a, a / alpha) if (act & 0xff) == ACT_LEAKY else a) - b) / (g * bn_c)
    else:
        a = _act(g * bn_c * zt + b, act, alpha)
    dxh = torch.sign(_v(xhat, rows) - _v(x, rows)) * scale
    _acc(dw, a.t() @ dxh, accumulate)
    _acc(dbias, dxh.sum().reshape(1), accumulate)
    da = dxh[:, None] * _v(w, Cin)[None, :]
    du = da * _dact(g * bn_c * zt + b, act, alpha)
    dzt = g * bn_c * du
    _acc(dgamma, bn_c * (du * zt).sum(0), accumulate)
    _acc(dbeta, du.sum(0), accumulate)
    _acc(dbias_prev, dzt.sum(0), accumulate)
    _w(dz, dzt)


def uad_mask_bn_act_fwd(x, mask, keep, gamma, beta, bn_c, act, alpha, z_out, a_out, rows, C, st):
    z = _v(x, rows, C)
    if mask is not None:
        z = z * _v(mask, rows, C) * keep
    _w(z_out, z)
    _w(a_out, _act(_affine(z, gamma, beta, C, bn_c), act, alpha))


def uad_mask_scale(g, mask, keep, out, n, st):
    _w(out, _v(g, n) * _v(mask, n) * keep)


def uad_l1_direct_term(x, xhat, scale, gx, n, st):
    _w(gx, _v(gx, n) - torch.sign(_v(xhat, n) - _v(x, n)) * scale)


def uad_mul_abs(l1, gx, out, n, st):
    _w(out, _v(l1, n) * _v(gx, n).abs())


def uad_loss_scalars(rec, kl, out3, B, st):
    r = _v(rec, B)
    k = _v(kl, B) if kl is not None else torch.zeros(B, dtype=D)
    _w(out3, torch.stack([r.mean(), k.mean(), (r + k).mean()]))


# ------------------------------------------------------------------------------------------------ dense / bottleneck
def uad_dense_fwd(x, w, bias, mask, mask_scale, gamma, beta, z_out, a_out, M, K, N, act, alpha, bn_c, ws, wsb, st):
    z = _v(x, M, K) @ _v(w, K, N)
    if bias is not None:
        z = z + _v(bias, N)
    if mask is not None:
        z = z * _v(mask, M, N) * mask_scale
    _w(z_out, z)
    _w(a_out, _act(_affine(z, gamma, beta, N, bn_c), act, alpha))


def uad_dense_bwd(x, w, dz, mask, mask_scale, dx, dw, dbias, M, K, N, accumulate, ws, wsb, st):
    d = _v(dz, M, N)
    if mask is not None:
        d = d * _v(mask, M, N) * mask_scale
    if dx is not None:
        _w(dx, d @ _v(w, K, N).t())
    _acc(dw, _v(x, M, K).t() @ d, accumulate)
    _acc(dbias, d.sum(0), accumulate)


def uad_reparam_kl_fwd(mu, ls, eps, sigma, z, kl, B, Z, st):
    m, l = _v(mu, B, Z), _v(ls, B, Z)
    s = torch.exp(l)
    _w(sigma, s)
    _w(z, m if eps is None else m + _v(eps, B, Z) * s)
    _w(kl, 0.5 * (m * m + s * s - torch.log(s * s) - 1).sum(1))


def uad_reparam_kl_bwd(mu, ls, eps, dz, kl_scale, dmu, dls, B, Z, st):
    m, l, d = _v(mu, B, Z), _v(ls, B, Z), _v(dz, B, Z)
    s = torch.exp(l)
    _w(dmu, d + kl_scale * m)
    _w(dls, d * _v(eps, B, Z) * s + kl_scale * (s * s - 1))


def uad_final1x1_l1_fwd(a, w, bias, x, xhat, l1, rec, B, HW, Cin, ws, wsb, st):
    y = _v(a, B * HW, Cin) @ _v(w, Cin) + _v(bias, 1)
    _w(xhat, y)
    if l1 is not None or rec is not None:
        r = (y - _v(x, B * HW)).abs()
        _w(l1, r)
        _w(rec, r.reshape(B, HW).sum(1))


def _final_bwd(a, w, dxh, da, dw, dbias, B, HW, Cin, accumulate):
    _w(da, dxh[:, None] * _v(w, Cin)[None, :])
    _acc(dw, _v(a, B * HW, Cin).t() @ dxh, accumulate)
    _acc(dbias, dxh.sum().reshape(1), accumulate)


def uad_final1x1_l1_bwd(a, w, x, xhat, scale, da, dw, dbias, B, HW, Cin, accumulate, ws, wsb, st):
    _final_bwd(a, w, torch.sign(_v(xhat, B * HW) - _v(x, B * HW)) * scale, da, dw, dbias, B, HW, Cin, accumulate)


def uad_final1x1_bwd(a, w, dxhat, da, dw, dbias, B, HW, Cin, accumulate, ws, wsb, st):
    _final_bwd(a, w, _v(dxhat, B * HW), da, dw, dbias, B, HW, Cin, accumulate)


# ------------------------------------------------------------------------------------------------ LayerNormalization([1,2])
def _ln_parts(x, stats, gamma, beta, B, HW, C):
    xt = _v(x, B, HW, C)
    st_ = _v(stats, 2, B, C)
    xh = (xt - st_[0][:, None, :]) * st_[1][:, None, :]
    g, b = _v(gamma, HW)[None, :, None], _v(beta, HW)[None, :, None]
    return xt, xh, st_[1][:, None, :], g, b


def uad_layernorm_hw_fwd_train(x, gamma, beta, y, stats, B, HW, C, eps, act, alpha, ws, wsb, st):
    xt = _v(x, B, HW, C)
    mean = xt.mean(1)
    rstd = 1.0 / torch.sqrt(xt.var(1, unbiased=False) + eps)
    _w(stats, torch.stack([mean, rstd]))
    n = (xt - mean[:, None, :]) * rstd[:, None, :] * _v(gamma, HW)[None, :, None] + _v(beta, HW)[None, :, None]
    _w(y, _act(n, act, alpha))


def uad_layernorm_hw_fwd(x, gamma, beta, y, B, HW, C, eps, act, alpha, ws, wsb, st):
    uad_layernorm_hw_fwd_train(x, gamma, beta, y, None, B, HW, C, eps, act, alpha, ws, wsb, st)


def _ln_reverse(dn, xh, r, g):
    """adjoint of x through xhat = (x - mean) * rstd given dn = adjoint of gamma*xhat+beta."""
    dxh = dn * g
    return r * (dxh - dxh.mean(1, keepdim=True) - xh * (dxh * xh).mean(1, keepdim=True))


def uad_layernorm_hw_bwd(dy, x, stats, gamma, beta, dx, dgamma, dbeta, B, HW, C, act, alpha, accumulate, ws, wsb, st):
    xt, xh, r, g, b = _ln_parts(x, stats, gamma, beta, B, HW, C)
    dn = _v(dy, B, HW, C) * _dact(g * xh + b, act, alpha)
    _acc(dgamma, (dn * xh).sum((0, 2)), accumulate)
    _acc(dbeta, dn.sum((0, 2)), accumulate)
    _w(dx, _ln_reverse(dn, xh, r, g))


def _ln_tangent(xd, xh, r):
    return r * (xd - xd.mean(1, keepdim=True) - xh * (xd * xh).mean(1, keepdim=True))


def uad_layernorm_hw_jvp(xdot, x, stats, gamma, beta, ydot, jstats, B, HW, C, act, alpha, ws, wsb, st):
    xt, xh, r, g, b = _ln_parts(x, stats, gamma, beta, B, HW, C)
    xd = _v(xdot, B, HW, C)
    _w(jstats, torch.stack([xd.mean(1), (xd * xh).mean(1)]))
    _w(ydot, _dact(g * xh + b, act, alpha) * g * _ln_tangent(xd, xh, r))


def uad_layernorm_hw_bwd2(dydot, dy, x, xdot, stats, jstats, gamma, beta, dxdot, dx, dgamma, dbeta, B, HW, C, act, alpha,
                          accumulate, ws, wsb, st):
    """Reverse of (y, ydot) = (act(LN(x)), d/de act(LN(x + e xdot))) by autograd on the closed forms (the activation's
    derivative is piecewise constant, so its own derivative contributes nothing - exactly what the kernel assumes)."""
    xt = _v(x, B, HW, C).clone().requires_grad_(True)
    xd = _v(xdot, B, HW, C).clone().requires_grad_(True)
    g0 = _v(gamma, HW).clone().requires_grad_(True)
    b0 = _v(beta, HW).clone().requires_grad_(True)
    # the forward's epsilon is not an argument: recover var + eps from the saved rstd so autograd sees rstd as a function of x
    rs = _v(stats, 2, B, C)[1][:, None, :]
    eps_t = (1.0 / (rs * rs) - _v(x, B, HW, C).var(1, unbiased=False, keepdim=True)).detach()
    mean = xt.mean(1, keepdim=True)
    r = 1.0 / torch.sqrt(xt.var(1, unbiased=False, keepdim=True) + eps_t)
    xh = (xt - mean) * r
    g, b = g0[None, :, None], b0[None, :, None]
    n = g * xh + b
    da = _dact(n.detach(), act, alpha)
    y = torch.where(n > 0, n, alpha * n) if (act & 0xff) == ACT_LEAKY else (torch.relu(n) if (act & 0xff) == ACT_RELU else n)
    ydot = da * g * _ln_tangent(xd, xh, r)
    total = (ydot * _v(dydot, B, HW, C)).sum()
    if dy is not None:
        total = total + (y * _v(dy, B, HW, C)).sum()
    gx, gxd, gg, gb = torch.autograd.grad(total, (xt, xd, g0, b0), allow_unused=True)
    _w(dxdot, gxd)
    _w(dx, gx)
    _acc(dgamma, gg, accumulate)
    _acc(dbeta, torch.zeros(HW, dtype=D) if gb is None else gb, accumulate)


# ------------------------------------------------------------------------------------------------ element-wise / reductions
def uad_activation(x, y, n, act, alpha, st):
    _w(y, _act(_v(x, n), act, alpha))


def uad_activation_bwd(dy, u, dx, n, act, alpha, st):
    _w(dx, _v(dy, n) * _dact(_v(u, n), act, alpha))


def uad_fill(y, v, n, st):
    _w(y, torch.full((n,), float(v), dtype=D))


def uad_axpby(a, x, b, y, n, st):
    _w(y, a * _v(x, n) + (b * _v(y, n) if b != 0 else 0.0))         # b == 0: y is write-only (the kernel does not read it)


def uad_sum_scaled(x, n, scale, out, ws, wsb, st):
    _w(out, (_v(x, n).sum() * scale).reshape(1))


def uad_mse(a, b, n, grad_scale, grad_a, loss_scale, loss_out, ws, wsb, st):
    d = _v(a, n) - _v(b, n)
    _w(grad_a, grad_scale * d)
    _w(loss_out, (loss_scale * (d * d).sum()).reshape(1))


def uad_gradient_penalty(ddx, B, H, WC, scale, u_out, gp_out, ws, wsb, st):
    t = _v(ddx, B, H, WC).clone().requires_grad_(True)
    gp = ((torch.sqrt((t * t).sum(1)) - 1.0) ** 2).mean() * scale
    u, = torch.autograd.grad(gp, t)
    _w(u_out, u)
    _w(gp_out, gp.detach().reshape(1))


def uad_interpolate(x, x_gen, alpha, out, B, per_sample, st):
    xt, xg = _v(x, B, per_sample), _v(x_gen, B, per_sample)
    _w(out, xt + _v(alpha, B)[:, None] * (xg - xt))


def uad_l1_map(x, xhat, l1, rec, B, HW, st):
    r = (_v(xhat, B, HW) - _v(x, B, HW)).abs()
    _w(l1, r)
    _w(rec, r.sum(1))


def uad_counter_add(counter, inc, st):
    c = _resolve(counter, 1, torch.int64)
    c[0] += int(inc)


def uad_adam_tf_step(params, grads, m, v, n, lr, b1, b2, eps, grad_scale, step_dev, st):
    lr_t = lr
    if step_dev is not None:
        t = int(_resolve(step_dev, 1, torch.int64)[0])
        lr_t = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    g = _v(grads, n) * grad_scale
    mn = b1 * _v(m, n) + (1 - b1) * g
    vn = b2 * _v(v, n) + (1 - b2) * g * g
    _w(m, mn)
    _w(v, vn)
    _w(params, _v(params, n) - lr_t * mn / (vn.sqrt() + eps))


# ------------------------------------------------------------------------------------------------ GMVAE latent block, restoration
def _gmvae_terms(z_mu, z_ls, z_s, M, S, dc, c_lambda):
    """The reference's graph nodes (models/gaussian_mixture_variational_autoencoder.py:64-71, trainers/GMVAE.py:66-88) - an
    implementation independent of csrc/uad_gmvae_latent.h (that header is checked against the same formulas in test_gmvae_latent.py)."""
    zs = z_s.unsqueeze(-1)
    pc = torch.softmax((-0.5 * (zs - M) ** 2 * torch.exp(S) - S + math.log(math.pi)).sum(1), dim=-1)
    kl = 0.5 * ((torch.exp(z_ls).unsqueeze(-1) + (z_mu.unsqueeze(-1) - M) ** 2) * (torch.exp(S) + 1e-6) - S - z_ls.unsqueeze(-1) - 1)
    con = (kl * pc.unsqueeze(1)).sum((1, 2))
    closs1 = (pc * torch.log(pc * dc + 1e-8)).sum(1)
    return pc, con, torch.maximum(closs1, torch.full_like(closs1, c_lambda))


def uad_gmvae_latent_fwd(z_mu, z_ls, z_s, M, S, pc, con, closs, B, dz, dc, c_lambda, st):
    p, c, cl = _gmvae_terms(_v(z_mu, B, dz), _v(z_ls, B, dz), _v(z_s, B, dz), _v(M, B, dz, dc), _v(S, B, dz, dc), dc, c_lambda)
    _w(pc, p)
    _w(con, c)
    _w(closs, cl)


def uad_gmvae_latent_bwd(z_mu, z_ls, z_s, M, S, scale, dz_mu, dz_ls, dz_s, dM, dS, B, dz, dc, c_lambda, st):
    t = [a.clone().requires_grad_(True) for a in (_v(z_mu, B, dz), _v(z_ls, B, dz), _v(z_s, B, dz), _v(M, B, dz, dc), _v(S, B, dz, dc))]
    _, c, cl = _gmvae_terms(*t, dc, c_lambda)
    for dst, g in zip((dz_mu, dz_ls, dz_s, dM, dS), torch.autograd.grad(scale * (c + cl).sum(), t)):
        _w(dst, g)


def uad_tv_restore_seed(x, xhat, tv_lambda, g, tv, B, H, W, ws, wsb, st):
    xt, xh = _v(x, B, H, W), _v(xhat, B, H, W)
    d = (xt - xh).clone().requires_grad_(True)
    t = (d[:, 1:] - d[:, :-1]).abs().sum((1, 2)) + (d[:, :, 1:] - d[:, :, :-1]).abs().sum((1, 2))
    T, = torch.autograd.grad(t.sum(), d)
    _w(g, torch.sign(xh - xt) - tv_lambda * T)
    _w(tv, t.detach())


def uad_restore_update(x, gx, g, lr, grads_out, n, st):
    gr = _v(gx, n) - _v(g, n)
    _w(grads_out, gr)
    _w(x, _v(x, n) - lr * gr)


# ------------------------------------------------------------------------------------------------ scoring / evaluation stencils
def uad_binary_erosion_cross(mask, out, N, H, W, iterations, st):
    import scipy.ndimage
    m = _resolve(mask, N * H * W, torch.uint8)[:N * H * W].reshape(N, H, W).numpy() != 0
    se = scipy.ndimage.generate_binary_structure(2, 1)
    res = np.stack([scipy.ndimage.binary_erosion(sl, structure=se, iterations=int(iterations)) for sl in m])
    _w(out, torch.from_numpy(res.astype(np.uint8)), torch.uint8)


def uad_residual_score(x, xhat, mask, prior_quantile, keep_positive, apply_prior, diff, n, st):
    xt = _resolve(x, n)[:n].clone()                         # float32 arithmetic, as the kernel (Evaluation.py:282-291)
    d = xt - _resolve(xhat, n)[:n]
    d = torch.clamp(d, min=0) if keep_positive else d.abs()
    if mask is not None:
        d = d * (_resolve(mask, n, torch.uint8)[:n] != 0).to(d.dtype)
    if apply_prior:
        d = torch.where(xt.to(D) < prior_quantile, torch.zeros_like(d), d)
    _w(diff, d)


def uad_median_filter3d_5(vol, out, Z, H, W, st):
    import scipy.ndimage
    v = _resolve(vol, Z * H * W)[:Z * H * W].reshape(Z, H, W).numpy()
    _w(out, torch.from_numpy(scipy.ndimage.median_filter(v, (5, 5, 5), mode='reflect')))


def uad_threshold_counts(diff, label, n, thresholds, n_thr, counts, mask_out, st):
    d = _resolve(diff, n)[:n].to(D)
    g = (_resolve(label, n, torch.uint8)[:n] != 0) if label is not None else torch.zeros(n, dtype=torch.bool)
    res = []
    for i in range(int(n_thr)):
        p = d > float(thresholds[i])
        res += [int((p & g).sum()), int(p.sum()), int(g.sum())]
        if i == 0 and mask_out is not None:
            _w(mask_out, p.to(torch.uint8), torch.uint8)
    _w(counts, torch.tensor(res, dtype=torch.int64), torch.int64)


_rng = np.random.default_rng(1234)


def uad_randn(out, n, seed, offset, offset_dev, st):
    _w(out, torch.from_numpy(_rng.standard_normal(n)))


def uad_uniform(out, n, seed, offset, offset_dev, st):
    _w(out, torch.from_numpy(_rng.random(n)))


def uad_dropout_mask(mask, n, rate, seed, offset, offset_dev, st):
    _w(mask, torch.from_numpy((_rng.random(n) >= rate).astype(np.float64)))


calls = []


def check_signature(name, args):
    """What ctypes would enforce on the GPU box: the argument COUNT and Python types of a call against abi.SIGNATURES (the
    emulator bypasses ctypes, so a float where the ABI takes an int, a numpy scalar ctypes rejects, or a missing argument would
    otherwise only surface on hardware)."""
    import ctypes as C

    from unsupervised_anomaly_detection_brain_mri_b200.abi import SIGNATURES
    assert name in SIGNATURES, f'{name} has no ctypes signature in abi.SIGNATURES'
    argtypes = SIGNATURES[name][1]
    assert len(args) == len(argtypes), f'{name}: {len(args)} arguments, the ABI takes {len(argtypes)}'
    for i, (a, t) in enumerate(zip(args, argtypes)):
        if t is C.c_void_p or (isinstance(t, type) and issubclass(t, C._Pointer)):
            ok = a is None or isinstance(a, torch.Tensor) or type(a) is int or isinstance(a, C.Array)
        elif t in (C.c_float, C.c_double):
            ok = type(a) in (int, float) or isinstance(a, float)
        else:                                                    # c_int, c_longlong, c_size_t, c_uint64
            ok = type(a) is int and (a >= 0 or t in (C.c_int, C.c_longlong))
        assert ok, f'{name}: argument {i} = {a!r} ({type(a).__name__}) does not fit {t.__name__}'


def call(name, *args):
    fn = globals().get(name)
    if fn is None:
        raise NotImplementedError(f'abi_emulator: {name} is not modelled')
    check_signature(name, args)
    calls.append(name)
    with torch.enable_grad():
        fn(*args)
    return 0


def ptr(t):
    return t


def install(monkeypatch, *modules):
    """Route the given engine modules' ABI calls through the emulator and make their engines live on the CPU."""
    from unsupervised_anomaly_detection_brain_mri_b200.engine import ConvAutoencoderEngine
    from unsupervised_anomaly_detection_brain_mri_b200.fanogan_engine import FanoganEngine
    for mod in modules:
        monkeypatch.setattr(mod, 'call', call)
        monkeypatch.setattr(mod, 'ptr', ptr)
    monkeypatch.setattr(FanoganEngine, '_st', lambda self: 0)
    monkeypatch.setattr(ConvAutoencoderEngine, '_st', lambda self: 0)
    del _registry[:]
    del calls[:]


def adopt(engine):
    """Register the buffers an engine passes as raw pointers (loss scalars, step counters, the Philox counter)."""
    for name in ('sc', 'scalars', 'rng_ctr', 'step_dev'):
        if hasattr(engine, name):
            register(getattr(engine, name))
    for name in ('steps', 'op_steps'):
        if hasattr(engine, name):
            register(*getattr(engine, name).values())
    return engine


def poison(engine):
    """Fill every float32 work buffer of an engine with NaN (parameters, gradient / Adam buffers, loss scalars and constant
    vectors excepted): any ABI call that READS a buffer no earlier call wrote then shows up as NaN in the results - on the GPU
    such a read would see whatever the allocator left there.  Call it after the engine is built, before inputs / noise are staged."""
    skip = {id(t) for t in (getattr(engine, 'scalars', None), getattr(engine, 'sc', None)) if t is not None}
    seen = 0

    def visit(o, name=''):
        nonlocal seen
        if isinstance(o, torch.Tensor):
            if o.dtype == torch.float32 and id(o) not in skip and 'ones' not in name:
                o.fill_(float('nan'))
                seen += 1
        elif isinstance(o, (list, tuple)):
            for v in o:
                visit(v, name)
        elif isinstance(o, dict):
            for k, v in o.items():
                visit(v, f'{name}.{k}')
        elif type(o).__name__ in ('_Branch', '_CriticPass'):
            for k, v in vars(o).items():
                visit(v, k)

    for k, v in vars(engine).items():
        if k in ('fp', 'm_gen', 'v_gen', 'ws'):
            continue
        visit(v, k)
    return seen
