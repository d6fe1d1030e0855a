"""ctypes binding of libuad_b200.so (the C ABI declared in include/uad_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, a ``UadError`` is raised.
PyTorch tensors are used only as device-memory containers - every compute call below goes to hand-written sm_100a
kernels through plain pointers.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libuad_b200.so')

ACT_NONE, ACT_LEAKY, ACT_RELU, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3, 4
ACT_FROM_OUTPUT = 0x100   # UAD_ACT_FROM_OUTPUT: backward kernels read the block's output a instead of its pre-BN input z
OP_CONV_FWD, OP_CONV_DGRAD, OP_CONV_WGRAD, OP_CONVT_FWD, OP_CONVT_DGRAD, OP_CONVT_WGRAD = range(6)
MATH_FP32_SIMT, MATH_TC_3XTF32, MATH_TC_1XTF32 = 0, 1, 2

_P, _I, _F, _D, _Z, _U64, _LL = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t, C.c_uint64, C.c_longlong

# name -> (restype, argtypes)   -- must list EVERY symbol of include/uad_b200.h (tests/test_abi.py checks this)
SIGNATURES = {
    'uad_last_error': (C.c_char_p, []),
    'uad_abi_version': (_I, []),
    'uad_launch_count': (_LL, []),
    'uad_conv_tc_supported': (_I, [_I] * 7),
    'uad_conv_workspace_bytes': (_Z, [_I] * 8),
    'uad_conv2d_fwd': (_I, [_P] * 7 + [_I] * 7 + [_F, _F, _I, _P, _Z, _P]),
    'uad_conv2d_dgrad': (_I, [_P] * 3 + [_I] * 7 + [_P, _Z, _P]),
    'uad_conv2d_wgrad': (_I, [_P] * 3 + [_I] * 8 + [_P, _Z, _P]),
    'uad_convT2d_fwd': (_I, [_P] * 7 + [_I] * 7 + [_F, _F, _I, _P, _Z, _P]),
    'uad_convT2d_fwd_head_supported': (_I, [_I] * 7),
    'uad_convT2d_fwd_head': (_I, [_P] * 9 + [_I] * 7 + [_F, _F, _I, _P, _Z, _P]),
    'uad_convT2d_dgrad': (_I, [_P] * 3 + [_I] * 7 + [_P, _Z, _P]),
    'uad_convT2d_wgrad': (_I, [_P] * 3 + [_I] * 8 + [_P, _Z, _P]),
    'uad_rowreduce_workspace_bytes': (_Z, [_LL, _I]),
    'uad_act_bn_bwd': (_I, [_P] * 8 + [_LL, _I, _I, _F, _F, _I, _P, _Z, _P]),
    'uad_dense_fwd': (_I, [_P] * 4 + [_F] + [_P] * 4 + [_I] * 4 + [_F, _F, _P, _Z, _P]),
    'uad_dense_workspace_bytes': (_Z, [_I, _I, _I]),
    'uad_dense_bwd': (_I, [_P] * 4 + [_F] + [_P] * 3 + [_I] * 4 + [_P, _Z, _P]),
    'uad_reparam_kl_fwd': (_I, [_P] * 6 + [_I, _I, _P]),
    'uad_reparam_kl_bwd': (_I, [_P] * 4 + [_F] + [_P] * 2 + [_I, _I, _P]),
    'uad_final1x1_l1_fwd': (_I, [_P] * 7 + [_I] * 3 + [_P, _Z, _P]),
    'uad_final1x1_l1_bwd': (_I, [_P] * 4 + [_F] + [_P] * 3 + [_I] * 4 + [_P, _Z, _P]),
    'uad_final1x1_l1_bwd_fused': (_I, [_P] * 6 + [_F] + [_P] * 6 + [_I] * 4 + [_F, _F, _I, _P, _Z, _P]),
    'uad_loss_scalars': (_I, [_P] * 3 + [_I, _P]),
    'uad_adam_tf_step': (_I, [_P] * 4 + [_Z] + [_F] * 5 + [_P, _P]),
    'uad_randn': (_I, [_P, _Z, _U64, _U64, _P, _P]),
    'uad_dropout_mask': (_I, [_P, _Z, _F, _U64, _U64, _P, _P]),
    'uad_counter_add': (_I, [_P, _U64, _P]),
    'uad_layernorm_hw_workspace_bytes': (_Z, [_I, _I, _I]),
    'uad_layernorm_hw_fwd': (_I, [_P] * 4 + [_I] * 3 + [_F, _I, _F, _P, _Z, _P]),
    'uad_activation': (_I, [_P, _P, _Z, _I, _F, _P]),
    'uad_residual_score': (_I, [_P] * 3 + [_D, _I, _I, _P, _Z, _P]),
    'uad_threshold_counts': (_I, [_P, _P, _Z, C.POINTER(C.c_double), _I, _P, _P, _P]),
    'uad_mul_abs': (_I, [_P] * 3 + [_Z, _P]),
    'uad_l1_direct_term': (_I, [_P, _P, _F, _P, _Z, _P]),
    'uad_debug_trace': (_I, [C.POINTER(C.c_longlong)]),
    'uad_axpby': (_I, [_F, _P, _F, _P, _Z, _P]),
    'uad_layernorm_hw_train_workspace_bytes': (_Z, [_I, _I, _I]),
    'uad_layernorm_hw_fwd_train': (_I, [_P] * 5 + [_I] * 3 + [_F, _I, _F, _P, _Z, _P]),
    'uad_layernorm_hw_bwd': (_I, [_P] * 8 + [_I] * 4 + [_F, _I, _P, _Z, _P]),
    'uad_layernorm_hw_jvp': (_I, [_P] * 7 + [_I] * 4 + [_F, _P, _Z, _P]),
    'uad_layernorm_hw_bwd2': (_I, [_P] * 12 + [_I] * 4 + [_F, _I, _P, _Z, _P]),
    'uad_final1x1_bwd': (_I, [_P] * 6 + [_I] * 4 + [_P, _Z, _P]),
    'uad_activation_bwd': (_I, [_P] * 3 + [_Z, _I, _F, _P]),
    'uad_fill': (_I, [_P, _F, _Z, _P]),
    'uad_uniform': (_I, [_P, _Z, _U64, _U64, _P, _P]),
    'uad_interpolate': (_I, [_P] * 4 + [_I, _Z, _P]),
    'uad_reduce_workspace_bytes': (_Z, []),
    'uad_sum_scaled': (_I, [_P, _Z, _D, _P, _P, _Z, _P]),
    'uad_mse': (_I, [_P, _P, _Z, _F, _P, _D, _P, _P, _Z, _P]),
    'uad_gradient_penalty': (_I, [_P] + [_I] * 3 + [_F, _P, _P, _P, _Z, _P]),
    'uad_l1_map': (_I, [_P] * 4 + [_I, _I, _P]),
    'uad_binary_erosion_cross': (_I, [_P, _P, _I, _I, _I, _I, _P]),
    'uad_median_filter3d_5': (_I, [_P, _P, _I, _I, _I, _P]),
    'uad_mask_bn_act_fwd': (_I, [_P, _P, _F, _P, _P, _F, _I, _F, _P, _P, _LL, _I, _P]),
    'uad_mask_scale': (_I, [_P, _P, _F, _P, _Z, _P]),
    'uad_tv_restore_workspace_bytes': (_Z, [_I, _I, _I]),
    'uad_tv_restore_seed': (_I, [_P, _P, _F, _P, _P, _I, _I, _I, _P, _Z, _P]),
    'uad_restore_update': (_I, [_P, _P, _P, _F, _P, _Z, _P]),
    'uad_gmvae_latent_fwd': (_I, [_P] * 8 + [_I, _I, _I, _F, _P]),
    'uad_gmvae_latent_bwd': (_I, [_P] * 5 + [_F] + [_P] * 5 + [_I, _I, _I, _F, _P]),
    'uad_peer_region_bytes': (_Z, [_Z]),
    'uad_peer_alloc': (_I, [_Z, C.POINTER(C.c_void_p)]),
    'uad_peer_free': (_I, [_P]),
    'uad_peer_ipc_handle': (_I, [_P, _P]),
    'uad_peer_ipc_open': (_I, [_P, C.POINTER(C.c_void_p)]),
    'uad_peer_ipc_close': (_I, [_P]),
    'uad_peer_adam_step': (_I, [C.POINTER(C.c_void_p), _I, _I, _Z, _Z, _Z, _P, _P] + [_F] * 5 + [_P, _P]),
}


class UadError(RuntimeError):
    pass


_lib = None
launches = 0   # number of ABI compute calls issued (bench.py reports kernels launched from this counter)


def lib():
    """Load libuad_b200.so once.  Raises UadError if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UadError(f'{LIB_PATH} not found - build it with `python -c "import __graft_entry__ as g; g.build()"` '
                           f'(there is no CPU fallback for the hot path)')
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.uad_abi_version() != 1:
            raise UadError('libuad_b200.so ABI version mismatch')
        _lib = L
    return _lib


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """Invoke an int-returning ABI function, raising UadError(uad_last_error()) on failure."""
    global launches
    L = lib()
    rc = getattr(L, name)(*args)
    launches += 1
    if rc != 0:
        raise UadError(f'{name}: {L.uad_last_error().decode()}')
    return rc
