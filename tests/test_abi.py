"""CPU: the C-ABI library loads and exports every symbol include/uad_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'uad_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(uad_[a-zA-Z0-9_]+)\s*\(', src)))


def test_library_built_and_exports_every_declared_symbol():
    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    if not os.path.exists(abi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    names = header_functions()
    assert len(names) >= 25
    L = ctypes.CDLL(abi.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f'{n} declared in include/uad_b200.h but not exported by libuad_b200.so'
    assert set(names) == set(abi.SIGNATURES), set(names) ^ set(abi.SIGNATURES)


def test_python_binding_loads_and_reports_version():
    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    L = abi.lib()
    assert L.uad_abi_version() == 1
    assert L.uad_launch_count() >= 0
    # host-only helpers are callable without a GPU
    assert L.uad_conv_workspace_bytes(abi.OP_CONV_WGRAD, 64, 256, 256, 1, 32, 5, 0) > 0
    assert L.uad_rowreduce_workspace_bytes(1 << 20, 32) > 0


def test_missing_library_fails_loudly(monkeypatch):
    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    monkeypatch.setattr(abi, '_lib', None)
    monkeypatch.setattr(abi, 'LIB_PATH', '/nonexistent/libuad_b200.so')
    import pytest
    with pytest.raises(abi.UadError):
        abi.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'unsupervised_anomaly_detection_brain_mri_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), os.path.join(dp, f)
    for f in ('run.py', 'mains/main_AE.py', 'mains/main_VAE.py', 'mains/main_ceVAE.py'):
        assert not re.search(r'^\s*(from|import)\s+oracle\b', open(os.path.join(ROOT, f)).read(), flags=re.M)


def _ws_bytes_in_subprocess(env_extra):
    """uad_conv_workspace_bytes is host-only code: query it in a fresh process (the developer switches are read once)."""
    import json
    import subprocess
    import sys
    code = ("import json, sys; sys.path.insert(0, %r)\n"
            "from unsupervised_anomaly_detection_brain_mri_b200 import abi\n"
            "L = abi.lib()\n"
            "cases = [(0, 64, 64, 64, 128), (1, 64, 64, 128, 128), (3, 64, 16, 128, 128), (4, 64, 32, 128, 64), (0, 64, 128, 32, 64),\n"
            "         (3, 16, 64, 64, 32), (2, 64, 128, 32, 64), (5, 64, 128, 32, 32), (2, 64, 32, 64, 128)]\n"
            "print(json.dumps([int(L.uad_conv_workspace_bytes(op, B, H, H, ci, co, 5, 1)) for op, B, H, ci, co in cases]))\n") % ROOT
    env = {k: v for k, v in os.environ.items() if not k.startswith('UAD_')}
    env.update(env_extra)
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, check=True).stdout
    return json.loads(out.strip().splitlines()[-1])


def test_workspace_query_is_independent_of_the_developer_switches():
    """The workspace a caller allocates covers every kernel a developer switch can select for the same op (UAD_HS=0 /
    UAD_WGRAD_SS=0 fall back to the converter-warp kernels), so the sizes do not depend on the switches."""
    base = _ws_bytes_in_subprocess({})
    assert base == _ws_bytes_in_subprocess({'UAD_HS': '0', 'UAD_WGRAD_SS': '0'})
    assert all(b > 0 for b in base)


def test_unbuilt_math_modes_are_rejected_not_aliased():
    """A math mode that is not built (bf16 storage = 3, or any other value) is an error at every conv entry point - host-side
    check, no GPU needed.  The three built modes are 0 (fp32 SIMT), 1 (3xTF32), 2 (1xTF32)."""
    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    L = abi.lib()
    rc = L.uad_conv2d_fwd(None, None, None, None, None, None, None, 1, 16, 16, 32, 32, 5, 0, 0.0, 1.0, 3, None, 0, None)
    assert rc != 0 and b'math_mode 3 is not built' in L.uad_last_error()
    rc = L.uad_convT2d_wgrad(None, None, None, 1, 16, 16, 32, 32, 5, 0, 7, None, 0, None)
    assert rc != 0 and b'math_mode 7 is not built' in L.uad_last_error()


def test_peer_optimizer_entry_points_validate_on_the_host():
    """csrc/uad_peer.cu: region sizing and the argument checks of uad_peer_adam_step run before any device work - no GPU needed."""
    import ctypes as C
    from unsupervised_anomaly_detection_brain_mri_b200 import abi
    L = abi.lib()
    n = 2194176                                            # a flat buffer of 64-float slots
    half = (n * 4 + 255) & ~255
    assert L.uad_peer_region_bytes(n) == 2 * half + 64 * 8
    regions = (C.c_void_p * 16)()
    assert L.uad_peer_adam_step(regions, 0, 17, n, 0, n, None, None, 1e-4, 0.5, 0.999, 1e-8, 1.0, None, None) != 0
    assert b'bad rank / world' in L.uad_last_error()
    assert L.uad_peer_adam_step(regions, 2, 2, n, 0, n, None, None, 1e-4, 0.5, 0.999, 1e-8, 1.0, None, None) != 0
    dummy = C.c_void_p(0x1000)
    assert L.uad_peer_adam_step(regions, 0, 2, n, 0, n, dummy, dummy, 1e-4, 0.5, 0.999, 1e-8, 0.5, None, None) != 0
    assert b'is not mapped' in L.uad_last_error()
    regions[0] = regions[1] = 0x10000
    assert L.uad_peer_adam_step(regions, 0, 2, n, 64, n, dummy, dummy, 1e-4, 0.5, 0.999, 1e-8, 0.5, None, None) != 0
    assert b'must lie inside the flat buffer' in L.uad_last_error()
