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
