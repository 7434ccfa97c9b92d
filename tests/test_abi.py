"""The C-ABI library builds, loads and exports every symbol include/gennet_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'gennet_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gn_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from gennet_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) > 40
    for s in syms:
        assert hasattr(lib, s), 'libgennet_b200.so does not export %s' % s


def test_binding_covers_header():
    from gennet_b200 import _lib
    assert set(header_symbols()) == set(_lib.exported_symbols())


def test_version_and_error_channel():
    from gennet_b200 import _lib
    lib = _lib.load()
    assert lib.gn_version() == 100
    # argument validation happens before any CUDA call, so it can be exercised without a GPU
    rc = lib.gn_fft_plan_create(1000, ctypes.byref(ctypes.c_void_p()))
    assert rc == -1 and b'power of two' in lib.gn_last_error()
    rc = lib.gn_dense_fwd_f32(None, None, None, None, 1, 1, 1, 0, 0.0, None)
    assert rc == -1 and b'null pointer' in lib.gn_last_error()
    rc = lib.gn_kde2d_pdf_f32(None, 4, None, 4, 1.0, 0.0, 1.0, 1.0, None, None)
    assert rc == -1 and b'null pointer' in lib.gn_last_error()
    rc = lib.gn_overlap_sums_f32(None, None, 10, None, None)
    assert rc == -1 and b'null pointer' in lib.gn_last_error()
    # split-operand tensor-core entry points: scaled fp16 pairs need both scale scalars; channel counts must tile
    one = ctypes.c_void_p(16)       # never dereferenced: validation fails first
    rc = lib.gn_split_f32_f16x2(one, one, None, 0, 64, None)
    assert rc == -1 and b'null pointer' in lib.gn_last_error()
    rc = lib.gn_conv1d_fwd_f16x2(one, None, one, one, None, one, None, 1, 64, 64, 60, 64, 5, 1, 0, 0, 0.0, None)
    assert rc == -1 and b'null pointer' in lib.gn_last_error()
    rc = lib.gn_conv1d_fwd_f16x2(one, one, one, one, None, one, None, 1, 64, 50, 60, 64, 5, 1, 0, 0, 0.0, None)
    assert rc == -1 and b'multiples of 64' in lib.gn_last_error()
    rc = lib.gn_conv1d_wgrad_f16x2(one, one, one, one, None, one, None, 1, 64, 64, 60, 64, 5, 1, 0, None)
    assert rc == -1 and b'Cin % 128' in lib.gn_last_error()
    rc = lib.gn_dense_dgrad_f16x2(one, one, one, one, None, one, None, 8, 100, 64, 0, 0.0, None)
    assert rc == -1 and b'K % 64' in lib.gn_last_error()


def test_no_cpu_fallback_without_device():
    import pytest
    import torch
    from gennet_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_lib.GennetError):
        _lib.require_device()
    with pytest.raises(_lib.GennetError):
        _lib.ptr(torch.zeros(4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'gennet_b200')
    n = 0
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), os.path.join(d, f)
                n += 1
    assert n >= 15
