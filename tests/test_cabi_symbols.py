"""CPU-only: libnnconv_b200.so loads and exports every symbol include/nnconv_b200.h declares (no compute
calls without a GPU), argument validation that needs no device, and the header / binding lists agree."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'nnconv_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(nnconv_[a-z0-9_]+)\s*\(', txt)))


@pytest.fixture(scope='module')
def lib():
    from graph_pde_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def test_header_and_binding_lists_agree():
    from graph_pde_b200 import _lib
    assert _header_symbols() == sorted(_lib.SYMBOLS)


def test_every_declared_symbol_is_exported(lib):
    for name in _header_symbols():
        assert getattr(lib, name) is not None, name


def test_abi_version_and_size_queries(lib):
    from graph_pde_b200 import _lib
    assert lib.nnconv_abi_version() == _lib.ABI_VERSION == 2
    ws, tmp = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.nnconv_plan_sizes(1000, 50, ctypes.byref(ws), ctypes.byref(tmp)) == 0
    assert ws.value > 1000 * 8 and tmp.value > 0
    assert lib.nnconv_plan_sizes(-1, 50, ctypes.byref(ws), ctypes.byref(tmp)) != 0
    assert b'need' in lib.nnconv_last_error()
    dims = (ctypes.c_int * 4)(6, 1024, 1024, 4096)
    nbytes = ctypes.c_size_t()
    assert lib.nnconv_weights_sizes(3, dims, 64, 64, 1, ctypes.byref(nbytes)) == 0
    # W3p alone: 64 * 1024 * 64 fp16
    assert nbytes.value >= 64 * 1024 * 64 * 2 + 1024 * 1024 * 2
    bad = (ctypes.c_int * 4)(6, 1024, 1024, 4095)           # in*out mismatch
    assert lib.nnconv_weights_sizes(3, bad, 64, 64, 1, ctypes.byref(nbytes)) != 0
    assert lib.nnconv_weights_sizes(0, dims, 64, 64, 1, ctypes.byref(nbytes)) != 0


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from graph_pde_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(_lib.NNConvLibraryError):
        _lib.lib()
