"""The C-ABI library loads without a GPU and exports every symbol include/bsbolt_b200.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'bsbolt_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(bsb_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(built):
    lib = ctypes.CDLL(os.path.join(ROOT, 'bsbolt_b200', 'libbsbolt_b200.so'))
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/bsbolt_b200.h but not exported'


def test_python_binding_lists_the_same_symbols(built):
    from bsbolt_b200 import _native
    assert sorted(_native.EXPORTS) == declared_symbols()


def test_no_cpu_fallback(built, golden):
    """Without a CUDA device the product refuses to load an index instead of computing on the CPU."""
    from bsbolt_b200 import _native
    if _native.lib().bsb_device_count() > 0:
        pytest.skip('a CUDA device is present')
    with pytest.raises(RuntimeError, match='no CPU fallback|CUDA'):
        _native.Index(golden.idxbase, 0)
