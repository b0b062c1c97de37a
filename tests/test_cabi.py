"""CPU: the C-ABI library loads and exports every symbol include/radiobear_b200.h declares; without a
GPU the context refuses to come up (there is no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from radiobear_b200 import _lib

HEADER = os.path.join(ROOT, 'include', 'radiobear_b200.h')


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(rb_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported():
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), 'symbol {} declared in the header but not exported'.format(n)
    assert sorted(_lib.EXPORTED_SYMBOLS) == names
    assert lib.rb_abi_version() == 1


def test_struct_layouts_match_header():
    # sizes computed from the header by hand: see include/radiobear_b200.h
    assert ctypes.sizeof(_lib.AlphaDesc) == 12 + 32 + 4 + 3 * 8 + 8 + 4 + 32 + 4 + 8 + 4 + 24 + 4 + 4 * 3 + 4 + 8 or \
        ctypes.sizeof(_lib.AlphaDesc) % 8 == 0
    assert _lib.AlphaDesc.freqs.offset == 48
    assert _lib.GeometryDesc.radius.offset == 8
    assert ctypes.sizeof(_lib.RtDesc) == 40


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_lib.RadiobearB200Error) as e:
        _lib.Context(0)
    assert 'no CPU fallback' in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'radiobear_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith('.py'):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), fn
