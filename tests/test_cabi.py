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


def test_struct_layouts_match_header(tmp_path):
    """The ctypes mirrors against the header itself: gcc compiles include/radiobear_b200.h and prints sizeof / offsetof."""
    import subprocess
    fields = {'rb_alpha_desc': (_lib.AlphaDesc, ['n_layers', 'formalism', 'freqs', 'gas_col', 'cloud', 'cloud_flags', 'units',
                                                 'scale', 'freqs_host', 'freqs_per_layer']),
              'rb_geometry_desc': (_lib.GeometryDesc, ['radius', 'n0', 'orientation', 'gtype', 'limb']),
              'rb_rt_desc': (_lib.RtDesc, ['n_freqs', 'alpha', 'T', 'disc_average', 'out_f32', 'tau_cut', 'alpha0']),
              'rb_gravity_model': (_lib.GravityModel, ['radius', 'GM_layer', 'n_J', 'Jn', 'RJ', 'n_vw', 'vwdat', 'latstep', 'max_lat'])}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "radiobear_b200.h"', 'int main(void) {']
    for st, (_, names) in fields.items():
        src.append('  printf("{0} %zu\\n", sizeof({0}));'.format(st))
        for n in names:
            src.append('  printf("{0}.{1} %zu\\n", offsetof({0}, {1}));'.format(st, n))
    src += ['  return 0;', '}']
    c = tmp_path / 'layout.c'
    c.write_text('\n'.join(src))
    exe = str(tmp_path / 'layout')
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(c), '-o', exe], check=True)
    out = dict(line.split() for line in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines())
    for st, (cls, names) in fields.items():
        assert ctypes.sizeof(cls) == int(out[st]), st
        for n in names:
            assert getattr(cls, n).offset == int(out['{}.{}'.format(st, n)]), (st, n)
    assert _lib.AlphaDesc.freqs.offset == 48 and ctypes.sizeof(_lib.RtDesc) == 48


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


def test_integration_doc_structs_match_the_binding():
    """The ctypes structures INTEGRATION.md shows a maintainer are the ones radiobear_b200/_lib.py binds (which the
    layout test above checks against the header compiled with gcc): same field names, same order."""
    import re
    from radiobear_b200 import _lib
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    shown = {m.group(1): re.findall(r"\('(\w+)'", m.group(2))
             for m in re.finditer(r"class (\w+)\(C\.Structure\):.*?_fields_ = \[(.*?)\]\n", text, re.S)}
    assert shown, 'INTEGRATION.md lost its binding stub'
    for name, fields in shown.items():
        assert fields == [f[0] for f in getattr(_lib, name)._fields_], name
