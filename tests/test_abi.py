"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/mlbp.h declares,
and the ctypes signatures in _lib.py agree with the header prototypes.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

from macaronicusermodeling_b200 import _lib, build

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, 'include', 'mlbp.h')


def header_prototypes():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'\b(int|void|const char \*)\s*(mlbp_\w+)\s*\(([^)]*)\)\s*;', src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        codes = ''
        for a in [x.strip() for x in args.split(',')]:
            if a in ('void', ''):
                continue
            if '*' in a:
                codes += 'p'
            elif 'int64_t' in a:
                codes += 'l'
            elif 'float' in a:
                codes += 'f'
            elif re.match(r'(const\s+)?int\b', a):
                codes += 'i'
            else:
                raise AssertionError('unparsed argument %r of %s' % (a, name))
        protos[name] = codes
    return protos


@pytest.fixture(scope='module')
def lib():
    build.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_declares_what_binding_uses():
    protos = header_prototypes()
    for name, codes in _lib._SIGNATURES.items():
        assert name in protos, name
        assert protos[name] == codes, (name, protos[name], codes)


def test_library_exports_every_declared_symbol(lib):
    for name in header_prototypes():
        assert hasattr(lib, name), 'libmlbp.so does not export %s' % name


def test_version_and_no_device_probe(lib):
    lib.mlbp_version.restype = ctypes.c_int
    assert lib.mlbp_version() >= 100
    lib.mlbp_device_ok.restype = ctypes.c_int
    assert lib.mlbp_device_ok() in (0, 1)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_lib.MlbpError):
        _lib.require_device()
