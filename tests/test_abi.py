"""CPU: the C-ABI library builds, loads, and exports every symbol include/mlb200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'mlb200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(mlb_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported(mlb):
    from madrona_learn_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    h = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(h, n)]
    assert not missing, missing


def test_binding_covers_header(mlb):
    from madrona_learn_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    assert _lib.lib().mlb_abi_version() == _lib.ABI_VERSION


def test_no_cpu_fallback(mlb):
    """Host tensors are rejected loudly instead of being computed on the CPU."""
    import torch
    from madrona_learn_b200 import _lib
    with pytest.raises(_lib.MLBError):
        _lib.ptr(torch.zeros(4))


def test_header_is_plain_c():
    """The drop-in boundary must be consumable from C (cgo / JNI / ctypes style bindings): the
    header compiles as C99 with nothing but <stddef.h>/<stdint.h>."""
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'mlb200.h')
    r = subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-fsyntax-only', '-x', 'c', hdr],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
