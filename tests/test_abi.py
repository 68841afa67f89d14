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


def test_tensorboard_tag_set(mlb):
    """TrainingMetrics.tensorboard_log writes the reference's tag set (ml/metrics.py:218-244): for every
    ring slot and metric `p0/<name> Mean|σ|Min|Max` at step base_update_idx + slot (host-side logic)."""
    import ctypes
    import numpy as np
    import torch
    from madrona_learn_b200 import _lib
    from madrona_learn_b200.metrics import REC, TrainingMetrics
    tm = TrainingMetrics(['Loss', 'Rewards'], buffer_size=2, start_update_idx=0, device='cpu')
    recs = [_lib.Metric(1.5, 8.0, -1.0, 4.0, 2), _lib.Metric(0.25, 0.0, 0.25, 0.25, 1),
            _lib.Metric(2.5, 18.0, 0.0, 5.0, 2), _lib.Metric(0.5, 0.0, 0.5, 0.5, 1)]
    raw = np.frombuffer(b''.join(bytes(r) for r in recs), dtype=np.uint8).copy()
    assert raw.size == 4 * REC == 4 * ctypes.sizeof(_lib.Metric)
    tm.ring.copy_(torch.from_numpy(raw))
    out = []

    class W:
        def scalar(self, tag, value, step):
            out.append((tag, float(value), step))
    tm.tensorboard_log(7, W())
    tags = [t for t, _, _ in out]
    assert tags[:4] == ['p0/Loss Mean', 'p0/Loss σ', 'p0/Loss Min', 'p0/Loss Max']
    assert len(out) == 2 * 2 * 4 and {s for _, _, s in out} == {7, 8}
    d = {(t, s): v for t, v, s in out}
    assert d[('p0/Loss Mean', 7)] == 1.5 and d[('p0/Loss σ', 7)] == 2.0 and d[('p0/Loss Max', 8)] == 5.0
    assert d[('p0/Rewards Min', 8)] == 0.5


def test_ffi_shim_type_checks(tmp_path):
    """csrc/ffi_xla.cc (the jax.ffi handlers a reference maintainer builds against jaxlib) cannot be built here:
    jaxlib's xla/ffi/api/ffi.h is not in this image.  g++ type-checks it against tests/ffi_mock (a stand-in for
    the slice of that API the shim uses): every call into include/mlb200.h is checked against the C prototypes and
    every binding's .Ctx/.Arg/.Ret/.Attr list against its implementation's parameters.  A deliberately wrong
    binding must FAIL the same check (the mock has teeth), and every handler wraps a declared entry point."""
    import re
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which('g++') is None:
        pytest.skip('g++ not available')
    cuda_inc = '/usr/local/cuda/include'
    if not os.path.exists(os.path.join(cuda_inc, 'cuda_runtime.h')):
        pytest.skip('CUDA headers not available')
    shim = os.path.join(root, 'madrona-learn_b200', 'csrc', 'ffi_xla.cc')
    base = ['g++', '-std=c++17', '-fsyntax-only', '-Wall', '-Werror', '-I', os.path.join(root, 'tests', 'ffi_mock'),
            '-I', cuda_inc]
    r = subprocess.run(base + [shim], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    src = open(shim).read()
    handlers = re.findall(r'XLA_FFI_DEFINE_HANDLER_SYMBOL\(\s*(\w+)_ffi,', src)
    assert len(handlers) >= 14
    header = open(os.path.join(root, 'include', 'mlb200.h')).read()
    for h in handlers:
        assert re.search(r'\b%s\(' % h, header), f'{h}_ffi wraps an entry point that include/mlb200.h does not declare'
    # negative control: drop one .Arg from a binding -> the static_assert in the mock's To() must fire
    bad = src.replace('.Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>(),',
                      '.Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>(),', 1)
    assert bad != src
    bad = bad.replace('#include "../../include/mlb200.h"', '#include "%s"' % os.path.join(root, 'include', 'mlb200.h'))
    f = tmp_path / 'bad_ffi.cc'
    f.write_text(bad)
    r = subprocess.run(base + [str(f)], capture_output=True, text=True)
    assert r.returncode != 0 and 'disagree on the parameter list' in r.stderr
