"""ctypes loader of oracle/liboracle_gae.so (C restatement; test infrastructure)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, 'liboracle_gae.so')
_h = None


def build():
    r = subprocess.run(['make', '-C', _HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('oracle C build failed: ' + r.stderr)
    return _LIB


def lib():
    global _h
    if _h is None:
        if not os.path.exists(_LIB):
            build()
        _h = ctypes.CDLL(_LIB)
    return _h


def gae(rewards, values, dones, bootstrap, gamma, gae_lambda, want_returns=True, threads=None):
    T, N = rewards.shape[0], rewards[0].size
    r = np.ascontiguousarray(rewards, np.float32)
    v = np.ascontiguousarray(values, np.float32)
    d = np.ascontiguousarray(dones).astype(np.uint8)
    b = np.ascontiguousarray(bootstrap, np.float32)
    adv = np.empty_like(r)
    ret = np.empty_like(r) if want_returns else None
    P = ctypes.c_void_p
    fn = lib().oracle_gae_f32
    threads = threads or os.cpu_count() or 1
    chunk = max(1024, -(-N // threads) // 1024 * 1024 + 1024)

    def run(n0):
        fn(P(r.ctypes.data), P(v.ctypes.data), P(d.ctypes.data), P(b.ctypes.data),
           P(adv.ctypes.data), P(ret.ctypes.data if want_returns else 0), ctypes.c_int(T),
           ctypes.c_longlong(N), ctypes.c_float(gamma),
           ctypes.c_float(float(gamma) * float(gae_lambda)), ctypes.c_longlong(n0),
           ctypes.c_longlong(min(N, n0 + chunk)))
    starts = list(range(0, N, chunk))
    if len(starts) == 1:
        run(0)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(run, starts))
    return adv, ret
