"""Micro-benchmark of the weight-streaming fused-layer kernels (HN = 512): device time per launch at the
cfg3 shapes, with the kernel's own cycle counters (experiment build only)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import madrona_learn_b200 as m
from madrona_learn_b200 import _lib
from madrona_learn_b200._lib import c_int, call, ptr

dev = torch.device('cuda', 0)
BF = torch.bfloat16
L = _lib.lib()
has_dbg = hasattr(L, 'mlb_stream_dbg')


def dbg(v):
    out = (ctypes.c_longlong * 8)()
    if has_dbg:
        L.mlb_stream_dbg(ctypes.c_int(v), out)
    return list(out)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dbg(-1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, dbg(-1)


def main():
    H = 512
    shapes = ((65536, 512), (262144, 512), (65536, 64))
    if len(sys.argv) > 2:
        shapes = ((int(sys.argv[1]), int(sys.argv[2])),)
    for M, K in shapes:
        x = torch.randn(M, K, device=dev).to(BF)
        wt = (torch.randn(H, K, device=dev) / K ** 0.5).to(BF)
        s = torch.ones(H, device=dev)
        b = torch.zeros(H, device=dev)
        y = torch.empty(M, H, device=dev, dtype=BF)
        xh = torch.empty(M, H, device=dev, dtype=BF)
        rstd = torch.empty(M, device=dev)
        dz = torch.empty(M, H, device=dev, dtype=BF)
        ds = torch.zeros(H, device=dev)
        db = torch.zeros(H, device=dev)
        dzin = torch.randn(M, K, device=dev).to(BF)
        fwd = lambda xh_=xh: call('mlb_dense_ln_relu_fwd_tc', ptr(x), ptr(wt), ptr(s), ptr(b), ptr(y), ptr(xh_),
                                  ptr(rstd), c_int(M), c_int(K), c_int(H), c_int(K), c_int(K))
        bwd = lambda: call('mlb_dense_dx_lnbwd_tc', ptr(dzin), ptr(wt), ptr(s), ptr(b), ptr(xh), ptr(rstd), ptr(dz),
                           ptr(ds), ptr(db), c_int(M), c_int(K), c_int(H), c_int(K), c_int(K))
        for mode in ((0, 1) if has_dbg else (0,)):
            dbg(mode)
            for name, fn in (('fwd', fwd), ('fwd_noxh', lambda: fwd(None)), ('dx', bwd)):
                us, prof = timeit(fn)
                tiles = max(prof[3], 1)
                print(json.dumps(dict(kernel=name, M=M, K=K, dbg=mode, us=round(us, 1),
                                      cta0_tiles=prof[3] // 10, us_per_tile=round(prof[2] / tiles / 1.9e3, 2),
                                      acc_wait_us_per_tile=round(prof[0] / tiles / 1.9e3, 2),
                                      panel_wait_us_per_tile=round(prof[1] / tiles / 1.9e3, 2))), flush=True)
        dbg(0)


main()
