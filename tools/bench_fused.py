"""Micro-benchmark of the tensor-core layer kernels at the cfg2 minibatch shape (65536 x 256 x 256)."""
import os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import madrona_learn_b200 as m
from madrona_learn_b200._lib import c_int, call, ptr
from madrona_learn_b200.engine import gemm_tc

dev = 'cuda:0'
rows, H, K = int(os.environ.get('ROWS', 65536)), int(os.environ.get('H', 256)), int(os.environ.get('K', 256))
BF = torch.bfloat16
X = torch.randn(rows, K, device=dev).to(BF)
Wt = (torch.randn(H, K, device=dev) * 0.06).to(BF)
W = (torch.randn(K, H, device=dev) * 0.06).to(BF)      # [in=K(prev width), out=H]
s = torch.ones(H, device=dev); b = torch.zeros(H, device=dev)
Y = torch.empty(rows, H, device=dev, dtype=BF); XH = torch.empty_like(Y)
rstd = torch.empty(rows, device=dev)
DZ = torch.randn(rows, H, device=dev).to(BF); DZo = torch.empty(rows, K, device=dev, dtype=BF)
XHp = torch.randn(rows, K, device=dev).to(BF); rstdp = torch.rand(rows, device=dev) + 0.5
sp = torch.ones(K, device=dev); bp = torch.zeros(K, device=dev)
gs = torch.zeros(K, device=dev); gb = torch.zeros(K, device=dev)
gW = torch.zeros(K, H, device=dev)
flush = torch.zeros(64 << 20, device=dev)


def t(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.add_(1); torch.cuda.synchronize(); torch.cuda._sleep(300000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


fwd = lambda: call('mlb_dense_ln_relu_fwd_tc', ptr(X), ptr(Wt), ptr(s), ptr(b), ptr(Y), ptr(XH), ptr(rstd),
                   c_int(rows), c_int(K), c_int(H), c_int(K), c_int(K))
fwd_inf = lambda: call('mlb_dense_ln_relu_fwd_tc', ptr(X), ptr(Wt), ptr(s), ptr(b), ptr(Y), ptr(None), ptr(None),
                       c_int(rows), c_int(K), c_int(H), c_int(K), c_int(K))
dx = lambda: call('mlb_dense_dx_lnbwd_tc', ptr(DZ), ptr(W), ptr(sp), ptr(bp), ptr(XHp), ptr(rstdp), ptr(DZo),
                  ptr(gs), ptr(gb), c_int(rows), c_int(H), c_int(K), c_int(H), c_int(H))
splitk = max(1, min(148 // (-(-K // 128) * -(-H // 128)), rows // 256))
dw = lambda: gemm_tc(X, DZ, gW, None, K, H, rows, K, H, H, 1, 1, 2, splitk)
gb_f = rows * (K * 2 + 2 * H * 2) / 1e9
print(f'rows={rows} K={K} H={H}')
for name, fn, bytes_ in (('fwd(train)', fwd, rows * (K * 2 + 2 * H * 2)), ('fwd(infer)', fwd_inf, rows * (K * 2 + H * 2)),
                         ('dx_lnbwd', dx, rows * (H * 2 + 2 * K * 2)), ('dW splitk', dw, rows * (K + H) * 2)):
    us = t(fn)
    fl = 2.0 * rows * K * H
    print(f'{name:12s} {us:8.1f} us   {fl / us / 1e6:7.1f} TFLOP/s   {bytes_ / us / 1e3:7.0f} GB/s (algorithmic)')
