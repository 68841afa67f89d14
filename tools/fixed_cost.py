"""Fixed-cost probe for the persistent tensor-core kernels: time vs number of tiles (diagnostic)."""
import os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import madrona_learn_b200 as m
from madrona_learn_b200._lib import c_int, call, ptr

dev = 'cuda:0'
BF = torch.bfloat16
H = 256
flush = torch.zeros(64 << 20, device=dev)
tiny = torch.zeros(1024, device=dev)


def t(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.add_(1); torch.cuda.synchronize(); torch.cuda._sleep(300000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


print('tiny torch kernel: %.1f us' % t(lambda: tiny.add_(1)))
for K in (64, 256):
    for rows in (128, 148 * 128, 2 * 148 * 128, 4 * 148 * 128, 8 * 148 * 128):
        X = torch.randn(rows, K, device=dev).to(BF)
        Wt = (torch.randn(H, K, device=dev) * 0.06).to(BF)
        s = torch.ones(H, device=dev); b = torch.zeros(H, device=dev)
        Y = torch.empty(rows, H, device=dev, dtype=BF); XH = torch.empty_like(Y)
        rstd = torch.empty(rows, device=dev)
        fwd = lambda: call('mlb_dense_ln_relu_fwd_tc', ptr(X), ptr(Wt), ptr(s), ptr(b), ptr(Y), ptr(XH), ptr(rstd),
                           c_int(rows), c_int(K), c_int(H), c_int(K), c_int(K))
        print(f'K={K} rows={rows} tiles/SM={rows / 128 / 148:.2f}: {t(fwd):.1f} us')
K = 256
for rows in (128, 148 * 128, 2 * 148 * 128, 4 * 148 * 128, 8 * 148 * 128):
    DZ = torch.randn(rows, H, device=dev).to(BF); DZo = torch.empty(rows, K, device=dev, dtype=BF)
    W = (torch.randn(K, H, device=dev) * 0.06).to(BF)
    XHp = torch.randn(rows, K, device=dev).to(BF); rstdp = torch.rand(rows, device=dev) + 0.5
    sp = torch.ones(K, device=dev); bp = torch.zeros(K, device=dev)
    gs = torch.zeros(K, device=dev); gb = torch.zeros(K, device=dev)
    dx = lambda: call('mlb_dense_dx_lnbwd_tc', ptr(DZ), ptr(W), ptr(sp), ptr(bp), ptr(XHp), ptr(rstdp), ptr(DZo),
                      ptr(gs), ptr(gb), c_int(rows), c_int(H), c_int(K), c_int(H), c_int(H))
    print(f'dx K={K} rows={rows} tiles/SM={rows / 128 / 148:.2f}: {t(dx):.1f} us')
    gW = torch.zeros(K, H, device=dev)
    from madrona_learn_b200.engine import gemm_tc
    splitk = max(1, min(148 // (-(-K // 128) * -(-H // 128)), rows // 256))
    dw = lambda: gemm_tc(XHp, DZ, gW, None, K, H, rows, K, H, H, 1, 1, 2, splitk)
    print(f'dW K={K} rows={rows}: {t(dw):.1f} us')
