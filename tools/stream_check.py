"""Direct check of the weight-streaming fused-layer kernels (HN = 512) at the cfg3 shapes against a torch
fp32 restatement of the same math on the same bf16 operands."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import madrona_learn_b200 as m
from madrona_learn_b200._lib import c_int, call, ptr

dev = torch.device('cuda', 0)
BF = torch.bfloat16


def check(M, K, H):
    g = torch.Generator(device=dev).manual_seed(M + K)
    x = torch.randn(M, K, device=dev, generator=g).to(BF)
    wt = (torch.randn(H, K, device=dev, generator=g) / K ** 0.5).to(BF)
    s = 1 + 0.1 * torch.randn(H, device=dev, generator=g)
    b = 0.1 * torch.randn(H, device=dev, generator=g)
    y = torch.empty(M, H, device=dev, dtype=BF)
    xh = torch.empty(M, H, device=dev, dtype=BF)
    rstd = torch.empty(M, device=dev)
    for xh_arg in (xh, None):
        y.zero_()
        call('mlb_dense_ln_relu_fwd_tc', ptr(x), ptr(wt), ptr(s), ptr(b), ptr(y), ptr(xh_arg),
             ptr(rstd if xh_arg is not None else None), c_int(M), c_int(K), c_int(H), c_int(K), c_int(K))
        torch.cuda.synchronize()
        rows = slice(0, M, max(1, M // 4096))
        z = x[rows].float() @ wt.float().t()
        mu = z.mean(-1, keepdim=True)
        var = (z * z).mean(-1, keepdim=True) - mu * mu
        r = torch.rsqrt(var.clamp_min(0) + 1e-6)
        xr = (z - mu) * r
        yr = torch.relu(xr * s + b)
        e = (y[rows].float() - yr).abs().max().item()
        print(f'fwd M={M} K={K} xh={xh_arg is not None} max|dy|={e:.4f}', flush=True)
        assert e < 0.06
    # backward: dz_in [M, K2] with K2 = H (inner layer) ; W [H, K2]
    for K2 in (H, 64):
        dzin = torch.randn(M, K2, device=dev, generator=g).to(BF)
        w = (torch.randn(H, K2, device=dev, generator=g) / K2 ** 0.5).to(BF)
        dz = torch.empty(M, H, device=dev, dtype=BF)
        ds = torch.zeros(H, device=dev)
        db = torch.zeros(H, device=dev)
        call('mlb_dense_dx_lnbwd_tc', ptr(dzin), ptr(w), ptr(s), ptr(b), ptr(xh), ptr(rstd), ptr(dz), ptr(ds), ptr(db),
             c_int(M), c_int(K2), c_int(H), c_int(K2), c_int(K2))
        torch.cuda.synchronize()
        rows = slice(0, M, max(1, M // 4096))
        dy = dzin[rows].float() @ w.float().t()
        xhf = xh[rows].float()
        du = torch.where(xhf * s + b > 0, dy, torch.zeros_like(dy))
        dxh = du * s
        m1 = dxh.mean(-1, keepdim=True)
        m2 = (dxh * xhf).mean(-1, keepdim=True)
        ref = rstd[rows, None] * (dxh - m1 - xhf * m2)
        e = (dz[rows].float() - ref).abs().max().item() / ref.abs().max().item()
        print(f'dx  M={M} K={K2} rel max err={e:.4f}', flush=True)
        assert e < 0.02
        if M <= 65536:
            dyf = dzin.float() @ w.float().t()
            duf = torch.where(xh.float() * s + b > 0, dyf, torch.zeros_like(dyf))
            e2 = ((ds - (duf * xh.float()).sum(0)).norm() / ds.norm()).item()
            e3 = ((db - duf.sum(0)).norm() / db.norm()).item()
            print(f'     dscale rel {e2:.4f} dbias rel {e3:.4f}', flush=True)
            assert e2 < 0.02 and e3 < 0.02


for M, K in ((32768, 64), (65536, 64), (65536, 512), (262144, 512), (100000, 512)):
    check(M, K, 512)
print('stream_check ok')
