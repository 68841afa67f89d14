"""Times mlb_gemm_tf32_tc against mlb_gemm_f32 on the cfg2 minibatch shapes (65536 rows) and the whole cfg2 update
with compute_dtype=float32 under both matmul precisions."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import madrona_learn_b200 as m  # noqa: E402
from madrona_learn_b200._lib import c_int, call, ptr  # noqa: E402

DEV = 'cuda:0'


def time_gemm(name, fn, M, N, K, ta, tb, acc, iters=20):
    A = torch.randn((K, M) if ta else (M, K), device=DEV)
    B = torch.randn((N, K) if tb else (K, N), device=DEV)
    C = torch.zeros(M, N, device=DEV)
    flush = torch.zeros(64 << 20, device=DEV)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2 * iters)]
    for i in range(iters + 3):
        flush.zero_()
        if i >= 3:
            e[2 * (i - 3)].record()
        call(fn, ptr(A), ptr(B), ptr(C), ptr(None), c_int(M), c_int(N), c_int(K), c_int(A.shape[1]), c_int(B.shape[1]),
             c_int(N), c_int(ta), c_int(tb), c_int(acc), c_int(1 if fn == 'mlb_gemm_tf32_tc' or not acc else 16))
        if i >= 3:
            e[2 * (i - 3) + 1].record()
    torch.cuda.synchronize()
    us = sorted(e[2 * i].elapsed_time(e[2 * i + 1]) * 1e3 for i in range(iters))[iters // 2]
    byt = 4 * (M * K + N * K + M * N)
    print(json.dumps(dict(gemm=name, fn=fn, M=M, N=N, K=K, us=round(us, 2), tflops=round(2 * M * N * K / us / 1e6, 1),
                          gbs=round(byt / us / 1e3, 1))), flush=True)


if __name__ == '__main__':
    R = 65536
    print('MLB_TF32_ROUND=%s MLB_TF32_STAGES=%s' % (os.environ.get('MLB_TF32_ROUND', '1'), os.environ.get('MLB_TF32_STAGES', '2')))
    for fn in ('mlb_gemm_tf32_tc',) + (('mlb_gemm_f32',) if 'sgemm' in sys.argv else ()):
        time_gemm('fwd Z=XW', fn, R, 256, 256, 0, 0, 0)
        time_gemm('fwd0 Z=XW (K=64)', fn, R, 256, 64, 0, 0, 0)
        time_gemm('dX=dZ W^T', fn, R, 256, 256, 0, 1, 0)
        time_gemm('dW=X^T dZ', fn, 256, 256, R, 1, 0, 1)
        time_gemm('head fwd', fn, R, 28, 256, 0, 0, 0)
        time_gemm('head dX', fn, R, 256, 28, 0, 1, 0)
        time_gemm('head dW', fn, 256, 28, R, 1, 0, 1)
    # fused Dense + LayerNorm + ReLU layer vs GEMM + LayerNorm kernel
    x = torch.randn(R, 256, device=DEV); w = torch.randn(256, 256, device=DEV) / 16
    sc = torch.ones(256, device=DEV); bi = torch.zeros(256, device=DEV)
    z = torch.empty(R, 256, device=DEV); y = torch.empty(R, 256, device=DEV); st = torch.empty(R, 2, device=DEV)
    flush = torch.zeros(64 << 20, device=DEV)
    from madrona_learn_b200._lib import c_ll
    for name, fn in (('fused layer', lambda: call('mlb_dense_ln_relu_fwd_tf32', ptr(x), ptr(w), ptr(sc), ptr(bi), ptr(z), ptr(y), ptr(st), c_ll(R), c_int(256), c_int(256), c_int(256))),
                     ('gemm + ln', lambda: (call('mlb_gemm_tf32_tc', ptr(x), ptr(w), ptr(z), ptr(None), c_int(R), c_int(256), c_int(256), c_int(256), c_int(256), c_int(256), c_int(0), c_int(0), c_int(0), c_int(1)),
                                            call('mlb_ln_relu_fwd_f32', ptr(z), ptr(sc), ptr(bi), ptr(y), ptr(st), c_ll(R), c_int(256))))):
        ev = []
        for i in range(13):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); ev.append((e0, e1))
        torch.cuda.synchronize()
        us = sorted(a.elapsed_time(b) * 1e3 for a, b in ev[3:])[5]
        print(json.dumps(dict(layer=name, rows=R, H=256, K=256, us=round(us, 2))), flush=True)
    if 'gemm-only' in sys.argv:
        sys.exit(0)
    import bench_configs as b
    m.set_matmul_precision('tf32')
    b.run('cfg2 f32 matmul=tf32', 8192, 32, 1, 256, 3, 4, 4, torch.float32, steps=10, warm=3)
    if 'cfg4' in sys.argv:
        b.run('cfg4 f32 matmul=tf32', 16384, 128, 4, 256, 2, 4, 2, torch.float32, rnn=256, normalize_values=True, steps=2, warm=2)
