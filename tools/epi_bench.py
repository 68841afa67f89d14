"""Times the persistent fused-layer kernels (forward and backward-dx) back to back over rotating,
larger-than-L2 buffer sets for the epilogue generation selected by MLB_TC_EPI; prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import madrona_learn_b200 as m  # noqa: F401
from madrona_learn_b200._lib import c_int, call, ptr


def main():
    dev = torch.device('cuda', 0)
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    BF = torch.bfloat16
    Wm = (torch.randn(H, H, device=dev) * 0.06).to(BF)
    sc, bi = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    fw, dx = [], []
    for _ in range(4):
        X = torch.randn(rows, H, device=dev).to(BF)
        G = (torch.randn(rows, H, device=dev) * 1e-3).to(BF)
        O1 = torch.empty(rows, H, device=dev, dtype=BF)
        XH = torch.randn(rows, H, device=dev).to(BF)
        rs = torch.rand(rows, device=dev) + 0.5
        ds, db = torch.zeros(H, device=dev), torch.zeros(H, device=dev)
        O2 = torch.empty(rows, H, device=dev, dtype=BF)
        fw.append(lambda X=X, O1=O1, O2=O2, rs=rs: call(
            'mlb_dense_ln_relu_fwd_tc', ptr(X), ptr(Wm), ptr(sc), ptr(bi), ptr(O1), ptr(O2), ptr(rs),
            c_int(rows), c_int(H), c_int(H), c_int(H), c_int(H)))
        dx.append(lambda G=G, O1=O1, XH=XH, rs=rs, ds=ds, db=db: call(
            'mlb_dense_dx_lnbwd_tc', ptr(G), ptr(Wm), ptr(sc), ptr(bi), ptr(XH), ptr(rs), ptr(O1), ptr(ds),
            ptr(db), c_int(rows), c_int(H), c_int(H), c_int(H), c_int(H)))
    t_f = bench.time_kernel_rotating(torch, fw, 40)
    t_d = bench.time_kernel_rotating(torch, dx, 40)
    by = rows * H * 2 * 3 + rows * 4 + H * H * 2
    pk = bench.peaks()[0]['hbm_gbs']
    print(json.dumps(dict(epi=os.environ.get('MLB_TC_EPI', '0'), rows=rows, H=H, fwd_us=t_f * 1e6, dx_us=t_d * 1e6,
                          fwd_frac=by / t_f / 1e9 / pk, dx_frac=by / t_d / 1e9 / pk)))


if __name__ == '__main__':
    main()
