"""GAE kernel at the mid-size sweep points for one CTA size (MLB_GAE_BLOCK); prints fraction of the HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import madrona_learn_b200 as mlb
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gae_sweep import peak, time_op
K = mlb.kernels
dev = 'cuda:0'
pk, _ = peak()
flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
out = {}
for T, N in ((128, 65536), (256, 65536), (32, 262144), (64, 262144), (16, 1048576), (256, 16384)):
    r, v = torch.randn(T, N, device=dev), torch.randn(T, N, device=dev)
    d = torch.rand(T, N, device=dev) < 0.02
    b = torch.randn(N, device=dev)
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    t = time_op(lambda: K.gae(r, v, d, b, 0.99, 0.95, advantages=adv, returns=ret), flush)
    by = 17.0 * T * N + 4.0 * N
    out[f'{T}x{N}'] = round(by / t / 1e9 / pk, 3)
print(json.dumps(dict(block=os.environ.get('MLB_GAE_BLOCK', 'auto'), frac=out)))
