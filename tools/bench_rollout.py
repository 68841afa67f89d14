"""Micro-benchmark of the one-kernel rollout policy step (mlb_policy_rollout_tc) vs hidden width / rows."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import madrona_learn_b200 as m
from madrona_learn_b200.engine import PolicyProgram

DEV = 'cuda:0'
buckets = [4, 8, 5, 5, 2, 2]
for H, L, rows in ((256, 3, 8192), (128, 3, 8192), (256, 3, 4096), (256, 1, 8192), (256, 3, 16384)):
    ac = m.ActorCritic(
        backbone=m.BackboneShared(prefix=None, encoder=m.BackboneEncoder(net=m.models.MLP(H, L))),
        actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(buckets)),
        critic=m.models.DenseLayerCritic())
    prog = PolicyProgram(ac, 64, {'act': m.DiscreteActionsConfig(buckets)}, DEV, torch.bfloat16)
    prog.init_params(1); prog.finalize_params()
    obs = torch.randn(rows, 64, device=DEV)
    store = torch.empty_like(obs)
    k0 = torch.tensor([1, 2], dtype=torch.int32, device=DEV); k1 = torch.zeros_like(k0)
    a = torch.zeros(rows, 6, dtype=torch.int32, device=DEV); lp = torch.zeros(rows, 6, device=DEV); v = torch.zeros(rows, device=DEV)
    for name, fn in (('sample', lambda: prog.rollout_step_fused(obs, store, rows, k0, k1, a, lp, v)),
                     ('greedy', lambda: prog.rollout_step_fused(obs, None, rows, None, None, a, None, v, deterministic=True))):
        for _ in range(3): fn()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for _ in range(50): fn()
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); e1.record(st); torch.cuda.synchronize()
        print(f'H={H} L={L} rows={rows} {name}: {e0.elapsed_time(e1) * 1e3 / 50:.1f} us/step (graph of 50)')
