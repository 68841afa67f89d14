"""Warm, in-situ phase timing of update_iter on cfg2: the rollout collection and the PPO learn
phase are captured as two separate CUDA graphs and replayed.  Diagnostic only (bench.py is the
measurement contract).  Usage: python tools/phase_times.py [--dtype bf16|f32] [--reps 30]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import madrona_learn_b200 as m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--reps', type=int, default=30)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    N = bench.WORKLOAD['worlds']
    env = m.SyntheticVectorEnv(N, bench.WORKLOAD['obs_dim'], len(bench.BUCKETS), seed=0, device=dev)
    os.environ['MLB_CUDA_GRAPH'] = '0'
    mgr = m.init_training(dev, bench.make_cfg(m, N, dtype=a.dtype), env.sim_fns(), bench.make_policy(m), None,
                          verbose=False)
    for _ in range(2):
        mgr.update_iter()
    torch.cuda.synchronize()
    from madrona_learn_b200.train import TrainHooks
    hooks = TrainHooks()
    rm, tsm, cfg = mgr.rollout_mgr, mgr.state, mgr.cfg
    algo = cfg.algo.setup()
    holder = {}

    def collect():
        _, rs, data, _, met = rm.collect(tsm, mgr.rollout, mgr.metrics, hooks.start_rollouts,
                                         hooks.finish_rollouts, hooks.rollout_metrics)
        holder['data'] = data

    def learn():
        algo.update(cfg, tsm.policy_states, tsm.train_states, holder['data'], hooks.optimize_metrics,
                    mgr.metrics, dist_ctx=None, ws=mgr.ppo_ws)

    def graph_time(fn):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                fn()
            for _ in range(3):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(s)
            for _ in range(a.reps):
                g.replay()
            e1.record(s)
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.reps

    tc = graph_time(collect)
    tl = graph_time(learn)
    print(f'collect {tc:.3f} ms   learn {tl:.3f} ms   sum {tc + tl:.3f} ms')


if __name__ == '__main__':
    main()
