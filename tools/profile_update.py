"""Per-kernel device-time breakdown of one update_iter (eager, CUPTI through torch.profiler):
python tools/profile_update.py cfg2|cfg3|cfg4 [updates].  Prints the kernels sorted by total device time."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['MLB_CUDA_GRAPH'] = '0'
import torch
from torch.profiler import ProfilerActivity, profile

import bench
import madrona_learn_b200 as m


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
    n_upd = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    wl = bench.CFG3 if name == 'cfg3' else bench.WORKLOAD
    dev = torch.device('cuda', 0)
    if name == 'cfg4':          # recurrent actor-critic: LSTM 256, 16384 worlds x 128 steps, 4 BPTT chunks, value-norm EMA
        N, T, C, H, L, RH = 16384, 128, 4, 256, 2, 256
        policy = m.Policy(actor_critic=m.ActorCritic(
            backbone=m.BackboneShared(prefix=None, encoder=m.RecurrentBackboneEncoder(
                net=m.models.MLP(H, L), rnn=m.rnn.LSTM(RH, 1))),
            actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(bench.BUCKETS)),
            critic=m.models.DenseLayerCritic()))
        env = m.SyntheticVectorEnv(N, 64, len(bench.BUCKETS), seed=0, device=dev)
        cfg = m.TrainConfig(
            num_worlds=N, num_agents_per_world=1, num_updates=1 << 30,
            actions={'act': m.DiscreteActionsConfig(bench.BUCKETS)}, steps_per_update=T, lr=3e-4,
            algo=m.PPOConfig(num_epochs=2, minibatch_size=N * C // 4, clip_coef=0.2, value_loss_coef=0.5,
                             entropy_coef={'act': 0.01}, max_grad_norm=0.5),
            num_bptt_chunks=C, gamma=0.99, seed=0, metrics_buffer_size=4, gae_lambda=0.95,
            dreamer_v3_critic=False, normalize_values=True, compute_dtype=torch.bfloat16)
        mgr = m.init_training(dev, cfg, env.sim_fns(), policy, None, verbose=False)
    else:
        N = wl['worlds']
        env = m.SyntheticVectorEnv(N, bench.WORKLOAD['obs_dim'], len(bench.BUCKETS), seed=0, device=dev)
        mgr = m.init_training(dev, bench.make_cfg(m, N, dtype=os.environ.get('PROFILE_DTYPE', 'bf16'), wl=wl), env.sim_fns(), bench.make_policy(m, wl),
                              None, verbose=False)
    mgr.update_iter()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n_upd):
            mgr.update_iter()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, 'device_time_total', None) or getattr(e, 'cuda_time_total', 0)
        if t:
            rows.append((t / n_upd, e.count / n_upd, e.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f'{name}: {tot / 1e3:.2f} ms of kernel time per update')
    out = []
    for t, cnt, key in rows[:25]:
        print(f'{t / 1e3:9.3f} ms {100 * t / tot:5.1f}%  x{cnt:6.1f}  {t / cnt:8.1f} us  {key[:110]}')
        out.append(dict(kernel=key, ms_per_update=t / 1e3, share=t / tot, launches=cnt, us_per_launch=t / cnt))
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(dict(config=name, kernel_ms_per_update=tot / 1e3, kernels=out),
              open(f'gpurun_out/r2_profile_{name}.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
