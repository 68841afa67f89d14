"""Times update_iter for the other BASELINE.json configs on ONE B200 (documentation numbers;
bench.py is the contract benchmark and runs configs[1] only).
  cfg3: PPO MLP 3x512, 65536 worlds x 64 steps (single-GPU shard = the whole thing), bf16 path
  cfg4: recurrent (LSTM 256) actor-critic, 16384 worlds x 128 steps, 4 BPTT chunks, value-norm EMA, fp32
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import madrona_learn_b200 as m  # noqa: E402

BUCKETS = [4, 8, 5, 5, 2, 2]
DEV = 'cuda:0'


def run(name, N, T, C, H, L, mbs, epochs, dtype, rnn=None, normalize_values=False, steps=5, warm=3):
    enc = (m.RecurrentBackboneEncoder(net=m.models.MLP(H, L), rnn=m.rnn.LSTM(rnn, 1)) if rnn
           else m.BackboneEncoder(net=m.models.MLP(H, L)))
    policy = m.Policy(actor_critic=m.ActorCritic(
        backbone=m.BackboneShared(prefix=None, encoder=enc),
        actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(BUCKETS)),
        critic=m.models.DenseLayerCritic()))
    env = m.SyntheticVectorEnv(N, 64, len(BUCKETS), seed=0, device=DEV)
    J = N * C
    cfg = m.TrainConfig(num_worlds=N, num_agents_per_world=1, num_updates=1 << 30,
                        actions={'act': m.DiscreteActionsConfig(BUCKETS)}, steps_per_update=T, lr=3e-4,
                        algo=m.PPOConfig(num_epochs=epochs, minibatch_size=J // mbs, clip_coef=0.2,
                                         value_loss_coef=0.5, entropy_coef={'act': 0.01}, max_grad_norm=0.5),
                        num_bptt_chunks=C, gamma=0.99, seed=0, metrics_buffer_size=4, gae_lambda=0.95,
                        dreamer_v3_critic=False, normalize_values=normalize_values, compute_dtype=dtype)
    mgr = m.init_training(DEV, cfg, env.sim_fns(), policy, None, verbose=False)
    for _ in range(warm):
        mgr.update_iter()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        mgr.update_iter()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = dict(config=name, worlds=N, steps_per_update=T, bptt_chunks=C, ms_per_update=ms,
               agent_steps_per_s=N * T / ms * 1e3, dtype=str(dtype), loss=mgr.metrics.latest()['Loss'].mean,
               mem_gb=torch.cuda.max_memory_allocated() / 1e9)
    print(json.dumps(out), flush=True)
    del mgr, env
    torch.cuda.empty_cache()
    return out


if __name__ == '__main__':
    res = []
    res.append(run('cfg2 PPO MLP 3x256, 8192x32', 8192, 32, 1, 256, 3, 4, 4, torch.bfloat16))
    res.append(run('cfg3 PPO MLP 3x512, 65536x64 (1 GPU)', 65536, 64, 1, 512, 3, 4, 4, torch.bfloat16))
    res.append(run('cfg4 LSTM-256 actor-critic, 16384x128, 4 BPTT chunks, value-norm EMA', 16384, 128, 4,
                   256, 2, 4, 2, torch.float32, rnn=256, normalize_values=True, steps=3, warm=2))
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(res, open('gpurun_out/bench_configs.json', 'w'), indent=1)
