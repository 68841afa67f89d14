mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_recurrent_gpu.py -m gpu -q -x -s 2>&1 | grep "PARITY\|passed\|failed\|Error" | tail -12
timeout 600 python - <<'PY' 2>&1 | grep -v Warn | tail -4
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import torch, bench_configs as b
for dt in (torch.bfloat16,):
    b.run('cfg4 LSTM-256 actor-critic, 16384x128, 4 BPTT chunks, value-norm EMA', 16384, 128, 4, 256, 2, 4, 2, dt, rnn=256, normalize_values=True, steps=3, warm=3)
PY
MLB_CUDA_GRAPH=0 timeout 300 python - <<'PY' 2>&1 | grep -v Warn | tail -24
import sys, os; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import torch
from torch.profiler import ProfilerActivity, profile
import madrona_learn_b200 as m
BUCKETS = [4, 8, 5, 5, 2, 2]
DEV='cuda:0'
N,T,C,H,L,RH=16384,128,4,256,2,256
enc = m.RecurrentBackboneEncoder(net=m.models.MLP(H, L), rnn=m.rnn.LSTM(RH, 1))
policy = m.Policy(actor_critic=m.ActorCritic(backbone=m.BackboneShared(prefix=None, encoder=enc),
    actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(BUCKETS)), critic=m.models.DenseLayerCritic()))
env = m.SyntheticVectorEnv(N, 64, len(BUCKETS), seed=0, device=DEV)
cfg = m.TrainConfig(num_worlds=N, num_agents_per_world=1, num_updates=1 << 30, actions={'act': m.DiscreteActionsConfig(BUCKETS)},
    steps_per_update=T, lr=3e-4, algo=m.PPOConfig(num_epochs=2, minibatch_size=N*C // 4, clip_coef=0.2, value_loss_coef=0.5,
    entropy_coef={'act': 0.01}, max_grad_norm=0.5), num_bptt_chunks=C, gamma=0.99, seed=0, metrics_buffer_size=4, gae_lambda=0.95,
    dreamer_v3_critic=False, normalize_values=True, compute_dtype=torch.bfloat16)
mgr = m.init_training(DEV, cfg, env.sim_fns(), policy, None, verbose=False)
mgr.update_iter(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    mgr.update_iter(); torch.cuda.synchronize()
rows=[]
for e in prof.key_averages():
    t = getattr(e, 'device_time_total', None) or 0
    if t: rows.append((t, e.count, e.key))
rows.sort(reverse=True); tot=sum(r[0] for r in rows)
print('cfg4 bf16: %.2f ms kernel time' % (tot/1e3))
for t,c,k in rows[:18]: print('%9.3f ms %5.1f%% x%5d %8.1f us  %s' % (t/1e3, 100*t/tot, c, t/c, k[:100]))
PY
