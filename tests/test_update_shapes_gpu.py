"""GPU parity of ONE FULL update_iter on the path bench.py times: compute_dtype=bfloat16 (tcgen05
fused layers, one-kernel rollout step), at the BASELINE shapes

    cfg2        8192 worlds x 32 steps, MLP 3x256, minibatch 2048 trajectories (65 536 rows)
    cfg3-width  MLP 3x512 (the non-persistent fused kernels + layer-by-layer rollout), 4096 worlds
                x 32 steps, minibatch 1024 trajectories (32 768 rows: the oracle stays in seconds)

against oracle/ppo.ppo_update with the SAME bf16 quantisation points (quant=bf16_round) and against the
exact fp32 oracle.  What is checked (VERDICT r1 "next" 1a-1c):
  * minibatch permutations and the advanced update key: bit-exact;
  * rollout: stored log-probs / values of sampled rows re-derived by the oracle forward;
  * GAE / returns on the stored buffers: bit-exact;
  * the update itself: per-minibatch loss of the last minibatch, the Adam first/second moments
    (which integrate the gradients of all minibatches) and the parameter delta.

Tolerances.  Adam's first steps are sign-like (u = -lr * m / (sqrt(v) + eps) ~ -lr * sign(g)), so
the parameter DELTA amplifies the sign of every gradient element whose magnitude is below the bf16
noise floor; the Adam moment `m` is linear in the gradients and is the well-conditioned quantity.
Stated bounds: m vs the same-quantisation oracle rel-L2 <= 5e-3 / cosine >= 0.9999, parameter delta
cosine >= 0.999; vs the exact fp32 oracle (the precision cost of bf16 storage) m cosine >= 0.999 and
rel-L2 <= 2e-2 -- the tolerance SURVEY 8c states -- and parameter delta cosine >= 0.99.  The measured figures are appended to $MLB_PARITY_LOG when set (profiles/r2_parity_*.jsonl).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import algo_common as oac
from oracle import layouts, nn as onn, ppo as oppo

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
BUCKETS = [4, 8, 5, 5, 2, 2]


def _log(rec):
    path = os.environ.get('MLB_PARITY_LOG')
    if path:
        with open(path, 'a') as f:
            f.write(json.dumps(rec) + '\n')
    print('PARITY', json.dumps(rec))


def _flat(tree):
    return np.concatenate([np.asarray(x, np.float64).reshape(-1) for x in onn.tree_leaves(tree)])


def _cos_rel(a, b):
    return (float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300)),
            float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300)))


def _arena_tree(prog, arena):
    return prog.to_oracle_params(arena)


@pytest.mark.parametrize('name,N,T,H,M,E', [('cfg2', 8192, 32, 256, 2048, 1),
                                            ('cfg3-width', 4096, 32, 512, 1024, 1)])
def test_bf16_update_iter_vs_oracle_at_baseline_shape(mlb, monkeypatch, name, N, T, H, M, E):
    monkeypatch.setenv('MLB_CUDA_GRAPH', '0')
    m = mlb
    D, L, A = 64, 3, len(BUCKETS)
    env = m.SyntheticVectorEnv(N, D, A, seed=7, device=DEV)
    policy = m.Policy(actor_critic=m.ActorCritic(
        backbone=m.BackboneShared(prefix=None, encoder=m.BackboneEncoder(net=m.models.MLP(H, L))),
        actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(BUCKETS)),
        critic=m.models.DenseLayerCritic()))
    cfg = m.TrainConfig(
        num_worlds=N, num_agents_per_world=1, num_updates=10, actions={'act': m.DiscreteActionsConfig(BUCKETS)},
        steps_per_update=T, lr=3e-4,
        algo=m.PPOConfig(num_epochs=E, minibatch_size=M, clip_coef=0.2, value_loss_coef=0.5,
                         entropy_coef={'act': 0.01}, max_grad_norm=0.5),
        num_bptt_chunks=1, gamma=0.99, seed=3, metrics_buffer_size=4, gae_lambda=0.95,
        dreamer_v3_critic=False, normalize_values=False, compute_dtype=torch.bfloat16)
    mgr = m.init_training(DEV, cfg, env.sim_fns(), policy, None, verbose=False)
    prog = mgr.state.policy_states.program
    assert prog.tc and prog.fused_rollout == (H <= 256)
    p0 = prog.to_oracle_params()
    key0 = mgr.state.train_states.update_prng_key.cpu().numpy().view(np.uint32).copy()
    mgr.update_iter()
    torch.cuda.synchronize()
    st = {k: v.cpu().numpy() for k, v in mgr.rollout_mgr.store.items()}
    boot = mgr.rollout_mgr.bootstrap.cpu().numpy()

    # ---- rollout: sampled rows re-derived by the oracle forward with the same quantisation -----
    rng = np.random.default_rng(0)
    sel = rng.choice(T * N, 4096, replace=False)
    obs_sel = st['obs'].reshape(T * N, D)[sel]
    act_sel = st['actions'].reshape(T * N, A)[sel]
    logits, critic, _ = onn.actor_critic_fwd(onn.cast_tree(p0, np.float64), obs_sel.astype(np.float64),
                                             onn.bf16_round)
    lp, _ = onn.action_stats(logits, act_sel, BUCKETS)
    lp_got = st['log_probs'].reshape(T * N, A)[sel]
    v_got = st['values'].reshape(T * N, 1)[sel]
    _, rel_lp = _cos_rel(lp_got.reshape(-1).astype(np.float64), lp.reshape(-1))
    _, rel_v = _cos_rel(v_got.reshape(-1).astype(np.float64), critic.reshape(-1))
    assert rel_lp < 5e-3 and rel_v < 1e-2, (rel_lp, rel_v)

    # ---- GAE / returns: bit-exact on the stored buffers ------------------------------------------
    adv = oac.compute_advantages(cfg.gamma, cfg.gae_lambda, st['rewards'], st['values'], st['dones'], boot)
    np.testing.assert_array_equal(st['advantages'], adv)
    np.testing.assert_array_equal(st['returns'], (adv + st['values']).astype(np.float32))

    # ---- the update, re-run by the oracle from (p0, key0) on the stored rollout --------------------
    ocfg = oppo.PPOCfg(BUCKETS, num_epochs=E, minibatch_size=M, clip_coef=0.2, value_loss_coef=0.5,
                       entropy_coef=0.01, max_grad_norm=0.5, lr=cfg.lr, gamma=cfg.gamma, gae_lambda=cfg.gae_lambda)
    roll = {k: layouts.reorder_seq_data(st[k])[0] for k in
            ('obs', 'actions', 'log_probs', 'advantages', 'returns', 'values', 'dones')}
    norms = oppo.initial_weight_norms(p0)
    res = {}
    for tag, q in (('quant', onn.bf16_round), ('exact', None)):
        p1, opt1, key1, _, last, perms = oppo.ppo_update(p0, oppo.adam_init(p0), norms, roll, ocfg, key0, None,
                                                         dtype=np.float32, quant=q)
        res[tag] = (p1, opt1, key1, last, perms)
    p1, opt1, key1, last, perms = res['quant']
    np.testing.assert_array_equal(mgr.ppo_ws.perm.cpu().numpy(), perms)                      # bit-exact
    np.testing.assert_array_equal(mgr.state.train_states.update_prng_key.cpu().numpy().view(np.uint32), key1)
    assert prog.adam_step.item() == opt1['t'] == E * (N // M)

    got_p = prog.to_oracle_params()
    got_m = _arena_tree(prog, prog.adam_m)
    rec = dict(case=name, rows_per_minibatch=M * T, hidden=H, rollout_log_prob_rel=rel_lp, rollout_value_rel=rel_v)
    for tag in ('quant', 'exact'):
        p1, opt1, _, last, _ = res[tag]
        cm, rm = _cos_rel(_flat(got_m), _flat(opt1['m']))
        cd, rd = _cos_rel(_flat(got_p) - _flat(p0), _flat(p1) - _flat(p0))
        rec.update({f'adam_m_cos_{tag}': cm, f'adam_m_rel_{tag}': rm, f'delta_cos_{tag}': cd,
                    f'delta_rel_{tag}': rd})
    lat = mgr.metrics.latest()
    rec['loss_gpu'], rec['loss_quant'], rec['loss_exact'] = (float(lat['Loss'].mean), float(res['quant'][3]['loss']),
                                                             float(res['exact'][3]['loss']))
    _log(rec)
    # measured on B200 (profiles/r2_parity_measured.jsonl): cfg2 m rel 1.1e-3 (quant) / 1.07e-2 (exact),
    # delta cosine 0.99996 / 0.9975; width 512: 2.6e-4 / 2.4e-3, 0.99995 / 0.9981
    assert rec['adam_m_cos_quant'] >= 0.9999 and rec['adam_m_rel_quant'] <= 5e-3, rec
    assert rec['delta_cos_quant'] >= 0.999, rec
    assert rec['adam_m_cos_exact'] >= 0.999 and rec['adam_m_rel_exact'] <= 2e-2, rec      # SURVEY 8c bound
    assert rec['delta_cos_exact'] >= 0.99, rec
    np.testing.assert_allclose(rec['loss_gpu'], rec['loss_quant'], rtol=2e-2, atol=2e-4)
    np.testing.assert_allclose(lat['Entropy'].mean, np.mean(res['quant'][3]['entropies']), rtol=2e-3)
