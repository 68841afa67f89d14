"""GPU parity of the DreamerV3 two-hot critic (the reference's DEFAULT critic, ml/cfg.py:85):
sampling kernel emits the symexp two-hot mean as the value estimate, the loss kernel the
two-hot cross-entropy and its gradient; full update_iter with dreamer_v3_critic=True."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import algo_common as oac
from oracle import dists as odists, layouts, nn as onn, ppo as oppo

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
BUCKETS = [4, 8, 5, 5, 2, 2]


def _rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30)


def _policy(m, H, L):
    return m.Policy(actor_critic=m.ActorCritic(
        backbone=m.BackboneShared(prefix=None, encoder=m.BackboneEncoder(net=m.models.MLP(H, L))),
        actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(BUCKETS)),
        critic=m.models.DreamerV3Critic()))


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_twohot_loss_and_values_vs_oracle(mlb, dtype):
    from madrona_learn_b200._lib import c_float, c_int, c_ll, c_size_t, call, ptr
    from madrona_learn_b200.engine import PolicyProgram
    m = mlb
    D, H, L, Tp, M = 32, 64, 2, 4, 256
    rows, A, V = Tp * M, len(BUCKETS), 63
    rng = np.random.default_rng(7)
    p = onn.init_params(rng, D, H, L, BUCKETS, critic_dim=V)
    p['actor']['kernel'] = (rng.standard_normal(p['actor']['kernel'].shape) * 0.2).astype(np.float32)
    p['critic']['kernel'] = (rng.standard_normal((H, V)) * 0.3).astype(np.float32)
    p['critic']['bias'] = (rng.standard_normal(V) * 0.1).astype(np.float32)
    prog = PolicyProgram(_policy(m, H, L).actor_critic, D, {'act': m.DiscreteActionsConfig(BUCKETS)}, DEV, dtype)
    assert prog.twohot and prog.V == V
    np.testing.assert_array_equal(np.array(list(prog._bins_c), np.float32), odists.bins(V))
    prog.load_oracle_params(p)
    cfg = oppo.PPOCfg(BUCKETS, entropy_coef=0.02, dreamer_v3_critic=True)
    mb = dict(obs=rng.standard_normal((Tp, M, D)).astype(np.float32),
              actions=np.stack([rng.integers(0, b, (Tp, M)) for b in BUCKETS], -1).astype(np.int32),
              advantages=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              returns=(rng.standard_normal((Tp, M, 1)) * 20).astype(np.float32),
              values=rng.standard_normal((Tp, M, 1)).astype(np.float32), mb_weights=np.ones((M, 1), np.float32))
    mb['returns'][0, :4, 0] = [0.0, 1e7, -1e7, odists.bins(V)[40]]          # bin edges / clipping cases
    mb['log_probs'] = (-np.abs(rng.standard_normal((Tp, M, A))) - 0.5).astype(np.float32)
    quant = onn.bf16_round if dtype == torch.bfloat16 else None
    ref = oppo.ppo_loss(p, mb, cfg, None, dtype=np.float64, quant=quant)
    dv = {k: torch.from_numpy(v).to(DEV) for k, v in mb.items()}
    obs_d = dv['obs'].view(rows, D)
    head = prog.forward_train(obs_d, rows)
    tol = 3e-3 if dtype == torch.bfloat16 else 1e-4
    h = head.cpu().numpy()
    assert _rel(h[:, 26:26 + V], ref['critic']) < tol
    # value estimate = two-hot mean (rollout path)
    out, _ = prog.apply_rollout(torch.zeros(2, dtype=torch.int32, device=DEV), (), {'obs': obs_d})
    np.testing.assert_allclose(out['critic'].cpu().numpy(), odists.twohot_mean(h[:, 26:26 + V]), rtol=1e-3, atol=1e-3)
    tw = prog.train_ws(rows)
    mean, rstd = oac.zscore_stats(mb['advantages'])
    adv_mr = torch.tensor([mean, rstd, 0, 0], dtype=torch.float32, device=DEV)
    obj_scale = (ctypes.c_float * A)(*[1.0 / (rows * A)] * A)
    ent_scale = (ctypes.c_float * A)(*[cfg.entropy_coef / (rows * A)] * A)
    prog.zero_grads()
    call('mlb_ppo_loss_f32', ptr(head), c_int(prog.NH), ptr(dv['actions']), ptr(dv['log_probs']),
         ptr(dv['advantages']), ptr(dv['returns']), ptr(None), ptr(None), ptr(adv_mr), ptr(None),
         prog._buckets_c, obj_scale, ent_scale, c_int(A), c_ll(rows), c_ll(M), c_float(cfg.clip_coef),
         c_float(cfg.value_loss_coef), c_int(prog.loss_flags), ptr(tw['dhead']), ptr(prog.head_bias_grad()),
         ptr(tw['stats_out']), ptr(tw['loss_ws']), c_size_t(tw['loss_ws'].numel()), prog._bins_c, c_int(V))
    from madrona_learn_b200 import _lib
    stt = _lib.PPOStats.from_buffer_copy(tw['stats_out'].cpu().numpy().tobytes())
    np.testing.assert_allclose(stt.loss, ref['loss'], rtol=10 * tol, atol=1e-5)
    np.testing.assert_allclose(stt.value_loss, cfg.value_loss_coef * np.mean(ref['value_losses']), rtol=10 * tol)
    np.testing.assert_allclose(stt.metrics[3].mean, np.mean(np.abs(ref['value_errs'])), rtol=10 * tol)
    dh = tw['dhead'].float().cpu().numpy()
    assert _rel(dh[:, 26:26 + V], ref['dcritic']) < (2e-2 if dtype == torch.bfloat16 else 1e-4)
    prog.backward(obs_d, rows)
    g = prog.to_oracle_params(prog.grads)
    gt = 4e-2 if dtype == torch.bfloat16 else 2e-4
    onn.tree_map(lambda a, b: np.testing.assert_array_less(_rel(a, b), gt), g, ref['grads'])


def test_default_critic_update_iter(mlb, monkeypatch):
    """TrainConfig's DEFAULT critic (dreamer_v3_critic=True) end to end vs the oracle update."""
    monkeypatch.setenv('MLB_CUDA_GRAPH', '0')
    m = mlb
    N, T, M, E, D, H, L = 48, 16, 12, 2, 16, 64, 2
    env = m.SyntheticVectorEnv(N, D, len(BUCKETS), seed=2, p_done=0.1, device=DEV)
    cfg = m.TrainConfig(num_worlds=N, num_agents_per_world=1, num_updates=10,
                        actions={'act': m.DiscreteActionsConfig(BUCKETS)}, steps_per_update=T, lr=3e-4,
                        algo=m.PPOConfig(num_epochs=E, minibatch_size=M, clip_coef=0.2, value_loss_coef=0.5,
                                         entropy_coef={'act': 0.01}, max_grad_norm=0.5),
                        num_bptt_chunks=1, gamma=0.99, seed=5, metrics_buffer_size=4, gae_lambda=0.95)
    assert cfg.dreamer_v3_critic
    mgr = m.init_training(DEV, cfg, env.sim_fns(), _policy(m, H, L), None, verbose=False)
    prog = mgr.state.policy_states.program
    # give the (zero-initialised) critic some signal so the test is not vacuous
    p0 = prog.to_oracle_params()
    rng = np.random.default_rng(0)
    p0['critic']['kernel'] = (rng.standard_normal(p0['critic']['kernel'].shape) * 0.2).astype(np.float32)
    prog.load_oracle_params(p0)
    key0 = mgr.state.train_states.update_prng_key.cpu().numpy().view(np.uint32).copy()
    mgr.update_iter()
    torch.cuda.synchronize()
    st = {k: v.cpu().numpy() for k, v in mgr.rollout_mgr.store.items()}
    A = len(BUCKETS)
    p64 = onn.cast_tree(p0, np.float64)
    logits, clog, _ = onn.actor_critic_fwd(p64, st['obs'].reshape(T * N, D).astype(np.float64))
    np.testing.assert_allclose(st['values'].reshape(T * N, 1), odists.twohot_mean(clog.astype(np.float32)),
                               rtol=1e-3, atol=1e-3)
    adv = oac.compute_advantages(cfg.gamma, cfg.gae_lambda, st['rewards'], st['values'], st['dones'],
                                 mgr.rollout_mgr.bootstrap.cpu().numpy())
    np.testing.assert_array_equal(st['advantages'], adv)
    ocfg = oppo.PPOCfg(BUCKETS, num_epochs=E, minibatch_size=M, entropy_coef=0.01, lr=cfg.lr, dreamer_v3_critic=True)
    roll = {k: layouts.reorder_seq_data(st[k])[0] for k in
            ('obs', 'actions', 'log_probs', 'advantages', 'returns', 'values', 'dones')}
    opt, norms = oppo.adam_init(p0), oppo.initial_weight_norms(p0)
    p1, opt1, key1, _, last, perms = oppo.ppo_update(p0, opt, norms, roll, ocfg, key0, None, dtype=np.float32)
    np.testing.assert_array_equal(mgr.ppo_ws.perm.cpu().numpy(), perms)
    got = prog.to_oracle_params()
    num = sum(float(np.sum(np.square(a.astype(np.float64) - b))) for a, b in zip(onn.tree_leaves(got), onn.tree_leaves(p1)))
    den = sum(float(np.sum(np.square(a.astype(np.float64) - b))) for a, b in zip(onn.tree_leaves(p1), onn.tree_leaves(p0)))
    assert np.sqrt(num / den) < 2e-2, np.sqrt(num / den)
    np.testing.assert_allclose(mgr.metrics.latest()['Value Loss'].mean, np.mean(last['value_losses']), rtol=5e-3)
