"""GPU parity of the multi-policy learner (multi_policy.py; ml/train.py:165-174 vmapped over num_train_policies,
ml/rollouts.py:579-588 `_sim_to_train`): P policies with their own parameters / optimiser / rollout stores on ONE
simulator, stepped in lockstep.  Every policy is verified like the single-policy learner (tests/test_train_gpu.py):
the simulator transitions replayed bit-exactly by the oracle env from the scattered actions, the stored
log-probs / values re-derived by the oracle forward of THAT policy's parameters on THAT policy's rows, GAE
bit-exact, and the PPO update re-run by the oracle from (params_p, key_p).  Both the block assignment
(`x.reshape(P, -1)`) and an interleaved assignment vector routed through pbt_reorder's chunk indices."""
import numpy as np
import pytest
import torch

from oracle import algo_common as oac
from oracle import env as oenv
from oracle import layouts, nn as onn, ppo as oppo

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
BUCKETS = [4, 8, 5, 5, 2, 2]


def _cfg(m, N, T, M, E, P, seed=5, dtype=torch.float32):
    return m.TrainConfig(
        num_worlds=N, num_agents_per_world=1, num_updates=10, actions={'act': m.DiscreteActionsConfig(BUCKETS)},
        steps_per_update=T, lr=3e-4,
        algo=m.PPOConfig(num_epochs=E, minibatch_size=M, clip_coef=0.2, value_loss_coef=0.5,
                         entropy_coef={'act': 0.01}, max_grad_norm=0.5),
        num_bptt_chunks=1, gamma=0.99, seed=seed, metrics_buffer_size=4, gae_lambda=0.95, dreamer_v3_critic=False,
        compute_dtype=dtype,
        pbt=m.PBTConfig(num_teams=1, team_size=1, num_train_policies=P, num_past_policies=0, self_play_portion=1.0,
                        cross_play_portion=0.0, past_play_portion=0.0))


def _policy(m, H=64, L=2):
    return m.Policy(actor_critic=m.ActorCritic(
        backbone=m.BackboneShared(prefix=None, encoder=m.BackboneEncoder(net=m.models.MLP(H, L))),
        actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(BUCKETS)), critic=m.models.DenseLayerCritic()))


def _ocfg(cfg):
    a = cfg.algo
    return oppo.PPOCfg(BUCKETS, num_epochs=a.num_epochs, minibatch_size=a.minibatch_size, clip_coef=a.clip_coef,
                       value_loss_coef=a.value_loss_coef, entropy_coef=a.entropy_coef['act'],
                       max_grad_norm=a.max_grad_norm, lr=cfg.lr, gamma=cfg.gamma, gae_lambda=cfg.gae_lambda)


@pytest.mark.parametrize('P,interleaved', [(2, False), (3, False), (2, True)])
def test_multi_policy_update_matches_oracle_per_policy(mlb, monkeypatch, P, interleaved):
    from madrona_learn_b200.multi_policy import MultiPolicyTrainingManager, init_multi_policy_training
    monkeypatch.setenv('MLB_CUDA_GRAPH', '0')
    m = mlb
    N, T, D, A = 24 * P, 8, 16, len(BUCKETS)
    B = N // P
    cfg = _cfg(m, N, T, 12, 2, P)
    env = m.SyntheticVectorEnv(N, D, A, seed=11, p_done=0.1, device=DEV)
    assign = None
    if interleaved:
        rng = np.random.default_rng(1)
        a = np.repeat(np.arange(P), B)
        rng.shuffle(a)
        assign = torch.from_numpy(a.astype(np.int32))
        mgr = init_multi_policy_training(torch.device(DEV), cfg, env.sim_fns(), _policy(m), None, m.TrainHooks(),
                                         None, None, None, assignments=assign)
        rows = [np.where(a == p)[0] for p in range(P)]
    else:
        mgr = m.init_training(DEV, cfg, env.sim_fns(), _policy(m), None, verbose=False)
        rows = [np.arange(p * B, (p + 1) * B) for p in range(P)]
    assert isinstance(mgr, MultiPolicyTrainingManager) and mgr.P == P
    np.testing.assert_array_equal(mgr.policy_assignments.cpu().numpy()[rows[P - 1]], P - 1)
    progs = [s.state.policy_states.program for s in mgr.subs]
    p0 = [pr.to_oracle_params() for pr in progs]
    assert not np.array_equal(p0[0]['mlp'][0]['kernel'], p0[1]['mlp'][0]['kernel'])     # independent inits
    key0 = [s.state.train_states.update_prng_key.cpu().numpy().view(np.uint32).copy() for s in mgr.subs]
    assert not np.array_equal(key0[0], key0[1])

    mgr.update_iter()
    torch.cuda.synchronize()
    sts = [{k: v.cpu().numpy() for k, v in s.rollout_mgr.store.items()} for s in mgr.subs]

    # (a) ONE simulator, stepped with the scattered actions of all policies: bit-exact replay
    ref_env = oenv.SyntheticEnv(N, D, A, seed=11, p_done=env.p_done)
    for t in range(T):
        acts = np.zeros((N, A), np.int32)
        for p in range(P):
            np.testing.assert_array_equal(sts[p]['obs'].reshape(T, B, D)[t], ref_env.obs[rows[p]])
            acts[rows[p]] = sts[p]['actions'].reshape(T, B, A)[t]
        _, r, d = ref_env.step(acts)
        for p in range(P):
            np.testing.assert_array_equal(sts[p]['rewards'].reshape(T, B)[t], r[rows[p]])
            np.testing.assert_array_equal(sts[p]['dones'].reshape(T, B)[t], d[rows[p]])
    np.testing.assert_array_equal(mgr.actions.cpu().numpy(), acts)               # the simulator's action buffer

    ocfg = _ocfg(cfg)
    sub_cfg = mgr.subs[0].cfg
    for p in range(P):
        st, s = sts[p], mgr.subs[p]
        # (b) stored log-probs / values == oracle forward of policy p's initial parameters on its rows
        p64 = onn.cast_tree(p0[p], np.float64)
        logits, critic, _ = onn.actor_critic_fwd(p64, st['obs'].reshape(T * B, D).astype(np.float64))
        lp, _ = onn.action_stats(logits, st['actions'].reshape(T * B, A), BUCKETS)
        np.testing.assert_allclose(st['log_probs'].reshape(T * B, A), lp, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(st['values'].reshape(T * B, 1), critic, rtol=1e-4, atol=1e-5)
        # (c) GAE per policy, bit-exact
        boot = s.rollout_mgr.bootstrap.cpu().numpy()
        _, bcrit, _ = onn.actor_critic_fwd(p64, ref_env.obs[rows[p]].astype(np.float64))
        np.testing.assert_allclose(boot.reshape(B, 1), bcrit, rtol=1e-4, atol=1e-5)
        adv = oac.compute_advantages(cfg.gamma, cfg.gae_lambda, st['rewards'], st['values'], st['dones'], boot)
        np.testing.assert_array_equal(st['advantages'], adv)
        # (d) policy p's PPO update re-run by the oracle from (p0[p], key0[p])
        roll = {k: layouts.reorder_seq_data(st[k])[0] for k in
                ('obs', 'actions', 'log_probs', 'advantages', 'returns', 'values', 'dones')}
        p1, opt1, key1, _, last, perms = oppo.ppo_update(p0[p], oppo.adam_init(p0[p]), oppo.initial_weight_norms(p0[p]),
                                                         roll, ocfg, key0[p], None, dtype=np.float32)
        np.testing.assert_array_equal(s.ppo_ws.perm.cpu().numpy(), perms)
        np.testing.assert_array_equal(s.state.train_states.update_prng_key.cpu().numpy().view(np.uint32), key1)
        got = progs[p].to_oracle_params()
        num = sum(float(np.sum(np.square(a_.astype(np.float64) - b_))) for a_, b_ in
                  zip(onn.tree_leaves(got), onn.tree_leaves(p1)))
        den = sum(float(np.sum(np.square(a_.astype(np.float64) - b_))) for a_, b_ in
                  zip(onn.tree_leaves(p1), onn.tree_leaves(p0[p])))
        assert den > 0 and np.sqrt(num / den) < 2e-2, (p, np.sqrt(num / den))
        np.testing.assert_allclose(s.metrics.latest()['Loss'].mean, last['loss'], rtol=5e-3, atol=1e-5)
    assert sub_cfg.num_worlds == N // P and mgr.update_idx == 1


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_multi_policy_graph_replay(mlb, dtype):
    """The whole P-policy update captured in one CUDA graph: eager update, capture, replays; every policy keeps
    learning on its own data (parameters move, differ between policies, metrics advance)."""
    m = mlb
    P, N, T = 2, 512, 16
    env = m.SyntheticVectorEnv(N, 64, len(BUCKETS), seed=3, device=DEV)
    mgr = m.init_training(DEV, _cfg(m, N, T, 128, 2, P, dtype=dtype), env.sim_fns(), _policy(m, 128, 2), None,
                          verbose=False)
    p0 = [s.state.policy_states.program.params.clone() for s in mgr.subs]
    for _ in range(4):
        mgr.update_iter()
    torch.cuda.synchronize()
    assert mgr._graph is not None and mgr.update_idx == 4
    d = [(s.state.policy_states.program.params - p0[i]) for i, s in enumerate(mgr.subs)]
    assert all(float(x.abs().max()) > 0 and bool(torch.isfinite(x).all()) for x in d)
    assert not torch.equal(d[0], d[1])
    for s in mgr.subs:
        assert np.isfinite(s.metrics.latest()['Loss'].mean) and s.update_idx == 4


def test_multi_policy_refuses_matchmaking(mlb):
    import dataclasses
    m = mlb
    env = m.SyntheticVectorEnv(64, 16, len(BUCKETS), seed=3, device=DEV)
    cfg = _cfg(m, 64, 8, 16, 1, 2)
    bad = dataclasses.replace(cfg, pbt=dataclasses.replace(cfg.pbt, num_past_policies=2, self_play_portion=0.5,
                                                           past_play_portion=0.5))
    with pytest.raises(NotImplementedError):
        m.init_training(DEV, bad, env.sim_fns(), _policy(m), None, verbose=False)
    with pytest.raises(ValueError):
        m.init_training(DEV, dataclasses.replace(cfg, num_worlds=63), env.sim_fns(), _policy(m), None, verbose=False)
