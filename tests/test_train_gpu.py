"""GPU end-to-end parity of init_training / TrainingManager.update_iter (BASELINE config 1
shape family: toy actor-critic on the synthetic vector env) against the oracle pipeline.

Sampling parity is distributional only, so the rollout is verified on the data
the GPU actually produced: env transitions replayed bit-exactly by the oracle env, stored
log-probs / values re-derived by the oracle forward, GAE bit-exact on the stored buffers,
and the whole PPO update (permutations bit-exact, losses, Adam, re-projection) re-run by
the oracle from the same initial parameters and update key.
"""
import os

import numpy as np
import pytest
import torch

from oracle import algo_common as oac
from oracle import env as oenv
from oracle import layouts, metrics as omet, nn as onn, ppo as oppo, prng
from oracle.moving_avg import EMANormalizer as OEMA

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
BUCKETS = [4, 8, 5, 5, 2, 2]


def _make(mlb, N=32, T=32, C=1, M=8, E=2, H=64, L=3, D=16, normalize_values=False, clipv=False,
          huber=False, seed=5, p_done=-1.0, lr=3e-4, separate=False):
    m = mlb
    env = m.SyntheticVectorEnv(N, D, len(BUCKETS), seed=11, p_done=p_done, device=DEV)
    enc = lambda: m.BackboneEncoder(net=m.models.MLP(H, L))
    policy = m.Policy(actor_critic=m.ActorCritic(
        backbone=(m.BackboneSeparate(prefix=None, actor_encoder=enc(), critic_encoder=enc()) if separate
                  else m.BackboneShared(prefix=None, encoder=enc())),
        actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(BUCKETS)),
        critic=m.models.DenseLayerCritic()))
    cfg = m.TrainConfig(
        num_worlds=N, num_agents_per_world=1, num_updates=10,
        actions={'act': m.DiscreteActionsConfig(BUCKETS)}, steps_per_update=T, lr=lr,
        algo=m.PPOConfig(num_epochs=E, minibatch_size=M, clip_coef=0.2, value_loss_coef=0.5,
                         entropy_coef={'act': 0.01}, max_grad_norm=0.5, clip_value_loss=clipv,
                         huber_value_loss=huber),
        num_bptt_chunks=C, gamma=0.99, seed=seed, metrics_buffer_size=4, gae_lambda=0.95,
        dreamer_v3_critic=False, normalize_values=normalize_values)
    mgr = m.init_training(DEV, cfg, env.sim_fns(), policy, None, verbose=False)
    return mgr, cfg, env


def _ocfg(cfg):
    a = cfg.algo
    return oppo.PPOCfg(BUCKETS, num_epochs=a.num_epochs, minibatch_size=a.minibatch_size,
                       clip_coef=a.clip_coef, value_loss_coef=a.value_loss_coef,
                       entropy_coef=a.entropy_coef['act'], max_grad_norm=a.max_grad_norm, lr=cfg.lr,
                       clip_value_loss=a.clip_value_loss, huber_value_loss=a.huber_value_loss,
                       normalize_values=cfg.normalize_values, gamma=cfg.gamma,
                       gae_lambda=cfg.gae_lambda)


def _vn_to_oracle(t):
    h = t.cpu().numpy()
    return dict(mu=h[0:1].copy(), inv_sigma=h[1:2].copy(), sigma=h[2:3].copy(), mu_biased=h[3:4].copy(),
                sigma_sq_biased=h[4:5].copy(), N=np.int32(h[5:6].view(np.int32)[0]))


@pytest.mark.parametrize('kw', [
    dict(),                                                        # config 1: 32 worlds x 32 steps
    dict(N=48, T=24, C=3, M=36, E=2, normalize_values=True),       # BPTT chunks + value-norm EMA
    dict(N=64, T=8, M=16, E=3, clipv=True, huber=True, p_done=0.1),
    # non-trivial value-normaliser state: the GAE input, the finish_rollouts hook arguments and
    # the 'Bootstrap Values' metric are in UN-normalised units (ml/rollouts.py:726-745, 810)
    dict(N=48, T=24, C=3, M=36, E=2, normalize_values=True, vn_init=(0.7, 2.5)),
    # BackboneSeparate (ml/actor_critic.py:247-303): actor and critic towers, block-diagonal fused head
    dict(N=48, T=16, M=24, E=2, L=2, separate=True, p_done=0.05),
])
def test_update_iter_matches_oracle(mlb, kw, monkeypatch):
    monkeypatch.setenv('MLB_CUDA_GRAPH', '0')
    kw = dict(kw)
    vn_init = kw.pop('vn_init', None)
    mgr, cfg, env = _make(mlb, **kw)
    if vn_init is not None:
        mu, sigma = vn_init
        corr = 1.0 - cfg.value_normalizer_decay ** 10
        with torch.no_grad():
            vs = mgr.state.train_states.value_normalizer_state
            vs[:5] = torch.tensor([mu, 1.0 / sigma, sigma, mu * corr, sigma * sigma * corr], device=DEV)
            vs[5:].view(torch.int32).fill_(10)
    prog = mgr.state.policy_states.program
    N, T, C = cfg.num_worlds, cfg.steps_per_update, cfg.num_bptt_chunks
    Tp, D, A = T // C, prog.obs_dim, len(BUCKETS)
    p0 = prog.to_oracle_params()
    key0 = mgr.state.train_states.update_prng_key.cpu().numpy().view(np.uint32).copy()
    vn0 = _vn_to_oracle(mgr.state.train_states.value_normalizer_state) if cfg.normalize_values else None
    obs0 = mgr.rollout.cur_obs['obs'].cpu().numpy().copy()

    mgr.update_iter()
    torch.cuda.synchronize()
    st = {k: v.cpu().numpy() for k, v in mgr.rollout_mgr.store.items()}
    boot = mgr.rollout_mgr.bootstrap.cpu().numpy()

    # (a) environment transitions, bit-exact replay from the stored actions
    ref_env = oenv.SyntheticEnv(N, D, A, seed=11, p_done=env.p_done)
    np.testing.assert_array_equal(obs0, ref_env.obs)
    obs_seq = st['obs'].reshape(T, N, D)
    act_seq = st['actions'].reshape(T, N, A)
    for t in range(T):
        np.testing.assert_array_equal(obs_seq[t], ref_env.obs)
        _, r, d = ref_env.step(act_seq[t])
        np.testing.assert_array_equal(st['rewards'].reshape(T, N)[t], r)
        np.testing.assert_array_equal(st['dones'].reshape(T, N)[t], d)
    np.testing.assert_array_equal(mgr.rollout.cur_obs['obs'].cpu().numpy(), ref_env.obs)

    # (b) stored log-probs / values == oracle forward of the initial parameters
    p64 = onn.cast_tree(p0, np.float64)
    logits, critic, _ = onn.actor_critic_fwd(p64, obs_seq.reshape(T * N, D).astype(np.float64))
    lp, _ = onn.action_stats(logits, act_seq.reshape(T * N, A), BUCKETS)
    np.testing.assert_allclose(st['log_probs'].reshape(T * N, A), lp, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(st['values'].reshape(T * N, 1), critic, rtol=1e-4, atol=1e-5)
    _, bcrit, _ = onn.actor_critic_fwd(p64, ref_env.obs.astype(np.float64))
    np.testing.assert_allclose(boot.reshape(N, 1), bcrit, rtol=1e-4, atol=1e-5)

    # (c) GAE + returns on the stored buffers: bit-exact
    vals, bvals = st['values'], boot
    if cfg.normalize_values:
        on = OEMA(cfg.value_normalizer_decay)
        vals, bvals = on.invert(vn0, st['values']), on.invert(vn0, boot)
    adv = oac.compute_advantages(cfg.gamma, cfg.gae_lambda, st['rewards'], vals, st['dones'], bvals)
    np.testing.assert_array_equal(st['advantages'], adv)
    np.testing.assert_array_equal(st['returns'], (adv + vals).astype(np.float32))

    # (d) the PPO update re-run by the oracle from (p0, key0)
    ocfg = _ocfg(cfg)
    roll = {k: layouts.reorder_seq_data(st[k])[0] for k in
            ('obs', 'actions', 'log_probs', 'advantages', 'returns', 'values', 'dones')}
    opt = oppo.adam_init(p0)
    norms = oppo.initial_weight_norms(p0)
    p1, opt1, key1, vn1, last, perms = oppo.ppo_update(p0, opt, norms, roll, ocfg, key0, vn0,
                                                       dtype=np.float32)
    np.testing.assert_array_equal(mgr.ppo_ws.perm.cpu().numpy(), perms)           # bit-exact
    np.testing.assert_array_equal(
        mgr.state.train_states.update_prng_key.cpu().numpy().view(np.uint32), key1)
    got = prog.to_oracle_params()
    num = sum(float(np.sum(np.square(a.astype(np.float64) - b))) for a, b in
              zip(onn.tree_leaves(got), onn.tree_leaves(p1)))
    den = sum(float(np.sum(np.square(a.astype(np.float64) - b))) for a, b in
              zip(onn.tree_leaves(p1), onn.tree_leaves(p0)))
    assert den > 0
    assert np.sqrt(num / den) < 2e-2, f'parameter-delta rel-L2 {np.sqrt(num / den)}'
    onn.tree_map(lambda a, b: np.testing.assert_allclose(a, b, atol=4 * cfg.lr), got, p1)
    assert prog.adam_step.item() == opt1['t']
    if cfg.normalize_values:
        g = _vn_to_oracle(mgr.state.train_states.value_normalizer_state)
        assert g['N'] == vn1['N']
        for k in ('mu', 'sigma', 'inv_sigma', 'mu_biased', 'sigma_sq_biased'):
            np.testing.assert_allclose(g[k], vn1[k], rtol=1e-4, atol=1e-7, err_msg=k)

    # (e) metrics ring: slot 0 holds this update's records
    lat = mgr.metrics.latest()
    for name, x in (('Rewards', st['rewards']), ('Values', vals), ('Est Returns', st['returns']),
                    ('Advantages', st['advantages']), ('Bootstrap Values', bvals)):
        e = omet.metric_from_data(x)
        m = lat[name]
        assert m.count == e['count'], name
        np.testing.assert_allclose([m.mean, m.min, m.max], [e['mean'], e['min'], e['max']],
                                   rtol=1e-4, atol=1e-6, err_msg=name)
        np.testing.assert_allclose(m.m2, e['m2'], rtol=1e-3, err_msg=name)
    np.testing.assert_allclose(lat['Loss'].mean, last['loss'], rtol=5e-3, atol=1e-5)
    np.testing.assert_allclose(lat['Entropy'].mean, np.mean(last['entropies']), rtol=1e-3)
    assert lat['Env Returns'].count == T * N
    assert mgr.update_idx == 1 and mgr.metrics.cur_buffer_offset == 1


def test_cuda_graph_replay_matches_eager(mlb, monkeypatch):
    """The captured update graph must do exactly what the eager path does."""
    res = {}
    for mode in ('0', '1'):
        monkeypatch.setenv('MLB_CUDA_GRAPH', mode)
        mgr, cfg, env = _make(mlb, N=64, T=16, M=16, E=2, seed=9)
        for _ in range(4):
            mgr.update_iter()
        torch.cuda.synchronize()
        assert (mgr._graph is not None) == (mode == '1')
        res[mode] = (mgr.state.policy_states.program.params.cpu().numpy().copy(),
                     mgr.rollout_mgr.store['actions'].cpu().numpy().copy(),
                     mgr.metrics.latest()['Loss'].mean)
    p_e, a_e, l_e = res['0']
    p_g, a_g, l_g = res['1']
    # split-K atomics reorder fp32 sums, so parameters agree to tolerance, not bitwise
    assert np.linalg.norm(p_e - p_g) / np.linalg.norm(p_e) < 1e-4
    assert (a_e == a_g).mean() > 0.98
    np.testing.assert_allclose(l_e, l_g, rtol=5e-2, atol=1e-4)


def test_policy_learns_synthetic_task(mlb, monkeypatch):
    """Sanity: PPO through the graph-replayed path improves the mean reward of the synthetic
    task (reward = obs0 * (a0 - 1.5) * 0.1 -> pick a0 by the sign of obs0)."""
    monkeypatch.setenv('MLB_CUDA_GRAPH', '1')
    mgr, cfg, env = _make(mlb, N=512, T=16, M=128, E=4, lr=3e-3, p_done=1 / 16, seed=1)
    rew = []
    for _ in range(40):
        mgr.update_iter()
        rew.append(mgr.metrics.latest()['Rewards'].mean)
    assert np.isfinite(rew).all()
    assert np.mean(rew[-5:]) > np.mean(rew[:5]) + 0.01, rew


def test_checkpoint_roundtrip(mlb, tmp_path, monkeypatch):
    monkeypatch.setenv('MLB_CUDA_GRAPH', '0')
    mgr, cfg, env = _make(mlb, N=32, T=8, M=8)
    mgr.update_iter()
    mgr.save_ckpt(str(tmp_path))
    p = mgr.state.policy_states.program.params.clone()
    k = mgr.state.train_states.update_prng_key.clone()
    mgr.update_iter()
    assert not torch.equal(p, mgr.state.policy_states.program.params)
    mgr.load_ckpt(os.path.join(str(tmp_path), '1'))
    assert torch.equal(p, mgr.state.policy_states.program.params)
    assert torch.equal(k, mgr.state.train_states.update_prng_key)
    assert mgr.update_idx == 1


def test_host_trace_env_delivers_the_trace(mlb):
    """HostTraceEnv (the e2e arm's simulator stand-in): one pinned record per step, one H2D copy;
    the device views must show exactly the host trace, step after step and across wrap-around."""
    env = mlb.HostTraceEnv(96, 5, obs_dim=16, seed=3, device=DEV)
    out = env.init()
    torch.cuda.synchronize()
    assert torch.equal(out['obs']['obs'].cpu(), env.h_obs0)
    for t in range(7):
        out = env.step({'actions': None})
        torch.cuda.synchronize()
        rec = env.h_rec[t % 5]
        N, D = 96, 16
        assert torch.equal(out['obs']['obs'].cpu().view(torch.uint8).flatten(), rec[:N * D * 4])
        assert torch.equal(out['rewards'].cpu().view(torch.uint8).flatten(), rec[N * D * 4:N * D * 4 + N * 4])
        assert torch.equal(out['dones'].cpu().flatten(), rec[N * D * 4 + N * 4:])
    assert env.h2d_bytes_per_update == 5 * (N * D * 4 + N * 5)


def test_checkpoint_roundtrip_and_reference_key_tree(mlb, tmp_path):
    """TrainStateManager.save / load (ml/train_state.py:145-196): top-level keys of the reference's
    checkpoint dict, the flax parameter key tree, and a bit-exact restore into a fresh manager via
    init_training(restore_ckpt=...)."""
    mgr, cfg, env = _make(mlb, N=64, T=8, M=32, E=1)
    for _ in range(2):
        mgr.update_iter()
    mgr.save_ckpt(str(tmp_path / 'ckpt'))                     # -> <dir>/<next_update>, like the reference
    path = str(tmp_path / 'ckpt' / str(int(mgr.update_idx)))
    ck = torch.load(path, map_location='cpu', weights_only=True)      # arrays / scalars only: no pickled code
    assert set(ck) == {'next_update', 'policy_states', 'train_states', 'pbt_rng', 'user_state'}
    # the reference's PolicyState / PolicyTrainState leaves (ml/train_state.py:34-47, 85-99)
    assert {'params', 'batch_stats', 'obs_preprocess_state', 'reward_hyper_params', 'episode_score', 'mmr'} <= set(ck['policy_states'])
    assert {'opt_state', 'value_normalizer_state', 'max_advantage_est_state', 'hyper_params', 'scaler',
            'update_prng_key', 'initial_weight_norms'} <= set(ck['train_states'])
    iwn = ck['train_states']['initial_weight_norms']
    assert iwn['backbone']['encoder']['net']['Dense_0']['kernel'] > 0 and iwn['actor']['impl']['kernel'] is None
    assert iwn['backbone']['encoder']['net']['LayerNorm_0']['impl']['scale'] is None
    tree = ck['policy_states']['params']
    assert set(tree) == {'backbone', 'actor', 'critic'}
    net = tree['backbone']['encoder']['net']
    assert {'Dense_0', 'LayerNorm_0'} <= set(net) and set(net['LayerNorm_0']['impl']) == {'scale', 'bias'}
    assert tree['actor']['impl']['kernel'].shape[1] == sum(BUCKETS) and torch.is_tensor(net['Dense_0']['kernel'])
    m = mlb
    env2 = m.SyntheticVectorEnv(64, 16, len(BUCKETS), seed=11, p_done=-1.0, device=DEV)
    policy = m.Policy(actor_critic=m.ActorCritic(
        backbone=m.BackboneShared(prefix=None, encoder=m.BackboneEncoder(net=m.models.MLP(64, 3))),
        actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(BUCKETS)),
        critic=m.models.DenseLayerCritic()))
    new = m.init_training(DEV, cfg, env2.sim_fns(), policy, None, restore_ckpt=path, verbose=False)
    a, b = mgr.state.policy_states.program, new.state.policy_states.program
    for x, y in ((a.params, b.params), (a.adam_m, b.adam_m), (a.adam_v, b.adam_v), (a.adam_step, b.adam_step),
                 (mgr.state.train_states.update_prng_key, new.state.train_states.update_prng_key)):
        assert torch.equal(x, y)
    assert new.update_idx == mgr.update_idx
    new.update_iter()          # the restored manager trains (segment table / bf16 copies rebuilt)
    torch.cuda.synchronize()


def test_finish_rollouts_hook_sees_unnormalized_values(mlb, monkeypatch):
    """with normalize_values the hook receives value_normalizer.invert(values / bootstrap)
    (ml/rollouts.py:726-745), not the stored normalised critic outputs."""
    import dataclasses
    monkeypatch.setenv('MLB_CUDA_GRAPH', '0')
    m = mlb
    seen = {}

    @dataclasses.dataclass(frozen=True)
    class Hooks(m.TrainHooks):
        def finish_rollouts(self, rollouts, bootstrap_values, unnormalized_values,
                            unnormalized_bootstrap_values, user_state):
            seen['v'] = unnormalized_values.clone()
            seen['b'] = unnormalized_bootstrap_values.clone()
            seen['raw_v'] = rollouts['values'].clone()
            seen['raw_b'] = bootstrap_values.clone()
            return rollouts, user_state

    env = m.SyntheticVectorEnv(32, 16, len(BUCKETS), seed=11, device=DEV)
    policy = m.Policy(actor_critic=m.ActorCritic(
        backbone=m.BackboneShared(prefix=None, encoder=m.BackboneEncoder(net=m.models.MLP(64, 2))),
        actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(BUCKETS)),
        critic=m.models.DenseLayerCritic()))
    cfg = m.TrainConfig(
        num_worlds=32, num_agents_per_world=1, num_updates=10,
        actions={'act': m.DiscreteActionsConfig(BUCKETS)}, steps_per_update=8, lr=3e-4,
        algo=m.PPOConfig(num_epochs=1, minibatch_size=16, clip_coef=0.2, value_loss_coef=0.5,
                         entropy_coef={'act': 0.01}, max_grad_norm=0.5),
        num_bptt_chunks=1, gamma=0.99, seed=5, metrics_buffer_size=4, gae_lambda=0.95,
        dreamer_v3_critic=False, normalize_values=True)
    mgr = m.init_training(DEV, cfg, env.sim_fns(), policy, None, user_hooks=Hooks(), verbose=False)
    with torch.no_grad():
        vs = mgr.state.train_states.value_normalizer_state
        vs[0], vs[1], vs[2] = -1.25, 1.0 / 3.0, 3.0
    mgr.update_iter()
    torch.cuda.synchronize()
    f = np.float32
    np.testing.assert_allclose(seen['v'].cpu().numpy(), seen['raw_v'].cpu().numpy() * f(3.0) + f(-1.25), rtol=1e-6)
    np.testing.assert_allclose(seen['b'].cpu().numpy(), seen['raw_b'].cpu().numpy() * f(3.0) + f(-1.25), rtol=1e-6)
    lat = mgr.metrics.latest()
    np.testing.assert_allclose(lat['Bootstrap Values'].mean, seen['b'].mean().item(), rtol=1e-5)
