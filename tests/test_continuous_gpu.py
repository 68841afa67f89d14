"""GPU parity of the continuous action group (ContinuousActionDistributions, ml/dists.py:211-284): sampling
(threefry -> uniform -> erf_inv -> Normal), log-density / entropy in the PPO loss kernel and their gradient into
the raw mean / raw std columns.  Kernels vs the reference-generated golden (tests/golden/continuous.npz) and vs
the oracle through a full forward / loss / backward, plus update_iter end to end on a task whose optimum is
known (the policy's mean must move towards it)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import algo_common as oac
from oracle import dists as odists, nn as onn, ppo as oppo

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'continuous.npz')


def _rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30)


def _policy(m, H, L, ccfg):
    return m.Policy(actor_critic=m.ActorCritic(
        backbone=m.BackboneShared(prefix=None, encoder=m.BackboneEncoder(net=m.models.MLP(H, L))),
        actor=m.models.DenseLayerContinuousActor(ccfg), critic=m.models.DenseLayerCritic()))


@pytest.mark.parametrize('gi', [0, 1])
def test_continuous_kernels_vs_reference_golden(mlb, gi):
    """mlb_ppo_loss_f32 (MLB_PPO_CONTINUOUS_ACTIONS) and mlb_sample_continuous_f32 against action_stats / best as
    executed from the reference source: entropy / surrogate metrics, the `best` action, log-probabilities."""
    from madrona_learn_b200 import _lib
    from madrona_learn_b200._lib import c_float, c_int, c_ll, c_size_t, call, ptr
    g = np.load(GOLD)
    lo, hi = float(g['stddev_min'][gi]), float(g['stddev_max'][gi])
    means, stds, acts = g['means'][:, gi], g['stds'][:, gi], g['actions'][:, gi]
    rows, n = means.shape
    ld = 8
    head = np.zeros((rows, ld), np.float32)
    head[:, :n], head[:, n:2 * n] = means, stds
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(DEV)
    head_d = d(head)
    # deterministic sample = best(); its log-prob is the density at the mean
    a_out = torch.zeros(rows, n, dtype=torch.int32, device=DEV)
    lp_out = torch.zeros(rows, n, device=DEV)
    vals = torch.zeros(rows, device=DEV)
    call('mlb_sample_continuous_f32', ptr(head_d), c_int(ld), ptr(None), c_int(n), c_float(lo), c_float(hi),
         c_ll(rows), c_int(0), c_int(1), ptr(a_out), ptr(lp_out), ptr(vals), None, c_int(1))
    np.testing.assert_allclose(a_out.view(torch.float32).cpu().numpy(), g['best'][:, gi], rtol=2e-6, atol=1e-7)
    # loss kernel with old_lp = the reference's log_probs and advantage 1: ratio == 1 -> action objective mean 1,
    # the entropy metric is the reference's mean entropy
    buckets = (ctypes.c_int32 * n)(*([0] * n))
    obj = (ctypes.c_float * n)(*([1.0 / (rows * n)] * n))
    ent = (ctypes.c_float * (n + 2))(*([1.0 / (rows * n)] * n + [lo, hi]))
    dhead = torch.zeros(rows, ld, device=DEV)
    stats = torch.zeros(ctypes.sizeof(_lib.PPOStats), dtype=torch.uint8, device=DEV)
    ws = torch.zeros(_lib.lib().mlb_ppo_loss_workspace(rows) + 16, dtype=torch.uint8, device=DEV)
    dbias = torch.zeros(ld, device=DEV)
    ones = d(np.ones((rows, 1), np.float32))
    acts_d, olp_d = d(acts.view(np.int32)), d(g['log_probs'][:, gi])     # keep alive across the launch
    call('mlb_ppo_loss_f32', ptr(head_d), c_int(ld), ptr(acts_d), ptr(olp_d),
         ptr(ones), ptr(ones), ptr(None), ptr(None), ptr(None), ptr(None), buckets, obj, ent, c_int(n), c_ll(rows),
         c_ll(rows), c_float(0.2), c_float(0.0), c_int(16), ptr(dhead), ptr(dbias), ptr(stats), ptr(ws),
         c_size_t(ws.numel()), None, c_int(1))
    st = _lib.PPOStats.from_buffer_copy(stats.cpu().numpy().tobytes())
    np.testing.assert_allclose(st.metrics[4].mean, np.mean(g['entropies'][:, gi]), rtol=2e-5)
    np.testing.assert_allclose(st.entropy, np.mean(g['entropies'][:, gi]), rtol=2e-5)
    np.testing.assert_allclose(st.metrics[1].mean, 1.0, rtol=1e-4)          # exp(lp - reference lp) == 1
    np.testing.assert_allclose(st.action_obj, 1.0, rtol=1e-4)


@pytest.mark.parametrize('partitionable', [False, True])
def test_continuous_sample_vs_oracle(mlb, partitionable):
    """The sampling kernel's draws against the oracle's jax.random.normal restatement (same threefry counters;
    erf_inv / tanh / sigmoid agree to float rounding) -- and they are N(mean, std) distributed."""
    from madrona_learn_b200._lib import c_float, c_int, c_ll, call, ptr
    rng = np.random.default_rng(3)
    rows, n, ld = 4099, 5, 12
    head = rng.standard_normal((rows, ld)).astype(np.float32)
    key = np.array([0x1234, 0xBEEF], np.uint32)
    lo, hi = 0.1, 1.2
    a_ref, lp_ref = odists.continuous_sample(head[:, :n], head[:, n:2 * n], key, lo, hi, partitionable)
    a_out = torch.zeros(rows, n, dtype=torch.int32, device=DEV)
    lp_out = torch.zeros(rows, n, device=DEV)
    vals = torch.zeros(rows, device=DEV)
    head_d, key_d = torch.from_numpy(head).to(DEV), torch.from_numpy(key.view(np.int32)).to(DEV)
    call('mlb_sample_continuous_f32', ptr(head_d), c_int(ld), ptr(key_d), c_int(n), c_float(lo), c_float(hi), c_ll(rows),
         c_int(int(partitionable)), c_int(0), ptr(a_out), ptr(lp_out), ptr(vals), None, c_int(1))
    a = a_out.view(torch.float32).cpu().numpy()
    np.testing.assert_allclose(a, a_ref, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(lp_out.cpu().numpy(), lp_ref, rtol=1e-4, atol=2e-4)
    np.testing.assert_array_equal(vals.cpu().numpy(), head[:, 2 * n])
    mean, std = odists.continuous_params(head[:, :n], head[:, n:2 * n], lo, hi)
    z = (a - mean) / std
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.03


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_continuous_loss_and_grads_vs_oracle(mlb, dtype):
    from madrona_learn_b200 import _lib
    from madrona_learn_b200._lib import c_float, c_int, c_ll, c_size_t, call, ptr
    from madrona_learn_b200.engine import PolicyProgram
    m = mlb
    D, H, L, Tp, M, n = 32, 64, 2, 4, 256, 4
    rows, A = Tp * M, n
    lo, hi = 0.05, 1.5
    rng = np.random.default_rng(12)
    p = onn.init_params(rng, D, H, L, [2 * n])
    p['actor']['kernel'] = (rng.standard_normal(p['actor']['kernel'].shape) * 0.3).astype(np.float32)
    p['actor']['bias'] = (rng.standard_normal(2 * n) * 0.3).astype(np.float32)
    ccfg = m.ContinuousActionsConfig(stddev_min=lo, stddev_max=hi, num_dims=n)
    prog = PolicyProgram(_policy(m, H, L, ccfg).actor_critic, D, {'act': ccfg}, DEV, dtype)
    assert prog.continuous == (lo, hi, n) and prog.sumA == 2 * n
    prog.load_oracle_params(p)
    cfg = oppo.PPOCfg([None] * n, entropy_coef=0.02, continuous=(lo, hi))
    acts = rng.standard_normal((Tp, M, n)).astype(np.float32)
    mb = dict(obs=rng.standard_normal((Tp, M, D)).astype(np.float32), actions=acts.view(np.int32),
              advantages=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              returns=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              values=rng.standard_normal((Tp, M, 1)).astype(np.float32), mb_weights=np.ones((M, 1), np.float32))
    quant = onn.bf16_round if dtype == torch.bfloat16 else None
    # old log-probs near the new ones so that both clip branches are populated
    lg = oppo.ppo_loss(p, dict(mb, log_probs=np.zeros((Tp, M, n), np.float32)), cfg, None, dtype=np.float64,
                       quant=quant, want_grads=False)
    mb['log_probs'] = (lg['new_log_probs'].reshape(Tp, M, n) + rng.standard_normal((Tp, M, n)) * 0.25).astype(np.float32)
    ref = oppo.ppo_loss(p, mb, cfg, None, dtype=np.float64, quant=quant)
    dv = {k: torch.from_numpy(np.ascontiguousarray(v)).to(DEV) for k, v in mb.items()}
    obs_d = dv['obs'].view(rows, D)
    head = prog.forward_train(obs_d, rows)
    tol = 3e-3 if dtype == torch.bfloat16 else 1e-4
    tw = prog.train_ws(rows)
    mean, rstd = oac.zscore_stats(mb['advantages'])
    adv_mr = torch.tensor([mean, rstd, 0, 0], dtype=torch.float32, device=DEV)
    obj_scale = (ctypes.c_float * A)(*[1.0 / (rows * A)] * A)
    ent_scale = (ctypes.c_float * (A + 2))(*([cfg.entropy_coef / (rows * A)] * A + [lo, hi]))
    prog.zero_grads()
    call('mlb_ppo_loss_f32', ptr(head), c_int(prog.NH), ptr(dv['actions']), ptr(dv['log_probs']),
         ptr(dv['advantages']), ptr(dv['returns']), ptr(None), ptr(None), ptr(adv_mr), ptr(None),
         None, obj_scale, ent_scale, c_int(A), c_ll(rows), c_ll(M), c_float(cfg.clip_coef),
         c_float(cfg.value_loss_coef), c_int(prog.loss_flags), ptr(tw['dhead']), ptr(prog.head_bias_grad()),
         ptr(tw['stats_out']), ptr(tw['loss_ws']), c_size_t(tw['loss_ws'].numel()), prog._bins_c, c_int(1))
    stt = _lib.PPOStats.from_buffer_copy(tw['stats_out'].cpu().numpy().tobytes())
    np.testing.assert_allclose(stt.loss, ref['loss'], rtol=10 * tol, atol=1e-4)
    np.testing.assert_allclose(stt.entropy, cfg.entropy_coef * np.mean(ref['entropies']), rtol=10 * tol)
    dh = tw['dhead'].float().cpu().numpy()
    assert _rel(dh[:, :2 * n], ref['dlogits']) < (2e-2 if dtype == torch.bfloat16 else 2e-4)
    prog.backward(obs_d, rows)
    gr = prog.to_oracle_params(prog.grads)
    gt = 4e-2 if dtype == torch.bfloat16 else 3e-4
    onn.tree_map(lambda a, b: np.testing.assert_array_less(_rel(a, b), gt), gr, ref['grads'])


class _TargetEnv:
    """sim_fns of a one-step continuous control task on the device: reward = -mean((a - target(obs))^2) with
    target = 0.8 * tanh(obs[:, :n]); observations are redrawn every step from a fixed pool (torch ops only, so
    the update graph captures it)."""

    def __init__(self, N, D, n, seed):
        g = torch.Generator(device=DEV).manual_seed(seed)
        self.pool = torch.randn(64, N, D, device=DEV, generator=g)
        self.obs = self.pool[0].clone()
        self.t = torch.zeros((), dtype=torch.int64, device=DEV)
        self.rewards = torch.zeros(N, 1, device=DEV)
        self.dones = torch.ones(N, 1, dtype=torch.uint8, device=DEV)
        self.n = n

    def init(self):
        return {'state': None, 'obs': {'obs': self.obs}}

    def step(self, si):
        a = si['actions']['act']
        assert a.dtype == torch.float32
        tgt = 0.8 * torch.tanh(self.obs[:, :self.n])
        self.rewards.copy_(-((a.view(-1, self.n) - tgt) ** 2).mean(-1, keepdim=True))
        self.t.add_(1)
        self.obs.copy_(self.pool.index_select(0, (self.t % 64).view(1))[0])
        return {'state': None, 'obs': {'obs': self.obs}, 'rewards': self.rewards, 'dones': self.dones}

    def sim_fns(self):
        return {'init': self.init, 'step': self.step}


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_continuous_update_iter_learns(mlb, dtype):
    """ContinuousActionsConfig end to end through the public API (sampling in the rollout, float actions handed
    to the simulator, PPO on Normal log-densities, graph replay): the mean reward rises."""
    m = mlb
    N, T, D, n = 512, 8, 16, 3
    ccfg = m.ContinuousActionsConfig(stddev_min=0.05, stddev_max=1.0, num_dims=n)
    env = _TargetEnv(N, D, n, seed=5)
    cfg = m.TrainConfig(
        num_worlds=N, num_agents_per_world=1, num_updates=40, actions={'act': ccfg}, steps_per_update=T, lr=1e-3,
        algo=m.PPOConfig(num_epochs=4, minibatch_size=N // 2, clip_coef=0.2, value_loss_coef=0.5,
                         entropy_coef={'act': 0.0}, max_grad_norm=0.5),
        num_bptt_chunks=1, gamma=0.0, seed=1, metrics_buffer_size=4, gae_lambda=0.95, dreamer_v3_critic=False,
        compute_dtype=dtype)
    mgr = m.init_training(DEV, cfg, env.sim_fns(), _policy(m, 64, 2, ccfg), None, verbose=False)
    r = []
    for i in range(40):
        mgr.update_iter()
        torch.cuda.synchronize()
        r.append(float(mgr.rollout_mgr.store['rewards'].mean()))
        assert np.isfinite(mgr.metrics.latest()['Loss'].mean)
    assert np.mean(r[-5:]) > np.mean(r[:5]) + 0.05, (r[:5], r[-5:])


def test_continuous_recurrent_update_iter_runs(mlb):
    """Continuous action group behind a recurrent (LSTM) encoder with BPTT chunks and the value normaliser:
    the layer-by-layer rollout path, sequence forward / BPTT and the Normal-density loss compose (CUDA graph replay
    included) and the policy improves on the target task."""
    m = mlb
    N, T, D, n = 512, 8, 16, 3
    ccfg = m.ContinuousActionsConfig(stddev_min=0.05, stddev_max=1.0, num_dims=n)
    env = _TargetEnv(N, D, n, seed=6)
    pol = m.Policy(actor_critic=m.ActorCritic(
        backbone=m.BackboneShared(prefix=None, encoder=m.RecurrentBackboneEncoder(
            net=m.models.MLP(64, 1), rnn=m.rnn.LSTM(64, 1))),
        actor=m.models.DenseLayerContinuousActor(ccfg), critic=m.models.DenseLayerCritic()))
    cfg = m.TrainConfig(
        num_worlds=N, num_agents_per_world=1, num_updates=40, actions={'act': ccfg}, steps_per_update=T, lr=1e-3,
        algo=m.PPOConfig(num_epochs=4, minibatch_size=N, clip_coef=0.2, value_loss_coef=0.5,
                         entropy_coef={'act': 0.0}, max_grad_norm=0.5),
        num_bptt_chunks=2, gamma=0.0, seed=1, metrics_buffer_size=4, gae_lambda=0.95, dreamer_v3_critic=False,
        normalize_values=True)
    mgr = m.init_training(DEV, cfg, env.sim_fns(), pol, None, verbose=False)
    r = []
    for i in range(40):
        mgr.update_iter()
        torch.cuda.synchronize()
        r.append(float(mgr.rollout_mgr.store['rewards'].mean()))
        assert np.isfinite(mgr.metrics.latest()['Loss'].mean)
    assert np.mean(r[-5:]) > np.mean(r[:5]) + 0.03, (r[:5], r[-5:])
