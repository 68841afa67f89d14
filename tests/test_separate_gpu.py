"""GPU parity of BackboneSeparate (ml/actor_critic.py:247-303): two encoder towers on the same observations, the
actor head on the actor tower's features and the critic head on the critic tower's.  The lowering runs ONE head
GEMM over [actor features | critic features] with a block-diagonal weight; the tests check the forward / PPO loss /
backward against the oracle's two-tower restatement (fp32 and bf16 paths), that the off-diagonal head blocks get
no gradient and stay exactly zero through optimiser steps, and the reference parameter-tree shape.  The whole
update_iter against the oracle is the `separate=True` case of tests/test_train_gpu.py."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import algo_common as oac
from oracle import nn as onn, ppo as oppo

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
BUCKETS = [4, 8, 5, 5, 2, 2]


def _rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30)


def _policy(m, H, L):
    enc = lambda: m.BackboneEncoder(net=m.models.MLP(H, L))
    return m.Policy(actor_critic=m.ActorCritic(
        backbone=m.BackboneSeparate(prefix=None, actor_encoder=enc(), critic_encoder=enc()),
        actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(BUCKETS)), critic=m.models.DenseLayerCritic()))


@pytest.mark.parametrize('dtype,H,L', [(torch.float32, 64, 2), (torch.bfloat16, 64, 2), (torch.bfloat16, 256, 3)])
def test_separate_loss_and_grads_vs_oracle(mlb, dtype, H, L):
    from madrona_learn_b200 import _lib
    from madrona_learn_b200._lib import c_float, c_int, c_ll, c_size_t, call, ptr
    from madrona_learn_b200.engine import PolicyProgram
    m = mlb
    D, Tp, M = 32, 4, 256
    rows, A = Tp * M, len(BUCKETS)
    nA = sum(BUCKETS)
    rng = np.random.default_rng(31)
    p = onn.init_params(rng, D, H, L, BUCKETS, separate=True)
    p['actor']['kernel'] = (rng.standard_normal(p['actor']['kernel'].shape) * 0.2).astype(np.float32)
    p['critic']['bias'] = np.array([0.3], np.float32)
    prog = PolicyProgram(_policy(m, H, L).actor_critic, D, {'act': m.DiscreteActionsConfig(BUCKETS)}, DEV, dtype)
    assert prog.NT == 2 and prog.feat == 2 * H and not prog.fused_rollout
    prog.load_oracle_params(p)
    t = prog.param_tree()
    assert set(t['backbone']) == {'actor_encoder', 'critic_encoder'}
    assert tuple(t['actor']['impl']['kernel'].shape) == (H, nA) and tuple(t['critic']['Dense_0']['kernel'].shape) == (H, 1)
    back = prog.to_oracle_params()
    onn.tree_map(lambda a, b: np.testing.assert_array_equal(a, b), back, p)
    cfg = oppo.PPOCfg(BUCKETS, entropy_coef=0.02)
    mb = dict(obs=rng.standard_normal((Tp, M, D)).astype(np.float32),
              actions=np.stack([rng.integers(0, b, (Tp, M)) for b in BUCKETS], -1).astype(np.int32),
              advantages=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              returns=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              values=rng.standard_normal((Tp, M, 1)).astype(np.float32), mb_weights=np.ones((M, 1), np.float32))
    mb['log_probs'] = (-np.abs(rng.standard_normal((Tp, M, A))) - 0.5).astype(np.float32)
    quant = onn.bf16_round if dtype == torch.bfloat16 else None
    ref = oppo.ppo_loss(p, mb, cfg, None, dtype=np.float64, quant=quant)
    dv = {k: torch.from_numpy(v).to(DEV) for k, v in mb.items()}
    obs_d = dv['obs'].view(rows, D)
    head = prog.forward_train(obs_d, rows)
    tol = 3e-3 if dtype == torch.bfloat16 else 1e-4
    h = head.cpu().numpy()
    assert _rel(h[:, :nA], ref['logits']) < tol
    assert _rel(h[:, nA:nA + 1], ref['critic']) < tol
    # rollout-side forward (separate workspaces, ping-pong buffers per tower) gives the same head
    hi = prog.forward_infer(obs_d, rows).cpu().numpy()
    np.testing.assert_allclose(hi[:, :nA + 1], h[:, :nA + 1], rtol=1e-5, atol=1e-5)
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
         ptr(tw['stats_out']), ptr(tw['loss_ws']), c_size_t(tw['loss_ws'].numel()), prog._bins_c, c_int(1))
    stt = _lib.PPOStats.from_buffer_copy(tw['stats_out'].cpu().numpy().tobytes())
    np.testing.assert_allclose(stt.loss, ref['loss'], rtol=10 * tol, atol=1e-5)
    prog.backward(obs_d, rows)
    torch.cuda.synchronize()
    g = prog.to_oracle_params(prog.grads)
    gt = 4e-2 if dtype == torch.bfloat16 else 2e-4
    onn.tree_map(lambda a, b: np.testing.assert_array_less(_rel(a, b), gt), g, ref['grads'])
    # off-diagonal blocks of the fused head: no gradient, weights exactly zero -- also after optimiser steps
    gW, _ = prog.head_views(prog.grads)
    assert float(gW[:H, nA:].abs().max()) == 0.0 and float(gW[H:, :nA].abs().max()) == 0.0
    assert float(gW[:H, :nA].abs().max()) > 0 and float(gW[H:, nA:nA + 1].abs().max()) > 0
    for _ in range(3):
        prog.optimizer_step(1e-3, 0.5)
        prog.zero_grads()
        head = prog.forward_train(obs_d, rows)
        call('mlb_ppo_loss_f32', ptr(head), c_int(prog.NH), ptr(dv['actions']), ptr(dv['log_probs']),
             ptr(dv['advantages']), ptr(dv['returns']), ptr(None), ptr(None), ptr(adv_mr), ptr(None),
             prog._buckets_c, obj_scale, ent_scale, c_int(A), c_ll(rows), c_ll(M), c_float(cfg.clip_coef),
             c_float(cfg.value_loss_coef), c_int(prog.loss_flags), ptr(tw['dhead']), ptr(prog.head_bias_grad()),
             ptr(tw['stats_out']), ptr(tw['loss_ws']), c_size_t(tw['loss_ws'].numel()), prog._bins_c, c_int(1))
        prog.backward(obs_d, rows)
    torch.cuda.synchronize()
    W, _ = prog.head_views(prog.params)
    assert float(W[:H, nA:].abs().max()) == 0.0 and float(W[H:, :nA].abs().max()) == 0.0
    if prog.tc:
        assert float(prog.wh_c[:H, nA:].float().abs().max()) == 0.0


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_separate_update_iter_runs(mlb, dtype):
    """BackboneSeparate end to end through the public API (layer-by-layer rollout, GAE, update, graph replay) and
    the checkpoint tree carries both encoders."""
    m = mlb
    N, T, D = 256, 16, 32
    env = m.SyntheticVectorEnv(N, D, len(BUCKETS), seed=3, device=DEV)
    cfg = m.TrainConfig(
        num_worlds=N, num_agents_per_world=1, num_updates=4, actions={'act': m.DiscreteActionsConfig(BUCKETS)},
        steps_per_update=T, lr=3e-4,
        algo=m.PPOConfig(num_epochs=2, minibatch_size=128, clip_coef=0.2, value_loss_coef=0.5,
                         entropy_coef={'act': 0.01}, max_grad_norm=0.5),
        num_bptt_chunks=1, gamma=0.99, seed=1, metrics_buffer_size=4, gae_lambda=0.95, dreamer_v3_critic=False,
        compute_dtype=dtype)
    mgr = m.init_training(DEV, cfg, env.sim_fns(), _policy(m, 64, 2), None, verbose=False)
    prog = mgr.state.policy_states.program
    p0 = prog.params.clone()
    v = []
    for i in range(4):
        mgr.update_iter()
        torch.cuda.synchronize()
        v.append(mgr.metrics.latest()['Value Loss'].mean)
        assert np.isfinite(mgr.metrics.latest()['Loss'].mean)
    assert all(np.isfinite(v))
    t0 = prog.param_tree(p0)
    t1 = prog.param_tree()
    for enc in ('actor_encoder', 'critic_encoder'):          # both towers moved
        assert not torch.equal(t0['backbone'][enc]['net']['Dense_0']['kernel'], t1['backbone'][enc]['net']['Dense_0']['kernel'])
    n = prog.initial_weight_norms_tree()
    assert n['backbone']['critic_encoder']['net']['Dense_1']['kernel'] > 0
