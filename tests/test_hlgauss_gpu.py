"""GPU parity of the HL-Gauss critic (cfg.hlgauss_critic, ml/models.py:177-306, ml/ppo.py:178-185): the value
estimate is the softmax-weighted bin centre (symmetric summation), the value loss the cross-entropy against the
histogram of a Gaussian around the return.  Kernel vs the reference-generated golden (tests/golden/hlgauss.npz)
and vs the oracle through a full forward / loss / backward and update_iter."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import algo_common as oac
from oracle import dists as odists, nn as onn, ppo as oppo

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
BUCKETS = [4, 8, 5, 5, 2, 2]
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'hlgauss.npz')


def _rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30)


def _policy(m, H, L, **kw):
    return m.Policy(actor_critic=m.ActorCritic(
        backbone=m.BackboneShared(prefix=None, encoder=m.BackboneEncoder(net=m.models.MLP(H, L))),
        actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(BUCKETS)),
        critic=m.models.HLGaussCritic.create(**kw)))


@pytest.mark.parametrize('name,nb,lo,hi,sm', [('default', 127, -100, 100, 0.75), ('small', 31, -10, 10, 0.5)])
def test_hlgauss_kernel_vs_reference_golden(mlb, name, nb, lo, hi, sm):
    """mlb_ppo_loss_f32 (value-loss rows, value-error metric) and the sampling kernel's value decode against
    HLGaussDist.loss / .mean as executed from the reference source."""
    from madrona_learn_b200 import _lib
    from madrona_learn_b200._lib import c_float, c_int, c_ll, c_size_t, call, ptr
    g = np.load(GOLD)
    cr = mlb.models.HLGaussCritic.create(num_bins=nb, min_bound=lo, max_bound=hi, smoothness=sm)
    np.testing.assert_array_equal(cr.centers, g[f'{name}_centers'])
    np.testing.assert_array_equal(cr.bounds, g[f'{name}_bounds'])
    logits, tgt = g[f'{name}_logits'], g[f'{name}_targets']
    rows, A = logits.shape[0], 1
    ld = (2 + nb + 3) // 4 * 4
    head = np.zeros((rows, ld), np.float32)
    head[:, 2:2 + nb] = logits                                  # one 2-way action component in front
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(DEV)
    tab = np.concatenate([cr.centers, cr.bounds, [cr.smoothness]]).astype(np.float32)
    tab_c = (ctypes.c_float * tab.size)(*tab.tolist())
    buckets = (ctypes.c_int32 * 1)(2)
    one = (ctypes.c_float * 1)(0.0)
    head_d, dhead = d(head), torch.zeros(rows, ld, device=DEV)
    stats = torch.zeros(ctypes.sizeof(_lib.PPOStats), dtype=torch.uint8, device=DEV)
    ws = torch.zeros(_lib.lib().mlb_ppo_loss_workspace(rows) + 16, dtype=torch.uint8, device=DEV)
    dbias = torch.zeros(ld, device=DEV)
    call('mlb_ppo_loss_f32', ptr(head_d), c_int(ld), ptr(d(np.zeros((rows, 1), np.int32))),
         ptr(d(np.full((rows, 1), np.log(0.5), np.float32))), ptr(d(np.zeros((rows, 1), np.float32))), ptr(d(tgt)),
         ptr(None), ptr(None), ptr(None), ptr(None), buckets, one, one, c_int(1), c_ll(rows), c_ll(rows),
         c_float(0.2), c_float(1.0), c_int(8), ptr(dhead), ptr(dbias), ptr(stats), ptr(ws), c_size_t(ws.numel()),
         tab_c, c_int(nb))
    st = _lib.PPOStats.from_buffer_copy(stats.cpu().numpy().tobytes())
    np.testing.assert_allclose(st.value_loss, np.mean(g[f'{name}_loss']), rtol=2e-5)
    np.testing.assert_allclose(st.metrics[3].mean, np.mean(np.abs(g[f'{name}_mean'] - tgt)), rtol=2e-5, atol=1e-5)
    # gradient rows: softmax - c, with c recovered from the oracle (itself pinned to the golden loss)
    c = odists.hlgauss_target(tgt, cr.centers, cr.bounds, cr.smoothness)
    sm_ = np.exp(logits - logits.max(-1, keepdims=True)); sm_ /= sm_.sum(-1, keepdims=True)
    np.testing.assert_allclose(dhead.cpu().numpy()[:, 2:2 + nb] * rows, sm_ - c, atol=2e-6)
    # value decode of the rollout sampling kernel
    vals = torch.zeros(rows, device=DEV)
    acts = torch.zeros(rows, 1, dtype=torch.int32, device=DEV)
    call('mlb_sample_discrete_f32', ptr(head_d), c_int(ld), ptr(None), buckets, c_int(1), c_ll(rows), c_int(0),
         c_int(1), ptr(acts), ptr(None), ptr(vals), tab_c, c_int(nb))
    np.testing.assert_allclose(vals.cpu().numpy(), g[f'{name}_mean'][:, 0], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_hlgauss_loss_and_grads_vs_oracle(mlb, dtype):
    from madrona_learn_b200._lib import c_float, c_int, c_ll, c_size_t, call, ptr
    from madrona_learn_b200.engine import PolicyProgram
    m = mlb
    D, H, L, Tp, M, V = 32, 64, 2, 4, 256, 127
    rows, A = Tp * M, len(BUCKETS)
    rng = np.random.default_rng(11)
    p = onn.init_params(rng, D, H, L, BUCKETS, critic_dim=V)
    p['actor']['kernel'] = (rng.standard_normal(p['actor']['kernel'].shape) * 0.2).astype(np.float32)
    p['critic']['kernel'] = (rng.standard_normal((H, V)) * 0.3).astype(np.float32)
    p['critic']['bias'] = (rng.standard_normal(V) * 0.1).astype(np.float32)
    pol = _policy(m, H, L)
    prog = PolicyProgram(pol.actor_critic, D, {'act': m.DiscreteActionsConfig(BUCKETS)}, DEV, dtype)
    assert prog.hlgauss and prog.V == V and not prog.twohot
    prog.load_oracle_params(p)
    cr = pol.actor_critic.critic
    cfg = oppo.PPOCfg(BUCKETS, entropy_coef=0.02, hlgauss=(cr.centers, cr.bounds, cr.smoothness))
    mb = dict(obs=rng.standard_normal((Tp, M, D)).astype(np.float32),
              actions=np.stack([rng.integers(0, b, (Tp, M)) for b in BUCKETS], -1).astype(np.int32),
              advantages=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              returns=(rng.standard_normal((Tp, M, 1)) * 40).astype(np.float32),
              values=rng.standard_normal((Tp, M, 1)).astype(np.float32), mb_weights=np.ones((M, 1), np.float32))
    mb['returns'][0, :4, 0] = [0.0, 1e7, -1e7, cr.bounds[40]]                 # clipping / bin-bound cases
    mb['log_probs'] = (-np.abs(rng.standard_normal((Tp, M, A))) - 0.5).astype(np.float32)
    quant = onn.bf16_round if dtype == torch.bfloat16 else None
    ref = oppo.ppo_loss(p, mb, cfg, None, dtype=np.float64, quant=quant)
    dv = {k: torch.from_numpy(v).to(DEV) for k, v in mb.items()}
    obs_d = dv['obs'].view(rows, D)
    head = prog.forward_train(obs_d, rows)
    tol = 3e-3 if dtype == torch.bfloat16 else 1e-4
    h = head.cpu().numpy()
    assert _rel(h[:, 26:26 + V], ref['critic']) < tol
    out, _ = prog.apply_rollout(torch.zeros(2, dtype=torch.int32, device=DEV), (), {'obs': obs_d})
    np.testing.assert_allclose(out['critic'].cpu().numpy(), odists.hlgauss_mean(h[:, 26:26 + V], cr.centers),
                               rtol=1e-3, atol=1e-3)
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
    dh = tw['dhead'].float().cpu().numpy()
    assert _rel(dh[:, 26:26 + V], ref['dcritic']) < (2e-2 if dtype == torch.bfloat16 else 1e-4)
    prog.backward(obs_d, rows)
    g = prog.to_oracle_params(prog.grads)
    gt = 4e-2 if dtype == torch.bfloat16 else 2e-4
    onn.tree_map(lambda a, b: np.testing.assert_array_less(_rel(a, b), gt), g, ref['grads'])


def test_hlgauss_update_iter_runs(mlb):
    """cfg.hlgauss_critic end to end through the public API (rollout value decode, GAE, update, graph replay)."""
    m = mlb
    N, T, D = 256, 16, 32
    env = m.SyntheticVectorEnv(N, D, len(BUCKETS), seed=3, device=DEV)
    cfg = m.TrainConfig(
        num_worlds=N, num_agents_per_world=1, num_updates=4, actions={'act': m.DiscreteActionsConfig(BUCKETS)},
        steps_per_update=T, lr=3e-4,
        algo=m.PPOConfig(num_epochs=2, minibatch_size=128, clip_coef=0.2, value_loss_coef=0.5,
                         entropy_coef={'act': 0.01}, max_grad_norm=0.5),
        num_bptt_chunks=1, gamma=0.99, seed=1, metrics_buffer_size=4, gae_lambda=0.95,
        dreamer_v3_critic=False, hlgauss_critic=True)
    mgr = m.init_training(DEV, cfg, env.sim_fns(), _policy(m, 64, 2, num_bins=63, min_bound=-20, max_bound=20), None,
                          verbose=False)
    l0 = None
    for i in range(4):
        mgr.update_iter()
        torch.cuda.synchronize()
        v = mgr.metrics.latest()['Value Loss'].mean
        assert np.isfinite(v)
        l0 = v if l0 is None else l0
    assert mgr.metrics.latest()['Value Loss'].mean < l0          # the critic is learning
    with pytest.raises(ValueError):                               # critic module and cfg flag must agree
        import dataclasses
        bad = dataclasses.replace(cfg, hlgauss_critic=False)
        m.init_training(DEV, bad, env.sim_fns(), _policy(m, 64, 2, num_bins=63), None, verbose=False).update_iter()
