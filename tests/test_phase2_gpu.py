"""GPU parity for the policy path (fp32): GEMM, LayerNorm+ReLU fwd/bwd, sampling, fused PPO
loss + head gradient, optimiser, synthetic env -- each against the oracle through the C-ABI.
Tolerances (stated): fp32 GEMM/LN rel 1e-4 vs float64 oracle; loss scalars rel 1e-4;
gradients rel-L2 1e-4; env bit-exact; Adam/renorm rel 1e-5.
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import env as oenv
from oracle import nn as onn
from oracle import ppo as oppo
from oracle import prng

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def _rel_l2(a, b):
    return np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize('M,N,K,ta,tb,splitk', [
    (300, 256, 64, 0, 0, 1), (1000, 28, 256, 0, 0, 1), (513, 200, 100, 0, 1, 1),
    (64, 256, 5000, 1, 0, 1), (256, 256, 4096, 1, 0, 8), (256, 28, 3000, 1, 0, 4),
    (2048, 512, 512, 0, 0, 1), (130, 70, 33, 1, 1, 1), (1, 1, 1, 0, 0, 1)])
def test_gemm(mlb, M, N, K, ta, tb, splitk):
    from madrona_learn_b200.engine import gemm
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    ref = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64)
    for acc in ([1] if splitk > 1 else [0, 1]):
        Cd = _dev(C0.copy())
        gemm(_dev(A), _dev(B), Cd, _dev(bias), M, N, K, A.shape[1], B.shape[1], N, ta, tb, acc, splitk)
        want = ref + bias + (C0 if acc else 0)
        assert _rel_l2(Cd.cpu().numpy(), want) < 2e-6


@pytest.mark.parametrize('rows,H', [(1, 4), (37, 64), (1000, 256), (513, 512), (64, 1024)])
def test_layernorm_relu_fwd_bwd(mlb, rows, H):
    from madrona_learn_b200._lib import c_int, c_ll, call, ptr
    rng = np.random.default_rng(rows + H)
    z = (rng.standard_normal((rows, H)) * 2 + 0.5).astype(np.float32)
    s = (1 + 0.2 * rng.standard_normal(H)).astype(np.float32)
    b = (0.2 * rng.standard_normal(H)).astype(np.float32)
    dy = rng.standard_normal((rows, H)).astype(np.float32)
    y_ref, cache = onn.layernorm_relu_fwd(z.astype(np.float64), s.astype(np.float64), b.astype(np.float64))
    dz_ref, ds_ref, db_ref = onn.layernorm_relu_bwd(dy.astype(np.float64), cache, s.astype(np.float64))
    zd, sd, bd, dyd = _dev(z), _dev(s), _dev(b), _dev(dy)
    y = torch.empty_like(zd)
    stats = torch.empty(rows, 2, device=DEV)
    call('mlb_ln_relu_fwd_f32', ptr(zd), ptr(sd), ptr(bd), ptr(y), ptr(stats), c_ll(rows), c_int(H))
    np.testing.assert_allclose(y.cpu().numpy(), y_ref, rtol=1e-4, atol=1e-5)
    dz = torch.empty_like(zd)
    ds = torch.zeros(H, device=DEV)
    db = torch.zeros(H, device=DEV)
    call('mlb_ln_relu_bwd_f32', ptr(dyd), ptr(zd), ptr(stats), ptr(sd), ptr(bd), ptr(dz), ptr(ds),
         ptr(db), c_ll(rows), c_int(H))
    assert _rel_l2(dz.cpu().numpy(), dz_ref) < 1e-5
    assert _rel_l2(ds.cpu().numpy(), ds_ref) < 1e-5
    assert _rel_l2(db.cpu().numpy(), db_ref) < 1e-5


def _program(mlb, D, H, L, buckets):
    m = mlb
    ac = m.ActorCritic(
        backbone=m.BackboneShared(prefix=None, encoder=m.BackboneEncoder(net=m.models.MLP(H, L))),
        actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(buckets)),
        critic=m.models.DenseLayerCritic())
    from madrona_learn_b200.engine import PolicyProgram
    return PolicyProgram(ac, D, {'act': m.DiscreteActionsConfig(buckets)}, DEV)


def _rand_params(rng, D, H, L, buckets):
    p = onn.init_params(rng, D, H, L, buckets)
    p['actor']['kernel'] = (rng.standard_normal(p['actor']['kernel'].shape) * 0.3).astype(np.float32)
    p['actor']['bias'] = (rng.standard_normal(p['actor']['bias'].shape) * 0.1).astype(np.float32)
    p['critic']['bias'] = np.array([0.05], np.float32)
    for l in p['mlp']:
        l['scale'] = (1 + 0.1 * rng.standard_normal(H)).astype(np.float32)
        l['bias'] = (0.1 * rng.standard_normal(H)).astype(np.float32)
    return p


def test_forward_and_sampling(mlb):
    buckets = [4, 8, 5, 5, 2, 2]
    D, H, L, N = 64, 256, 3, 4096
    rng = np.random.default_rng(0)
    p = _rand_params(rng, D, H, L, buckets)
    prog = _program(mlb, D, H, L, buckets)
    prog.load_oracle_params(p)
    obs = rng.standard_normal((N, D)).astype(np.float32)
    head = prog.forward_infer(_dev(obs), N).cpu().numpy()
    logits, critic, _ = onn.actor_critic_fwd(onn.cast_tree(p, np.float64), obs.astype(np.float64))
    np.testing.assert_allclose(head[:, :26], logits, rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(head[:, 26:27], critic, rtol=1e-4, atol=2e-5)
    assert np.all(head[:, 27:] == 0)
    key = prng.key(77)
    kd = torch.from_numpy(key.view(np.int32).copy()).to(DEV)
    out, _ = prog.apply_rollout(kd, (), {'obs': _dev(obs)})
    acts = out['actions']['act'].cpu().numpy()
    lps = out['log_probs']['act'].cpu().numpy()
    ref_a, ref_lp = onn.sample_actions(head[:, :26], key, buckets)
    # identical algorithm (threefry bits -> gumbel -> argmax); logf ulp differences may flip
    # a near-tie, so demand >= 99.9 % identical actions and exact log-probs where they agree
    same = acts == ref_a
    assert same.mean() > 0.999
    np.testing.assert_allclose(lps[same], ref_lp[same], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out['critic'].cpu().numpy(), head[:, 26:27])
    # distributional check: empirical frequencies of component 1 vs softmax probabilities
    l1 = head[:, 4:12].astype(np.float64)
    pr = np.exp(l1 - l1.max(1, keepdims=True))
    pr /= pr.sum(1, keepdims=True)
    freq = np.bincount(acts[:, 1], minlength=8) / N
    np.testing.assert_allclose(freq, pr.mean(0), atol=0.03)
    # deterministic (best) path
    out, _ = prog.apply_actor_only((), {'obs': _dev(obs)})
    np.testing.assert_array_equal(out['actions']['act'].cpu().numpy(), onn.best_actions(head[:, :26], buckets))


def test_rollout_key_chain(mlb):
    from madrona_learn_b200._lib import c_int, call, ptr
    key = prng.key(5)
    kd = torch.from_numpy(key.view(np.int32).copy()).to(DEV)
    pk = torch.zeros(2, dtype=torch.int32, device=DEV)
    k = key
    for _ in range(3):
        call('mlb_rollout_keys', ptr(kd), ptr(pk), c_int(0))
        ks = prng.split(k, 2)
        k, step_key = ks[0], ks[1]
        np.testing.assert_array_equal(kd.cpu().numpy().view(np.uint32), k)
        np.testing.assert_array_equal(pk.cpu().numpy().view(np.uint32), prng.split(step_key, 1)[0])


@pytest.mark.parametrize('clipv,huber,vn', [(False, False, False), (True, False, False),
                                            (False, True, True), (True, True, True)])
def test_ppo_loss_and_full_backward(mlb, clipv, huber, vn):
    from madrona_learn_b200 import _lib
    from madrona_learn_b200._lib import c_float, c_int, c_ll, c_size_t, call, ptr
    buckets = [4, 8, 5, 5, 2, 2]
    D, H, L, Tp, M = 32, 64, 3, 8, 96
    rows, A = Tp * M, len(buckets)
    rng = np.random.default_rng(3)
    p = _rand_params(rng, D, H, L, buckets)
    cfg = oppo.PPOCfg(buckets, clip_value_loss=clipv, huber_value_loss=huber, normalize_values=vn,
                      entropy_coef=0.02, clip_coef=0.2, value_loss_coef=0.5)
    mb = dict(
        obs=rng.standard_normal((Tp, M, D)).astype(np.float32),
        actions=np.stack([rng.integers(0, b, (Tp, M)) for b in buckets], -1).astype(np.int32),
        advantages=(rng.standard_normal((Tp, M, 1)) * 2 + 0.3).astype(np.float32),
        returns=(rng.standard_normal((Tp, M, 1)) * 3 + 1).astype(np.float32),
        values=rng.standard_normal((Tp, M, 1)).astype(np.float32),
        mb_weights=np.ones((M, 1), np.float32))
    # old log-probs near the new ones so that some ratios fall inside and some outside the clip
    lg, _, _ = onn.actor_critic_fwd(onn.cast_tree(p, np.float64), mb['obs'].reshape(rows, D).astype(np.float64))
    lp0, _ = onn.action_stats(lg, mb['actions'].reshape(rows, A), buckets)
    mb['log_probs'] = (lp0 + 0.25 * rng.standard_normal(lp0.shape)).reshape(Tp, M, A).astype(np.float32)
    from oracle.moving_avg import EMANormalizer
    vn_state = None
    if vn:
        norm = EMANormalizer(cfg.value_normalizer_decay)
        vn_state = norm.init_estimates(1)
        vn_state = norm.update_estimates(vn_state, (np.array([0.4], np.float32), np.array([2.0], np.float32)))
    ref = oppo.ppo_loss(p, mb, cfg, vn_state, dtype=np.float64)

    prog = _program(mlb, D, H, L, buckets)
    prog.load_oracle_params(p)
    obs_d = _dev(mb['obs'].reshape(rows, D))
    head = prog.forward_train(obs_d, rows)
    tw = prog.train_ws(rows)
    from oracle import algo_common as oac
    mean, rstd = oac.zscore_stats(mb['advantages'])
    adv_mr = _dev(np.array([mean, rstd, 0, 0], np.float32))
    vnp = None
    if vn:
        nv = ref['new_vn_state']
        vnp = _dev(np.array([vn_state['mu'][0], vn_state['sigma'][0], nv['mu'][0], nv['inv_sigma'][0]], np.float32))
    obj_scale = (ctypes.c_float * A)(*[1.0 / (rows * A)] * A)
    ent_scale = (ctypes.c_float * A)(*[cfg.entropy_coef / (rows * A)] * A)
    flags = (1 if clipv else 0) | (2 if huber else 0)
    dv = {k: _dev(mb[k]) for k in ('actions', 'log_probs', 'advantages', 'returns', 'values')}  # keep alive
    prog.zero_grads()          # the loss kernel accumulates the head bias gradients
    call('mlb_ppo_loss_f32', ptr(head), c_int(prog.NH), ptr(dv['actions']), ptr(dv['log_probs']),
         ptr(dv['advantages']), ptr(dv['returns']), ptr(dv['values']), ptr(None),
         ptr(adv_mr), ptr(vnp), prog._buckets_c, obj_scale, ent_scale, c_int(A), c_ll(rows), c_ll(M),
         c_float(cfg.clip_coef), c_float(cfg.value_loss_coef), c_int(flags | prog.loss_flags), ptr(tw['dhead']),
         ptr(prog.head_bias_grad()), ptr(tw['stats_out']), ptr(tw['loss_ws']), c_size_t(tw['loss_ws'].numel()), None, c_int(0))
    st = _lib.PPOStats.from_buffer_copy(tw['stats_out'].cpu().numpy().tobytes())
    np.testing.assert_allclose(st.loss, ref['loss'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(st.action_obj, np.mean(ref['action_obj']), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(st.value_loss, cfg.value_loss_coef * np.mean(ref['value_losses']), rtol=1e-4)
    np.testing.assert_allclose(st.entropy, cfg.entropy_coef * np.mean(ref['entropies']), rtol=1e-4)
    from oracle import metrics as omet
    for i, x in enumerate([np.array([ref['loss']]), ref['action_obj'], ref['value_losses'],
                           np.abs(ref['value_errs']), ref['entropies']]):
        e = omet.metric_from_data(x.astype(np.float32))
        g = st.metrics[i]
        assert g.count == e['count']
        np.testing.assert_allclose([g.mean, g.min, g.max], [e['mean'], e['min'], e['max']], rtol=2e-4, atol=1e-6)
        np.testing.assert_allclose(g.m2, e['m2'], rtol=1e-3, atol=1e-6)
    dhead = tw['dhead'].cpu().numpy()
    assert _rel_l2(dhead[:, :26], ref['dlogits']) < 1e-4
    assert _rel_l2(dhead[:, 26:27], ref['dcritic']) < 1e-4
    assert np.all(dhead[:, 27:] == 0)
    prog.backward(obs_d, rows)
    g = prog.to_oracle_params(prog.grads)
    onn.tree_map(lambda a, b: np.testing.assert_array_less(_rel_l2(a, b), 1e-4), g, ref['grads'])


def test_optimizer_matches_oracle(mlb):
    buckets = [3, 2]
    D, H, L = 8, 16, 2
    rng = np.random.default_rng(9)
    p = _rand_params(rng, D, H, L, buckets)
    prog = _program(mlb, D, H, L, buckets)
    prog.load_oracle_params(p)
    cfg = oppo.PPOCfg(buckets, lr=3e-3, max_grad_norm=0.5)
    norms = oppo.initial_weight_norms(p)
    opt = oppo.adam_init(p)
    pp = p
    for it in range(4):
        g = onn.tree_map(lambda a: (rng.standard_normal(a.shape) * (0.01 if it == 2 else 1.0)).astype(np.float32), p)
        # load grads into the arena through a scratch program layout
        q = _program(mlb, D, H, L, buckets)
        q.load_oracle_params(g)
        prog.grads.copy_(q.params)
        pp, opt, gn = oppo.optimizer_step(pp, g, opt, cfg, norms, np.float32)
        prog.optimizer_step(cfg.lr, cfg.max_grad_norm)
        got = prog.to_oracle_params()
        onn.tree_map(lambda a, b: np.testing.assert_allclose(a, b, rtol=2e-5, atol=1e-6), got, pp)
        np.testing.assert_allclose(np.sqrt(prog.grad_sumsq.item()), gn, rtol=1e-6)
    assert prog.adam_step.item() == 4


@pytest.mark.parametrize('p_done', [-1.0, 1.0 / 8])
def test_synthetic_env_bit_exact(mlb, p_done):
    N, D, A = 257, 12, 3
    env = mlb.SyntheticVectorEnv(N, D, A, seed=123, p_done=p_done, device=DEV)
    ref = oenv.SyntheticEnv(N, D, A, seed=123, p_done=p_done)
    out = env.init()
    np.testing.assert_array_equal(out['obs']['obs'].cpu().numpy(), ref.obs)
    rng = np.random.default_rng(0)
    for t in range(70):
        a = rng.integers(0, 4, (N, A)).astype(np.int32)
        out = env.step({'actions': {'act': _dev(a)}})
        o, r, d = ref.step(a)
        np.testing.assert_array_equal(out['obs']['obs'].cpu().numpy(), o)
        np.testing.assert_array_equal(out['rewards'].cpu().numpy()[:, 0], r)
        np.testing.assert_array_equal(out['dones'].cpu().numpy()[:, 0].astype(bool), d)
    assert d.any() or p_done < 0


def test_fused_optimizer_step_matches_three_launch_path(mlb):
    """mlb_optimizer_step_fused (one launch, device-wide barriers) == mlb_sumsq_f32 + mlb_adam_step_f32 +
    mlb_renorm_segments: same arithmetic, only the fp64 partial-sum grouping of the norms differs
    (rel 1e-6 on parameters / moments after 3 steps), on both the fp32 and the bf16-copy paths."""
    from madrona_learn_b200.engine import PolicyProgram
    m = mlb
    buckets = [4, 8, 5, 5, 2, 2]
    for dtype in (torch.float32, torch.bfloat16):
        progs = []
        for fused in (True, False):
            ac = m.ActorCritic(
                backbone=m.BackboneShared(prefix=None, encoder=m.BackboneEncoder(net=m.models.MLP(128, 2))),
                actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(buckets)),
                critic=m.models.DenseLayerCritic())
            p = PolicyProgram(ac, 32, {'act': m.DiscreteActionsConfig(buckets)}, DEV, dtype)
            p.init_params(3)
            p.finalize_params()
            assert p._fused_opt
            p._fused_opt = fused
            progs.append(p)
        g = torch.Generator(device=DEV).manual_seed(0)
        for step in range(3):
            grads = torch.randn(progs[0].num_params, device=DEV, generator=g) * (0.3 if step else 30.0)  # clip on/off
            for p in progs:
                p.grads.copy_(grads)
                p.optimizer_step(3e-3, 0.5)
        torch.cuda.synchronize()
        a, b = progs
        assert int(a.adam_step.item()) == int(b.adam_step.item()) == 3
        for x, y in ((a.params, b.params), (a.adam_m, b.adam_m), (a.adam_v, b.adam_v)):
            assert torch.allclose(x, y, rtol=1e-6, atol=1e-9), (x - y).abs().max().item()
        assert abs(a.grad_sumsq.item() - b.grad_sumsq.item()) <= 1e-12 * b.grad_sumsq.item()
        if dtype == torch.bfloat16:
            for i in range(a.L):
                assert torch.equal(a.w_t[i], b.w_t[i]) or (a.w_t[i].float() - b.w_t[i].float()).abs().max() < 1e-2
