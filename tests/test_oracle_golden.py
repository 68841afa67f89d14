"""CPU: the oracle restatement vs golden outputs produced by the REFERENCE'S OWN SOURCE run
under oracle/jax_shim (tests/golden/make_golden.py), plus the known-answer vectors we hold
for the third-party PRNG.  This is what pins the oracle (DESIGN.md "Oracle")."""
import json
import os

import numpy as np
import pytest

from oracle import algo_common as oac
from oracle import cgae, dists as odists, layouts, metrics as omet, nn as onn, prng
from oracle.moving_avg import EMANormalizer

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = ['small', 'chunks', 'ragged', 'nodone', 'alldone', 'one']


@pytest.mark.parametrize('name', CASES)
def test_gae_returns_zscore_match_reference(name):
    z = np.load(os.path.join(G, 'algo_common.npz'))
    g = lambda k: z[f'{name}/{k}']
    gamma, lam = g('cfg')
    adv = oac.compute_advantages(gamma, lam, g('r'), g('v'), g('d'), g('b'))
    np.testing.assert_array_equal(adv, g('adv'))                       # bit-exact
    np.testing.assert_array_equal(oac.compute_returns(gamma, g('r'), g('d'), g('b')), g('ret'))
    np.testing.assert_allclose(oac.zscore_data(g('adv')), g('z'), rtol=1e-5, atol=1e-6)
    # the C restatement agrees bit-for-bit too
    T = g('r').shape[0] * g('r').shape[1]
    a2, _ = cgae.gae(g('r').reshape(T, -1), g('v').reshape(T, -1), g('d').reshape(T, -1),
                     g('b').reshape(-1), gamma, lam)
    np.testing.assert_array_equal(a2.reshape(g('adv').shape), g('adv'))


def test_ema_normalizer_matches_reference():
    z = np.load(os.path.join(G, 'ema.npz'))
    vals, hist = z['vals'], z['hist']
    iters, batch, dims = vals.shape
    sub = 8
    norm = EMANormalizer(0.999)
    est = norm.init_estimates(dims)
    for i in range(iters):
        stats = norm.init_input_stats(est)
        for j in range(sub):
            stats = norm.update_input_stats(stats, j, vals[i].reshape(sub, batch // sub, dims)[j])
        est = norm.update_estimates(est, stats)
        got = np.concatenate([est[k] for k in ('mu', 'inv_sigma', 'sigma', 'mu_biased', 'sigma_sq_biased')] +
                             [stats[0], stats[1]])
        np.testing.assert_allclose(got, hist[i], rtol=2e-5, atol=1e-6)
    assert est['N'] == z['N']
    np.testing.assert_allclose(norm.invert(est, vals[5]), z['inverted'], rtol=1e-5)
    est2, normed = norm.normalize_and_update_estimates(est, vals[3])
    np.testing.assert_allclose(normed, z['normalized'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(est2['mu'], z['est2_mu'], rtol=2e-5)
    np.testing.assert_allclose(est2['inv_sigma'], z['est2_inv_sigma'], rtol=2e-5)


def test_value_normalizer_recurrence_matches_reference():
    z = np.load(os.path.join(G, 'ema_value_norm.npz'))
    norm = EMANormalizer(0.99999)
    e = norm.init_estimates(1)
    for i in range(z['rets'].shape[0]):
        e, nr = norm.normalize_and_update_estimates(e, z['rets'][i])
        got = [e['mu'][0], e['inv_sigma'][0], e['sigma'][0], e['mu_biased'][0], e['sigma_sq_biased'][0]]
        np.testing.assert_allclose(got, z['hist'][i], rtol=3e-5, atol=1e-7)
    np.testing.assert_allclose(nr, z['last_normalized'], rtol=1e-4, atol=1e-5)


def test_metric_matches_reference():
    z = np.load(os.path.join(G, 'metric.npz'))
    pack = lambda m: np.array([m['mean'], m['m2'], m['min'], m['max'], m['count']], np.float64)
    m1, m2 = omet.metric_from_data(z['x1']), omet.metric_from_data(z['x2'])
    np.testing.assert_allclose(pack(m1), z['m1'], rtol=1e-5)
    np.testing.assert_allclose(pack(m2), z['m2'], rtol=1e-5)
    np.testing.assert_allclose(pack(omet.metric_merge(m1, m2)), z['merged'], rtol=1e-5)


def test_action_stats_match_reference():
    z = np.load(os.path.join(G, 'dists.npz'))
    buckets = list(z['buckets'])
    lp, ent = onn.action_stats(z['logits'].astype(np.float64), z['actions'], buckets)
    np.testing.assert_allclose(lp, z['log_probs'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ent, z['entropies'], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(onn.best_actions(z['logits'], buckets), z['best'])


def test_twohot_matches_reference():
    z = np.load(os.path.join(G, 'twohot.npz'))
    np.testing.assert_allclose(odists.twohot_mean(z['logits']), z['mean'], rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(odists.twohot_loss(z['logits'], z['targets']), z['loss'], rtol=1e-5)


def test_minibatch_relayout_matches_reference():
    z = np.load(os.path.join(G, 'minibatch.npz'))
    idx = z['idx']
    for k in ('obs', 'rewards'):
        store = z[f'store_{k}']
        ref = z[f'mb_{k}']
        np.testing.assert_array_equal(layouts.minibatch({'x': layouts.reorder_seq_data(store)[0]}, idx)['x'], ref)
        np.testing.assert_array_equal(layouts.minibatch_from_store(store, idx), ref)     # fused path
    np.testing.assert_array_equal(layouts.reorder_rnn_data(z['rnn'])[0][idx], z['mb_rnn_start_states'])


def test_reorder_chunks_match_reference_and_roundtrip():
    z = np.load(os.path.join(G, 'reorder_chunks.npz'))
    for i in range(4):
        a = z[f'v{i}_in']
        P, C = 6, 4
        B = a.size // C + P - 1
        tp, ts = layouts.compute_reorder_chunks(a, P, C, B)
        np.testing.assert_array_equal(tp, z[f'v{i}_to_policy'])
        np.testing.assert_array_equal(ts, z[f'v{i}_to_sim'])
        # the property the reference's own test asserts (tests/test_rollouts.py:36-56)
        pb = np.where(tp < a.size, a[np.clip(tp, 0, a.size - 1)], -1)
        np.testing.assert_array_equal(pb.reshape(-1)[ts], a)


def test_threefry_known_answers():
    kat = json.load(open(os.path.join(G, 'threefry_kat.json')))
    for v in kat['threefry2x32']:
        y0, y1 = prng.threefry2x32(v['key'][0], v['key'][1], v['ctr'][0], v['ctr'][1])
        assert [int(y0), int(y1)] == v['out']
    np.testing.assert_array_equal(prng.split(prng.key(0)), kat['split_prngkey0'])


def test_permutation_properties():
    for part in (False, True):
        for n in (1, 2, 17, 1625, 1626, 5000):
            p = prng.permutation(prng.key(n), n, part)
            assert sorted(p.tolist()) == list(range(n))
        a = prng.permutation(prng.key(3), 100, part)
        b = prng.permutation(prng.key(4), 100, part)
        assert not np.array_equal(a, b)
    assert prng.shuffle_rounds(1625) == 1 and prng.shuffle_rounds(1626) == 2
    assert prng.shuffle_rounds(8192) == 2 and prng.shuffle_rounds(1 << 20) == 2


PPO_CASES = ['plain', 'clipv_huber', 'valuenorm', 'valuenorm_clipv', 'twohot', 'returns_only']


def _ppo_case(name):
    """(mb, cfg, vn_state, z) of one tests/golden/ppo_loss.npz case (the reference's own
    `_ppo_update`, ml/ppo.py:109-362, executed under oracle/jax_shim)."""
    from oracle import ppo as oppo
    z = np.load(os.path.join(G, 'ppo_loss.npz'))
    g = lambda k: z[f'{name}/{k}']
    clipv, huber, vn, twohot, use_adv = [int(x) for x in g('flags')]
    cfg = oppo.PPOCfg([4, 8, 5, 5, 2, 2], clip_coef=0.2, value_loss_coef=0.5, entropy_coef=0.02,
                      clip_value_loss=bool(clipv), huber_value_loss=bool(huber), normalize_values=bool(vn),
                      dreamer_v3_critic=bool(twohot), compute_advantages=bool(use_adv))
    M = g('actions').shape[1]
    mb = dict(actions=g('actions'), log_probs=g('old_log_probs'), advantages=g('advantages'),
              returns=g('returns'), values=g('old_values'), mb_weights=np.ones((M, 1), np.float32))
    vn_state = None
    if vn:
        b = g('vn_before')
        vn_state = dict(mu=np.float32(b[0:1]), inv_sigma=np.float32(b[1:2]), sigma=np.float32(b[2:3]),
                        mu_biased=np.float32(b[3:4]), sigma_sq_biased=np.float32(b[4:5]), N=np.int32(b[5]))
    return mb, cfg, vn_state, g


@pytest.mark.parametrize('name', PPO_CASES)
def test_ppo_loss_matches_reference(name):
    """oracle/ppo.ppo_loss_heads vs the reference's loss_fn (ml/ppo.py:129-262) on the same head
    outputs: the five recorded metrics and the value-normaliser state after the minibatch."""
    from oracle import ppo as oppo
    mb, cfg, vn_state, g = _ppo_case(name)
    Tp, M, A = mb['actions'].shape
    rows = Tp * M
    out = oppo.ppo_loss_heads(g('logits').reshape(rows, -1), g('critic').reshape(rows, -1), mb, cfg, vn_state,
                              dtype=np.float32, want_grads=False)
    np.testing.assert_allclose(out['loss'], g('loss'), rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(out['action_obj'], g('action_obj'), rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(out['value_losses'], g('value_loss').reshape(rows, 1), rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(np.abs(out['value_errs']), g('value_errs').reshape(rows, 1), rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(out['entropies'], g('entropy'), rtol=2e-5, atol=1e-6)
    if vn_state is not None:
        a = g('vn_after')
        nv = out['new_vn_state']
        got = [nv['mu'][0], nv['inv_sigma'][0], nv['sigma'][0], nv['mu_biased'][0], nv['sigma_sq_biased'][0], nv['N']]
        np.testing.assert_allclose(got, a, rtol=2e-5, atol=1e-7)


def test_reprojection_matches_reference():
    """oracle/ppo.optimizer_step's kernel re-projection + LayerNorm renorm vs ml/ppo.py:300-338."""
    from oracle import ppo as oppo
    z = np.load(os.path.join(G, 'ppo_loss.npz'))
    p = {'mlp': [{'kernel': z[f'reproj/Dense_{i}_in'], 'scale': z[f'reproj/LayerNorm_{i}_scale_in'],
                  'bias': z[f'reproj/LayerNorm_{i}_bias_in']} for i in range(2)]}
    zero = onn.tree_map(np.zeros_like, p)
    cfg = oppo.PPOCfg([2], lr=0.0, max_grad_norm=0.5)
    norms = {'mlp': [float(x) for x in z['reproj/norms']]}
    new_p, _, _ = oppo.optimizer_step(p, zero, oppo.adam_init(p), cfg, norms, np.float32)
    for i in range(2):
        np.testing.assert_allclose(new_p['mlp'][i]['kernel'], z[f'reproj/Dense_{i}_out'], rtol=2e-6)
        np.testing.assert_allclose(new_p['mlp'][i]['scale'], z[f'reproj/LayerNorm_{i}_scale_out'], rtol=2e-6)
        np.testing.assert_allclose(new_p['mlp'][i]['bias'], z[f'reproj/LayerNorm_{i}_bias_out'], rtol=2e-6)


@pytest.mark.parametrize('name', ['default', 'filter', 'importance'])
def test_minibatch_selection_matches_reference(name):
    """oracle/ppo's restatement of the three selection branches of _ppo (ml/ppo.py:374-488) vs the
    reference's own _ppo run under the shim: number of minibatches, the (trajectory, step) composition
    of every minibatch of every epoch, the minibatch weights, the EMAEstimate state, the key advance."""
    from oracle import ppo as oppo
    from oracle.moving_avg import EMAEstimate
    z = np.load(os.path.join(G, 'ppo_select.npz'))
    g = lambda k: z[f'{name}/{k}']
    J, Tp, M, E = [int(x) for x in g('dims')]
    adv, val, ret = g('advantages'), g('values'), g('returns')
    tag = np.arange(J * Tp, dtype=np.int32).reshape(J, Tp, 1)
    roll = dict(advantages=adv, values=val, returns=ret, tag=tag, dones=np.zeros((J, Tp, 1), bool))
    key = g('key0')
    b = g('est_before')
    est0 = dict(mu=np.float32([b[0]]), mu_biased=np.float32([b[1]]), N=np.int32(b[2]))
    valid = weights = nmb = None
    est1 = est0
    if name == 'filter':
        valid, nmb, est1 = oppo.select_filter_advantages(adv, est0, 0.9, M)
        roll = oppo.flatten_time(roll)
    elif name == 'importance':
        ks = prng.split(key, 2)                                  # gen_update_rnd (:429)
        key = ks[1]
        valid, weights, _ = oppo.select_importance(adv, val, ret, ks[0], 2, M)
        nmb = 2
    else:
        valid, nmb = np.arange(J, dtype=np.int32), J // M
    assert nmb == int(g('num_minibatches'))
    np.testing.assert_allclose([est1['mu'][0], est1['mu_biased'][0], est1['N']], g('est_after'), rtol=1e-6)
    w = np.ones((roll['dones'].shape[0], 1), np.float32) if weights is None else weights
    tags, ws = [], []
    for e in range(E):
        ks = prng.split(key, 2)
        rnd, key = ks[0], ks[1]
        inds = prng.permutation(rnd, valid)
        inds = inds[np.argsort(np.where(inds == -1, 1, 0), kind='stable')]
        for i in range(nmb):
            mi = inds[i * M:(i + 1) * M]
            tags.append(layouts.minibatch(roll, mi)['tag'].reshape(-1))
            ws.append(w[mi].reshape(-1))
    np.testing.assert_array_equal(np.stack(tags), g('mb_tags'))          # bit-exact composition
    np.testing.assert_allclose(np.stack(ws), g('mb_weights'), rtol=1e-5)
    np.testing.assert_array_equal(key, g('key1'))


def test_hlgauss_oracle_matches_reference_golden():
    """oracle/dists.py HL-Gauss restatement vs HLGaussDist / HLGaussCritic.create executed from the reference
    source (tests/golden/hlgauss.npz, ml/models.py:177-306)."""
    from oracle import dists
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'hlgauss.npz'))
    for name, nb, lo, hi in (('default', 127, -100, 100), ('small', 31, -10, 10)):
        c, b = dists.hlgauss_bins(nb, lo, hi)
        np.testing.assert_array_equal(c, d[name + '_centers'])
        np.testing.assert_array_equal(b, d[name + '_bounds'])
        np.testing.assert_allclose(dists.hlgauss_mean(d[name + '_logits'], c), d[name + '_mean'], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(dists.hlgauss_loss(d[name + '_logits'], d[name + '_targets'], c, b,
                                                      float(d[name + '_smoothness'])), d[name + '_loss'], rtol=2e-6)
        t = dists.hlgauss_target(d[name + '_targets'], c, b, float(d[name + '_smoothness']))
        np.testing.assert_allclose(t.sum(-1), 1.0, rtol=1e-5)


def test_continuous_oracle_matches_reference_golden():
    """oracle/dists.py ContinuousActionDistributions restatement vs the reference class executed under the shim
    (tests/golden/continuous.npz, ml/dists.py:211-284): two groups with different std ranges."""
    from oracle import dists
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'continuous.npz'))
    for gi in range(2):
        lo, hi = float(g['stddev_min'][gi]), float(g['stddev_max'][gi])
        lp, ent = dists.continuous_action_stats(g['means'][:, gi], g['stds'][:, gi], g['actions'][:, gi], lo, hi)
        np.testing.assert_allclose(lp, g['log_probs'][:, gi], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(ent, g['entropies'][:, gi], rtol=1e-6, atol=1e-6)
        mean, std = dists.continuous_params(g['means'][:, gi], g['stds'][:, gi], lo, hi)
        np.testing.assert_allclose(mean, g['best'][:, gi], rtol=1e-6, atol=1e-7)
        assert (std >= lo).all() and (std <= hi).all()
        # analytic gradient of the restatement vs central differences (fp64)
        rng = np.random.default_rng(gi)
        dlp, dent = rng.standard_normal(lp.shape), rng.standard_normal(lp.shape)
        m64, s64, a64 = (g[k][:, gi].astype(np.float64) for k in ('means', 'stds', 'actions'))
        dm, ds = dists.continuous_action_stats_bwd(m64, s64, a64, lo, hi, dlp, dent)
        f = lambda m_, s_: sum((w * x).sum() for w, x in zip(
            (dlp, dent), dists.continuous_action_stats(m_, s_, a64, lo, hi, dtype=np.float64)))
        e = 1e-6
        for (i, j) in ((0, 0), (5, 2), (63, 1)):
            dmn, dsn = np.zeros_like(m64), np.zeros_like(s64)
            dmn[i, j] = dsn[i, j] = e
            np.testing.assert_allclose((f(m64 + dmn, s64) - f(m64 - dmn, s64)) / (2 * e), dm[i, j], rtol=1e-5, atol=1e-8)
            np.testing.assert_allclose((f(m64, s64 + dsn) - f(m64, s64 - dsn)) / (2 * e), ds[i, j], rtol=1e-5, atol=1e-8)


def test_continuous_sample_is_standard_normal():
    """oracle continuous_sample (jax.random.normal restated: threefry bits -> uniform(-1, 1) -> sqrt2 * erf_inv):
    the draws are N(mean, std) -- moments and a KS distance -- and erfinv_xla inverts scipy's erf."""
    from oracle import dists
    import scipy.special, scipy.stats
    x = np.linspace(-0.999999, 0.999999, 20001).astype(np.float32)
    np.testing.assert_allclose(scipy.special.erf(dists.erfinv_xla(x).astype(np.float64)), x, atol=3e-7)
    rows, n = 8192, 3
    z = np.zeros((rows, n), np.float32)
    key = np.array([7, 9], np.uint32)
    a, lp = dists.continuous_sample(z, z - 2.0, key, 0.0, 2.0)            # mean 0, std 2 * sigmoid(0) = 1
    assert abs(a.mean()) < 0.02 and abs(a.std() - 1) < 0.02
    assert scipy.stats.kstest(a.ravel(), 'norm').statistic < 0.01
    np.testing.assert_allclose(lp, scipy.stats.norm.logpdf(a.astype(np.float64)), atol=1e-5)


def test_multi_policy_train_order_matches_reference():
    """The multi-policy learner's block layout (policy p owns rows [p * B, (p + 1) * B) of the simulator batch) is
    the reference's `_compute_sim_to_train_indices` for self-play-only matchmaking, executed from the reference
    source (tests/golden/multi_policy.npz, ml/rollouts.py:1053-1104); the refusals of everything else."""
    import dataclasses
    import madrona_learn_b200 as m
    from madrona_learn_b200.multi_policy import _check_pbt, sim_to_train_indices
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'multi_policy.npz'))
    for name in ('selfplay_p2', 'selfplay_p3_teams'):
        P, nt, ts, sp, cp, pp = (int(x) for x in g[name + '_cfg'])
        idx = g[name + '_idx']
        np.testing.assert_array_equal(idx, sim_to_train_indices(P, sp))
        assert int(g[name + '_nper']) == sp // P == idx.shape[1]
    # with cross-play / past-play batches the table is no longer a reshape (team 0 of those matches only):
    # that mode is refused, not approximated
    P, nt, ts, sp, cp, pp = (int(x) for x in g['mixed_p2_cfg'])
    assert g['mixed_p2_idx'].shape == (P, (sp + cp // nt + pp // nt) // P)
    assert not np.array_equal(g['mixed_p2_idx'].reshape(-1), np.arange(g['mixed_p2_idx'].size))
    # an explicit assignment vector: rows of policy p in simulator order
    a = np.array([1, 0, 1, 1, 0, 0], np.int32)
    np.testing.assert_array_equal(sim_to_train_indices(2, 6, a), [[1, 4, 5], [0, 2, 3]])
    B = [4, 3]
    mk = lambda **kw: m.TrainConfig(
        num_worlds=8, num_agents_per_world=2, num_updates=1, actions={'act': m.DiscreteActionsConfig(B)},
        steps_per_update=4, lr=1e-3, algo=m.PPOConfig(num_epochs=1, minibatch_size=4, clip_coef=0.2,
                                                      value_loss_coef=0.5, entropy_coef=0.01, max_grad_norm=0.5),
        num_bptt_chunks=1, gamma=0.99, seed=0, metrics_buffer_size=2,
        pbt=m.PBTConfig(**{**dict(num_teams=2, team_size=1, num_train_policies=2, num_past_policies=0,
                                  self_play_portion=1.0, cross_play_portion=0.0, past_play_portion=0.0), **kw}))
    assert _check_pbt(mk()) == 2
    for bad in (dict(num_past_policies=1), dict(self_play_portion=0.5, cross_play_portion=0.5),
                dict(self_play_portion=0.5, past_play_portion=0.5)):
        with pytest.raises(NotImplementedError):
            _check_pbt(mk(**bad))
    with pytest.raises(ValueError):
        _check_pbt(mk(num_train_policies=3))          # 8 worlds do not split into 3 blocks
    with pytest.raises(ValueError):
        _check_pbt(mk(team_size=2))                   # teams x team_size != agents per world
