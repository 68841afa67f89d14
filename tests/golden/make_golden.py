"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES from
/root/reference/src/madrona_learn under oracle/jax_shim (a NumPy stand-in for the jax API;
jax itself is not installable in this image).  Run once in the build container:

    python tests/golden/make_golden.py

The fixtures are committed; tests/test_oracle_golden.py checks the oracle restatement (and,
on the GPU box, tests/test_golden_gpu.py checks the CUDA kernels) against them.  Nothing here
runs on the GPU box (/root/reference does not exist there).
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.jax_shim import _arr, extract_function, install, load_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def ppo_update_golden(rng, jax, jnp, lax, ac, ma, di):
    """Runs the reference's OWN `_ppo_update` (ml/ppo.py:109-362, AST-extracted, unmodified) forward
    under the shim: loss_fn (:129-262), the kernel re-projection and the LayerNorm renorm
    (:300-338).  The policy forward is replaced by given head outputs (logits | value), optax's
    transformation by a zero update, so only reference code is exercised:
        inputs : head outputs, actions, old log-probs, advantages, returns, old values
        outputs: the five recorded metrics (:351-362), the new value-normaliser state, the
                 re-projected parameters."""
    import functools
    import types as _t
    from flax.core import FrozenDict
    f32 = np.float32
    buckets = [4, 8, 5, 5, 2, 2]
    A, sumA = len(buckets), sum(buckets)
    optax = _t.SimpleNamespace(
        l2_loss=lambda p, t: 0.5 * jnp.square(p - t),                       # optax 0.1.9 l2_loss
        huber_loss=lambda p, t, delta=1.0: (lambda e: 0.5 * jnp.square(jnp.minimum(jnp.abs(e), delta)) +
                                            delta * (jnp.abs(e) - jnp.minimum(jnp.abs(e), delta)))(p - t),
        apply_updates=lambda p, u: jax.tree.map(lambda a, b: a + b, p, u))

    def value_and_grad(fn, has_aux=True):
        return lambda params: (fn(params), jax.tree.map(lambda x: jnp.zeros_like(x), params))
    jax_ns = _t.SimpleNamespace(**{k: getattr(jax, k) for k in dir(jax) if not k.startswith('__')})
    jax_ns.value_and_grad = value_and_grad

    class Ctx:
        def __init__(self, *a): pass
        def __enter__(self): return self
        def __exit__(self, *a): return False
    ns = dict(jax=jax_ns, jnp=jnp, lax=lax, optax=optax, partial=functools.partial, profile=Ctx,
              zscore_data=ac.zscore_data, FrozenDict=FrozenDict, Any=object, TrainConfig=object,
              PolicyState=object, PolicyTrainState=object, TrainingMetrics=object)
    ppo_update = extract_function('ppo.py', '_ppo_update', ns)

    class Obj:
        def __init__(self, **kw): self.__dict__.update(kw)
        def update(self, **kw):
            d = dict(self.__dict__); d.update(kw); return Obj(**d)

    class Rec:
        def __init__(self): self.out = None
        def record(self, d): self.out = d; return self

    out = {}
    Tp, M = 5, 12
    H = 16
    for name, kw in {'plain': {}, 'clipv_huber': dict(clipv=True, huber=True),
                     'valuenorm': dict(vn=True), 'valuenorm_clipv': dict(vn=True, clipv=True),
                     'twohot': dict(twohot=True), 'returns_only': dict(use_adv=False)}.items():
        V = 63 if kw.get('twohot') else 1
        logits = (rng.standard_normal((Tp, M, sumA)) * 1.5).astype(f32)
        critic = (rng.standard_normal((Tp, M, V)) * (1.0 if V > 1 else 2.0)).astype(f32)
        acts = np.stack([rng.integers(0, b, (Tp, M)) for b in buckets], -1).astype(np.int32)
        dist = di.DiscreteActionDistributions(actions_num_buckets=buckets, all_logits=_arr(logits))
        lp_new, ent = dist.action_stats(_arr(acts))
        old_lp = (np.asarray(lp_new) + 0.25 * rng.standard_normal((Tp, M, A))).astype(f32)
        adv = (rng.standard_normal((Tp, M, 1)) * 2 + 0.3).astype(f32)
        ret = (rng.standard_normal((Tp, M, 1)) * (30 if V > 1 else 3) + 1).astype(f32)
        oldv = (critic[..., :1] + 0.3 * rng.standard_normal((Tp, M, 1))).astype(f32)
        mbw = np.ones((M, 1), f32)
        cfg = _t.SimpleNamespace(
            compute_advantages=kw.get('use_adv', True), normalize_advantages=True, normalize_returns=True,
            dreamer_v3_critic=bool(kw.get('twohot')), hlgauss_critic=False,
            algo=_t.SimpleNamespace(clip_value_loss=bool(kw.get('clipv')), huber_value_loss=bool(kw.get('huber')),
                                    entropy_coef={'act': 0.02}))
        vn = vn_state = None
        if kw.get('vn'):
            vn = ma.EMANormalizer(decay=0.99999, norm_dtype=jnp.float32, inv_dtype=jnp.float32)
            vn_state = vn.init_estimates(_arr(np.zeros((1, 1), f32)))
            for _ in range(3):          # a non-trivial state
                vn_state, _n = vn.normalize_and_update_estimates(
                    vn_state, _arr((rng.standard_normal((40, 1)) * 4 + 2).astype(f32)))
        params = {'backbone': {'encoder': {'net': {
            'Dense_0': {'kernel': _arr(rng.standard_normal((8, H)).astype(f32))},
            'LayerNorm_0': {'impl': {'scale': _arr((1 + 0.2 * rng.standard_normal(H)).astype(f32)),
                                     'bias': _arr((0.2 * rng.standard_normal(H)).astype(f32))}},
            'Dense_1': {'kernel': _arr(rng.standard_normal((H, H)).astype(f32))},
            'LayerNorm_1': {'impl': {'scale': _arr((1 + 0.2 * rng.standard_normal(H)).astype(f32)),
                                     'bias': _arr((0.2 * rng.standard_normal(H)).astype(f32))}}}}},
            'actor': {'impl': {'kernel': _arr(rng.standard_normal((H, sumA)).astype(f32)),
                               'bias': _arr(rng.standard_normal(sumA).astype(f32))}},
            'critic': {'Dense_0': {'kernel': _arr(rng.standard_normal((H, V)).astype(f32)),
                                   'bias': _arr(rng.standard_normal(V).astype(f32))}}}
        norms = {'backbone': {'encoder': {'net': {
            'Dense_0': {'kernel': f32(3.0)}, 'LayerNorm_0': {'impl': {'scale': None, 'bias': None}},
            'Dense_1': {'kernel': f32(5.5)}, 'LayerNorm_1': {'impl': {'scale': None, 'bias': None}}}}},
            'actor': {'impl': {'kernel': None, 'bias': None}}, 'critic': {'Dense_0': {'kernel': None, 'bias': None}}}
        crit_out = di.SymExpTwoHotDistribution.create(_arr(critic)) if V > 1 else _arr(critic)

        def apply_fn(variables, rnn, dones, actions, obs, train, method, mutable):
            return ({'log_probs': {'act': lp_new}, 'entropies': {'act': ent}, 'critic': crit_out},
                    {'batch_stats': {}})
        tx = _t.SimpleNamespace(update=lambda g, o, p: (jax.tree.map(lambda x: jnp.zeros_like(x), p), o))
        hp = _t.SimpleNamespace(clip_coef=f32(0.2), value_loss_coef=f32(0.5))
        ps = Obj(apply_fn=apply_fn, params=params, batch_stats={})
        ts = Obj(value_normalizer=vn, value_normalizer_state=vn_state, hyper_params=hp, tx=tx, opt_state=None,
                 scaler=None, initial_weight_norms=norms)
        mb = FrozenDict(rnn_start_states=None, dones=None, actions={'act': _arr(acts)}, obs=None,
                        log_probs={'act': _arr(old_lp)}, advantages=_arr(adv), returns=_arr(ret), values=_arr(oldv))
        rec = Rec()
        ps2, ts2, rec = ppo_update(cfg, mb, _arr(mbw), ps, ts, rec)
        o = rec.out
        pre = name + '/'
        out.update({pre + 'logits': logits, pre + 'critic': critic, pre + 'actions': acts, pre + 'old_log_probs': old_lp,
                    pre + 'advantages': adv, pre + 'returns': ret, pre + 'old_values': oldv,
                    pre + 'flags': np.array([int(cfg.algo.clip_value_loss), int(cfg.algo.huber_value_loss),
                                             int(bool(kw.get('vn'))), int(V > 1), int(cfg.compute_advantages)]),
                    pre + 'loss': np.asarray(o['Loss'], f32), pre + 'action_obj': np.asarray(o['Action Obj']),
                    pre + 'value_loss': np.asarray(o['Value Loss']), pre + 'value_errs': np.asarray(o['Value Errors']),
                    pre + 'entropy': np.asarray(o['Entropy'])})
        if vn is not None:
            keys = ('mu', 'inv_sigma', 'sigma', 'mu_biased', 'sigma_sq_biased')
            out[pre + 'vn_before'] = np.array([float(vn_state[k][0]) for k in keys] + [float(vn_state['N'])], np.float64)
            s2 = ts2.value_normalizer_state
            out[pre + 'vn_after'] = np.array([float(s2[k][0]) for k in keys] + [float(s2['N'])], np.float64)
        if name == 'plain':     # re-projection / LayerNorm renorm of the (un-updated) parameters
            net, net2 = params['backbone']['encoder']['net'], ps2.params['backbone']['encoder']['net']
            for k in ('Dense_0', 'Dense_1'):
                out[f'reproj/{k}_in'], out[f'reproj/{k}_out'] = np.asarray(net[k]['kernel']), np.asarray(net2[k]['kernel'])
            for k in ('LayerNorm_0', 'LayerNorm_1'):
                for leaf in ('scale', 'bias'):
                    out[f'reproj/{k}_{leaf}_in'] = np.asarray(net[k]['impl'][leaf])
                    out[f'reproj/{k}_{leaf}_out'] = np.asarray(net2[k]['impl'][leaf])
            out['reproj/norms'] = np.array([3.0, 5.5], f32)
            assert np.array_equal(np.asarray(ps2.params['actor']['impl']['kernel']), np.asarray(params['actor']['impl']['kernel']))
    np.savez_compressed(os.path.join(OUT, 'ppo_loss.npz'), **out)


def select_golden(rng, jax, jnp, lax, ma):
    """Runs the reference's OWN `_ppo` (ml/ppo.py:366-488, AST-extracted, unmodified) under the shim
    for the three selection branches: default, filter_advantages, importance_sample_trajectories.
    `_ppo_update` is replaced by a recorder of (mb_inds-derived minibatch, mb_weights); jax.random
    (third party) by oracle/prng.py's restatement of split / permutation / choice."""
    import types as _t
    import flax
    from flax.core import FrozenDict
    from oracle import prng
    f32 = np.float32
    ns_rd = dict(jax=jax, jnp=jnp, lax=lax, flax=flax, FrozenDict=FrozenDict, Any=object)
    RolloutData = extract_function('rollouts.py', 'RolloutData', ns_rd)

    class Ctx:
        def __init__(self, *a): pass
        def __enter__(self): return self
        def __exit__(self, *a): return False

    rec = []

    def fake_ppo_update(cfg, mb, mb_weights, policy_state, train_state, metrics):
        rec.append((np.asarray(mb['tag']).copy(), np.asarray(mb_weights).copy()))
        return policy_state, train_state, metrics

    def permutation(key, x):
        return _arr(prng.permutation(np.asarray(key), np.asarray(x)))

    def choice(key, n, shape, replace, p):
        assert not replace
        g = prng.gumbel_from_bits(prng.random_bits(np.asarray(key), (n,))) + np.log(np.asarray(p, f32))
        return _arr(np.argsort(-g, kind='stable')[:shape[0]].astype(np.int32))
    random = _t.SimpleNamespace(permutation=permutation, choice=choice)
    ns = dict(jax=jax, jnp=jnp, lax=lax, random=random, profile=Ctx, _ppo_update=fake_ppo_update,
              TrainConfig=object, PolicyState=object, PolicyTrainState=object, RolloutData=object,
              TrainingMetrics=object, Callable=object)
    ppo_fn = extract_function('ppo.py', '_ppo', ns)

    class TS:
        def __init__(self, key, est, est_state):
            self.key, self.max_advantage_est, self.max_advantage_est_state = key, est, est_state
        def gen_update_rnd(self):
            ks = prng.split(self.key, 2)
            return ks[0], TS(ks[1], self.max_advantage_est, self.max_advantage_est_state)
        def update(self, max_advantage_est_state=None):
            return TS(self.key, self.max_advantage_est, max_advantage_est_state)

    out = {}
    J, Tp, M, E = 24, 3, 8, 2
    for name in ('default', 'filter', 'importance'):
        adv = (rng.standard_normal((J, Tp, 1)) * np.exp(rng.standard_normal((J, 1, 1)) * 2)).astype(f32)
        if name == 'filter':
            adv[rng.random((J, Tp, 1)) < 0.6] *= f32(1e-4)            # most elements fall below 1 % of the max
        if name == 'importance':
            adv = (adv * f32(0.05)).astype(f32)                       # keep softmax / the (1/J)/p weights in a sane range
        val = rng.standard_normal((J, Tp, 1)).astype(f32)
        ret = (val + rng.standard_normal((J, Tp, 1))).astype(f32)
        tag = np.arange(J * Tp, dtype=np.int32).reshape(J, Tp, 1)   # identifies (trajectory, step)
        data = FrozenDict(advantages=_arr(adv), values=_arr(val), returns=_arr(ret), dones=_arr(np.zeros((J, Tp, 1), bool)),
                          tag=_arr(tag), rnn_start_states=_arr(np.zeros((J, Tp), f32)))   # flattens to J*T' rows like the rest
        rd = RolloutData(data=data, num_train_seqs_per_policy=J, num_train_policies=1)
        est = ma.EMAEstimate(decay=0.9)
        est_state = est.init_estimates(_arr(np.zeros((1,), f32)))
        for x in (3.0, 5.0):                                          # a non-trivial estimator state
            est_state = est.update_estimates(est_state, _arr(np.array(x, f32)))
        key0 = prng.key(1234 + len(name))
        ts = TS(key0, est, est_state)
        cfg = _t.SimpleNamespace(filter_advantages=name == 'filter', importance_sample_trajectories=name == 'importance',
                                 importance_sample_num_minibatches=2,
                                 algo=_t.SimpleNamespace(minibatch_size=M, num_epochs=E))
        rec.clear()
        _, ts2, _ = ppo_fn(cfg, None, ts, rd, lambda m, *a: m, None)
        pre = name + '/'
        out[pre + 'advantages'], out[pre + 'values'], out[pre + 'returns'] = adv, val, ret
        out[pre + 'key0'] = np.asarray(key0, np.uint32)
        out[pre + 'key1'] = np.asarray(ts2.key, np.uint32)
        out[pre + 'dims'] = np.array([J, Tp, M, E], np.int32)
        out[pre + 'num_minibatches'] = np.int32(len(rec) // E)
        out[pre + 'mb_tags'] = np.stack([r[0].reshape(-1) for r in rec]) if rec else np.zeros((0, M), np.int32)
        out[pre + 'mb_weights'] = np.stack([r[1].reshape(-1) for r in rec]) if rec else np.zeros((0, M), f32)
        out[pre + 'est_before'] = np.array([float(est_state['mu'][0]), float(est_state['mu_biased'][0]), float(est_state['N'])])
        s2 = ts2.max_advantage_est_state
        out[pre + 'est_after'] = np.array([float(s2['mu'][0]), float(s2['mu_biased'][0]), float(s2['N'])])
    np.savez_compressed(os.path.join(OUT, 'ppo_select.npz'), **out)


def hlgauss_golden(rng, jax, jnp):
    """HLGaussDist / HLGaussCritic.create (ml/models.py:177-306), AST-extracted and executed unmodified:
    bin centres and bounds, mean() (symmetric summation) and loss() (Gaussian-histogram cross-entropy)."""
    import types as _t
    import flax
    import scipy.special
    f32 = np.float32
    jax_ns = _t.SimpleNamespace(**{k: getattr(jax, k) for k in dir(jax) if not k.startswith('__')})
    jax_ns.scipy = _t.SimpleNamespace(special=_t.SimpleNamespace(
        erf=lambda x: _arr(scipy.special.erf(np.asarray(x, np.float32)).astype(np.float32))))
    jax_ns.nn = _t.SimpleNamespace(**{k: getattr(jax.nn, k) for k in dir(jax.nn) if not k.startswith('__')})
    jax_ns.nn.initializers = _t.SimpleNamespace(constant=lambda v: None)      # only names a default argument
    nn = _t.SimpleNamespace(Module=object, compact=lambda f: f, Dense=None)
    ns = dict(jax=jax_ns, jnp=jnp, np=np, flax=flax, nn=nn, Callable=object)
    Dist = extract_function('models.py', 'HLGaussDist', ns)
    Critic = extract_function('models.py', 'HLGaussCritic', ns)
    # the staticmethod only builds the bin tables; capture its arguments instead of constructing the Module
    caught = {}
    ns['HLGaussCritic'] = lambda **kw: caught.update(kw)
    out = {}
    for name, (nb, lo, hi, sm) in {'default': (127, -100, 100, 0.75), 'small': (31, -10, 10, 0.5)}.items():
        Critic.create(dtype=jnp.float32, num_bins=nb, min_bound=lo, max_bound=hi, smoothness=sm)
        centers, bounds = np.asarray(caught['centers'], f32), np.asarray(caught['bounds'], f32)
        logits = (rng.standard_normal((40, nb)) * 1.5).astype(f32)
        tgt = (rng.standard_normal((40, 1)) * 0.6 * hi).astype(f32)      # some beyond the outer centres (clipped)
        tgt[:3, 0] = [centers[0], centers[-1], 0.0]
        d = Dist(logits=_arr(logits), smoothness=sm, centers=_arr(centers), bounds=_arr(bounds))
        out[f'{name}_centers'], out[f'{name}_bounds'], out[f'{name}_smoothness'] = centers, bounds, f32(sm)
        out[f'{name}_logits'], out[f'{name}_targets'] = logits, tgt
        out[f'{name}_mean'], out[f'{name}_loss'] = np.asarray(d.mean()), np.asarray(d.loss(_arr(tgt)))
    np.savez_compressed(os.path.join(OUT, 'hlgauss.npz'), **out)


def continuous_golden(rng, jax, jnp, di):
    """ContinuousActionDistributions.action_stats / .best (ml/dists.py:211-284) executed from the reference
    module: two action groups with different std ranges; means / stds in the reference's [rows, groups, dims]."""
    import types as _t
    import scipy.stats
    f32 = np.float32
    if not hasattr(jax, 'scipy'):
        jax.scipy = _t.SimpleNamespace()
    jax.scipy.stats = _t.SimpleNamespace(norm=_t.SimpleNamespace(
        logpdf=lambda x, loc, scale: _arr(scipy.stats.norm.logpdf(np.asarray(x, np.float64), np.asarray(loc, np.float64),
                                                                     np.asarray(scale, np.float64)).astype(np.float32))))
    from madrona_learn.cfg import ContinuousActionsConfig
    cfgs = [ContinuousActionsConfig(stddev_min=0.05, stddev_max=1.5, num_dims=3),
            ContinuousActionsConfig(stddev_min=0.2, stddev_max=0.8, num_dims=3)]
    rows = 64
    means = (rng.standard_normal((rows, 2, 3)) * 1.5).astype(f32)
    stds = (rng.standard_normal((rows, 2, 3)) * 2.0).astype(f32)
    acts = (rng.standard_normal((rows, 2, 3))).astype(f32)
    d = di.ContinuousActionDistributions(cfgs=cfgs, means=_arr(means), stds=_arr(stds))
    lp, ent = d.action_stats(_arr(acts))
    np.savez_compressed(os.path.join(OUT, 'continuous.npz'), means=means, stds=stds, actions=acts,
                        log_probs=np.asarray(lp), entropies=np.asarray(ent), best=np.asarray(d.best()),
                        stddev_min=np.array([c.stddev_min for c in cfgs], f32),
                        stddev_max=np.array([c.stddev_max for c in cfgs], f32))


def multi_policy_golden(jax, jnp):
    """`_compute_sim_to_train_indices` / `_compute_num_train_agents_per_policy` (ml/rollouts.py:1053-1104,
    AST-extracted, unmodified) under the shim: the simulator-order -> training-order index table the reference's
    multi-policy rollouts use.  Pins (a) the self-play layout the multi-policy learner lowers (policy p owns the
    contiguous block p of the simulator batch == `x.reshape(P, -1)`, ml/rollouts.py:579-588) and (b) the general
    table with cross-play / past-play batches (team 0 of every non-self-play match only), for the record."""
    import types as _t
    ns = dict(jax=jax, jnp=jnp)
    nper = extract_function('rollouts.py', '_compute_num_train_agents_per_policy', ns)
    idxf = extract_function('rollouts.py', '_compute_sim_to_train_indices', ns)
    out = {}
    cases = {  # name: (P, num_teams, team_size, self, cross, past)   batch sizes in agents
        'selfplay_p2': (2, 1, 1, 48, 0, 0),
        'selfplay_p3_teams': (3, 2, 2, 72, 0, 0),
        'mixed_p2': (2, 2, 1, 32, 16, 16),
    }
    for name, (P, nt, ts, sp, cp, pp) in cases.items():
        pbt = _t.SimpleNamespace(num_current_policies=P, num_teams=nt, team_size=ts, self_play_batch_size=sp,
                                 cross_play_batch_size=cp, past_play_batch_size=pp)
        cfg = _t.SimpleNamespace(pbt=pbt, sim_batch_size=sp + cp + pp)
        out[name + '_cfg'] = np.array([P, nt, ts, sp, cp, pp], np.int64)
        out[name + '_idx'] = np.asarray(idxf(cfg)).astype(np.int64)
        out[name + '_nper'] = np.int64(nper(cfg))
    np.savez_compressed(os.path.join(OUT, 'multi_policy.npz'), **out)


def main():
    install()
    import jax
    import jax.numpy as jnp
    from jax import lax
    ac = load_reference('algo_common')
    ma = load_reference('moving_avg')
    me = load_reference('metrics')
    di = load_reference('dists')
    rng = np.random.default_rng(20261018)
    f32 = np.float32

    # ---- GAE / returns / zscore: ml/algo_common.py:45-140 ---------------------------------
    cases = {}
    for name, (C, Tp, P, B, pd, g, lam) in {
            'small': (1, 32, 1, 32, 0.05, 0.99, 0.95), 'chunks': (4, 8, 1, 48, 0.1, 0.998, 0.9),
            'ragged': (3, 5, 1, 7, 0.3, 0.9, 1.0), 'nodone': (1, 16, 1, 33, 0.0, 0.99, 0.95),
            'alldone': (1, 4, 1, 9, 1.0, 0.99, 0.95), 'one': (1, 1, 1, 1, 0.0, 0.5, 0.5)}.items():
        cfg = types.SimpleNamespace(gamma=g, gae_lambda=lam)
        r = rng.standard_normal((C, Tp, P, B, 1)).astype(f32)
        v = rng.standard_normal((C, Tp, P, B, 1)).astype(f32)
        d = rng.random((C, Tp, P, B, 1)) < pd
        b = rng.standard_normal((P, B, 1)).astype(f32)
        adv = np.asarray(ac.compute_advantages(cfg, _arr(r), _arr(v), _arr(d), _arr(b)))
        ret = np.asarray(ac.compute_returns(cfg, _arr(r), _arr(d), _arr(b)))
        z = np.asarray(ac.zscore_data(_arr(adv)))
        for k, x in dict(r=r, v=v, d=d, b=b, adv=adv, ret=ret, z=z,
                         cfg=np.array([g, lam], np.float64)).items():
            cases[f'{name}/{k}'] = x
    np.savez_compressed(os.path.join(OUT, 'algo_common.npz'), **cases)

    # ---- EMANormalizer: ml/moving_avg.py:48-198 (tests/test_ema.py recipe) -----------------
    norm = ma.EMANormalizer(decay=0.999, norm_dtype=jnp.float32, inv_dtype=jnp.float32)
    iters, batch, dims, sub = 40, 256, 2, 8
    means = rng.random((iters, dims)) * 100 - 5
    stds = rng.random((iters, dims)) * 2000 + 2
    means[-1], stds[-1] = -20, 0.01
    vals = (rng.standard_normal((iters, batch, dims)) * stds[:, None] + means[:, None]).astype(f32)
    est = norm.init_estimates(_arr(vals[0]))
    hist = []
    for i in range(iters):
        stats = norm.init_input_stats(est)
        for j in range(sub):
            stats = norm.update_input_stats(stats, j, _arr(vals[i].reshape(sub, batch // sub, dims)[j]))
        est = norm.update_estimates(est, stats)
        hist.append(np.concatenate([np.asarray(est[k]).reshape(-1) for k in
                                    ('mu', 'inv_sigma', 'sigma', 'mu_biased', 'sigma_sq_biased')] +
                                   [np.asarray(stats[0]), np.asarray(stats[1])]))
    est2, normed = norm.normalize_and_update_estimates(est, _arr(vals[3]))
    np.savez_compressed(os.path.join(OUT, 'ema.npz'), vals=vals, hist=np.stack(hist).astype(f32),
                        N=np.int32(est['N']), normalized=np.asarray(normed),
                        inverted=np.asarray(norm.invert(est, _arr(vals[5]))),
                        est2_mu=np.asarray(est2['mu']), est2_inv_sigma=np.asarray(est2['inv_sigma']))
    # scalar (value-normaliser) case, decay 0.99999, normalize_and_update per "minibatch"
    vn = ma.EMANormalizer(decay=0.99999, norm_dtype=jnp.float32, inv_dtype=jnp.float32)
    e = vn.init_estimates(_arr(np.zeros((1, 1), f32)))
    rets = (rng.standard_normal((12, 64, 1)) * 3 + 1.5).astype(f32)
    vh = []
    for i in range(12):
        e, nr = vn.normalize_and_update_estimates(e, _arr(rets[i]))
        vh.append([float(e['mu'][0]), float(e['inv_sigma'][0]), float(e['sigma'][0]),
                   float(e['mu_biased'][0]), float(e['sigma_sq_biased'][0])])
    np.savez_compressed(os.path.join(OUT, 'ema_value_norm.npz'), rets=rets, hist=np.array(vh, f32),
                        last_normalized=np.asarray(nr))

    # ---- Metric: ml/metrics.py:31-98 -------------------------------------------------------
    x1 = (rng.standard_normal((7, 33)) * 4 - 2).astype(f32)
    x2 = (rng.standard_normal((5, 11)) + 3).astype(f32)
    m1 = me.Metric.init_from_data(True, _arr(x1))
    m2 = me.Metric.init_from_data(True, _arr(x2))
    mm = m1.merge(m2)
    pack = lambda m: np.array([m.mean, m.m2, m.min, m.max, m.count], np.float64)
    np.savez_compressed(os.path.join(OUT, 'metric.npz'), x1=x1, x2=x2, m1=pack(m1), m2=pack(m2),
                        merged=pack(mm))

    # ---- DiscreteActionDistributions.action_stats / best: ml/dists.py:46-77 ---------------
    buckets = [4, 8, 5, 5, 2, 2]
    logits = (rng.standard_normal((50, sum(buckets))) * 2).astype(f32)
    acts = np.stack([rng.integers(0, b, 50) for b in buckets], -1).astype(np.int32)
    dist = di.DiscreteActionDistributions(actions_num_buckets=buckets, all_logits=_arr(logits))
    lp, ent = dist.action_stats(_arr(acts))
    np.savez_compressed(os.path.join(OUT, 'dists.npz'), logits=logits, actions=acts,
                        log_probs=np.asarray(lp), entropies=np.asarray(ent),
                        best=np.asarray(dist.best()), buckets=np.array(buckets))
    # two-hot critic (default critic of the reference; "next" row) ml/dists.py:119-208
    tl = (rng.standard_normal((20, 63))).astype(f32)
    tgt = (rng.standard_normal((20, 1)) * 50).astype(f32)
    th = di.SymExpTwoHotDistribution.create(_arr(tl))
    np.savez_compressed(os.path.join(OUT, 'twohot.npz'), logits=tl, targets=tgt,
                        mean=np.asarray(th.mean()), loss=np.asarray(th.two_hot_cross_entropy_loss(_arr(tgt))))

    # ---- buffer relayout + minibatch: ml/rollouts.py:311-334, 788-804 (AST-extracted) ------
    import flax
    from flax.core import FrozenDict
    ns = dict(jax=jax, jnp=jnp, lax=lax, flax=flax, FrozenDict=FrozenDict, Any=object)
    RolloutData = extract_function('rollouts.py', 'RolloutData', ns)
    C, Tp, P, B = 3, 4, 1, 5
    store = {'obs': rng.integers(0, 1000, (C, Tp, P, B, 6)).astype(np.int32),
             'rewards': rng.standard_normal((C, Tp, P, B, 1)).astype(f32)}
    rnn = rng.standard_normal((C, P, B, 3)).astype(f32)

    def reorder_seq_data(x):                       # verbatim semantics of :791-793
        t = x.transpose(2, 0, 3, 1, *range(4, len(x.shape)))
        return t.reshape(t.shape[0], -1, *t.shape[3:])
    data = FrozenDict({k: _arr(reorder_seq_data(v)[0]) for k, v in store.items()})
    data = data.copy({'rnn_start_states': _arr(rnn.transpose(1, 0, 2, 3).reshape(P, C * B, 3)[0])})
    rd = RolloutData(data=data, num_train_seqs_per_policy=C * B, num_train_policies=P)
    idx = rng.permutation(C * B)[:6].astype(np.int32)
    mb = rd.minibatch(_arr(idx))
    np.savez_compressed(os.path.join(OUT, 'minibatch.npz'), idx=idx, rnn=rnn,
                        **{f'store_{k}': v for k, v in store.items()},
                        **{f'mb_{k}': np.asarray(v) for k, v in mb.items()})

    # ---- _compute_reorder_chunks KAT inputs: tests/test_rollouts.py:36-81 -------------------
    crc = extract_function('rollouts.py', '_compute_reorder_chunks', dict(jax=jax, jnp=jnp, lax=lax))
    vecs = [[1, 1, 0, 0, 2, 2, 5, 3, 2, 1, 0, 3, 3], [1, 1, 0, 0, 2, 2, 4, 5, 2, 1, 0, 3],
            [1, 1, 0, 0, 2, 2, 4, 3, 2, 1, 0, 3],
            list(rng.permutation([0, 0, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5]))]
    out = {}
    for i, vec in enumerate(vecs):
        a = np.array(vec, np.int32)
        Pn, Cn = 6, 4
        Bn = a.size // Cn + Pn - 1
        tp, ts = crc(_arr(a), Pn, Cn, Bn)
        out[f'v{i}_in'], out[f'v{i}_to_policy'], out[f'v{i}_to_sim'] = a, np.asarray(tp), np.asarray(ts)
    np.savez_compressed(os.path.join(OUT, 'reorder_chunks.npz'), **out)
    # ---- the composite PPO loss + re-projection: ml/ppo.py:109-362 (own generator so the
    # fixtures above keep their random stream and stay bit-identical) ------------------------
    ppo_update_golden(np.random.default_rng(20261019), jax, jnp, lax, ac, ma, di)
    select_golden(np.random.default_rng(20261020), jax, jnp, lax, ma)
    hlgauss_golden(np.random.default_rng(20261021), jax, jnp)
    continuous_golden(np.random.default_rng(20261022), jax, jnp, di)
    multi_policy_golden(jax, jnp)
    print('golden fixtures written to', OUT)


if __name__ == '__main__':
    main()
