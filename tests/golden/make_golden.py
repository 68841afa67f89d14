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
    print('golden fixtures written to', OUT)


if __name__ == '__main__':
    main()
