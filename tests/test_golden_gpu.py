"""GPU: CUDA kernels (through the C-ABI) vs the committed golden vectors that were produced by
the reference's own source under oracle/jax_shim (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
_dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


@pytest.mark.parametrize('name', ['small', 'chunks', 'ragged', 'nodone', 'alldone', 'one'])
def test_gae_returns_zscore_vs_reference_golden(mlb, name):
    K = mlb.kernels
    z = np.load(os.path.join(G, 'algo_common.npz'))
    g = lambda k: z[f'{name}/{k}']
    gamma, lam = g('cfg')
    T = g('r').shape[0] * g('r').shape[1]
    r, v, d, b = (g('r').reshape(T, -1), g('v').reshape(T, -1), g('d').reshape(T, -1), g('b').reshape(-1))
    adv, ret = K.gae(_dev(r), _dev(v), _dev(d), _dev(b), float(gamma), float(lam))
    np.testing.assert_array_equal(adv.cpu().numpy(), g('adv').reshape(T, -1))          # bit-exact
    ret2 = K.discounted_returns(_dev(r), _dev(d), _dev(b), float(gamma))
    np.testing.assert_array_equal(ret2.cpu().numpy(), g('ret').reshape(T, -1))
    zs = K.zscore(_dev(g('adv').reshape(-1)))
    np.testing.assert_allclose(zs.cpu().numpy(), g('z').reshape(-1), rtol=1e-5, atol=1e-6)


def test_ema_update_vs_reference_golden(mlb):
    K = mlb.kernels
    z = np.load(os.path.join(G, 'ema.npz'))
    hist = z['hist']
    dims = 2
    st = K.ema_state_init(dims, DEV)
    for i in range(hist.shape[0]):
        K.ema_update(st, dims, _dev(hist[i, 10:12]), _dev(hist[i, 12:14]), 0.999)
        np.testing.assert_allclose(st[:10].cpu().numpy(), hist[i, :10], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(K.ema_invert(st, dims, _dev(z['vals'][5])).cpu().numpy(), z['inverted'], rtol=1e-5)


def test_value_norm_scan_vs_reference_golden(mlb):
    K = mlb.kernels
    z = np.load(os.path.join(G, 'ema_value_norm.npz'))
    rets = z['rets']                                   # [12, 64, 1]: 12 "minibatches"
    Kmb, M = rets.shape[0], rets.shape[1]
    x = _dev(rets.reshape(Kmb * M)[None, :].copy())    # T=1, N = Kmb*M, one trajectory per element
    tm = K.traj_moments(x, 1)
    perm = _dev(np.arange(Kmb * M, dtype=np.int32)[None, :])
    mbm = K.mb_moments(tm, perm, M, 1, 0.0)
    st = K.ema_state_init(1, DEV)
    out = K.ema_scan(st, mbm, 0.99999).cpu().numpy()
    for i in range(Kmb):
        np.testing.assert_allclose(out[i, 2:4], z['hist'][i, 0:2], rtol=3e-5, atol=1e-7)   # mu_new, inv_sigma_new
        if i:
            np.testing.assert_allclose(out[i, 0:2], z['hist'][i - 1, [0, 2]], rtol=3e-5, atol=1e-7)
    np.testing.assert_allclose(st[:5].cpu().numpy(), z['hist'][-1], rtol=3e-5, atol=1e-7)
    nr = K.ema_normalize(st, 1, _dev(rets[-1]))
    np.testing.assert_allclose(nr.cpu().numpy(), z['last_normalized'], rtol=1e-4, atol=1e-5)


def test_metric_vs_reference_golden(mlb):
    K = mlb.kernels
    z = np.load(os.path.join(G, 'metric.npz'))
    for x, ref in ((z['x1'], z['m1']), (z['x2'], z['m2'])):
        g = K.metrics_to_host(K.metric(_dev(x)), 1)[0]
        np.testing.assert_allclose([g['mean'], g['m2'], g['min'], g['max'], g['count']], ref, rtol=1e-5)


def test_minibatch_gather_vs_reference_golden(mlb):
    K = mlb.kernels
    z = np.load(os.path.join(G, 'minibatch.npz'))
    idx = _dev(z['idx'])
    for k in ('obs', 'rewards'):
        store = z[f'store_{k}']
        C, Tp, P, B = store.shape[:4]
        out = K.mb_gather(_dev(store[:, :, 0]), idx, C, Tp, B)
        np.testing.assert_array_equal(out.cpu().numpy(), z[f'mb_{k}'])                # bit-exact
    rnn = z['rnn']
    out = K.mb_gather_rnn(_dev(rnn[:, 0]), idx, rnn.shape[0], rnn.shape[2])
    np.testing.assert_array_equal(out.cpu().numpy(), z['mb_rnn_start_states'])


# ------------------------------------------------------------------------------------------
# composite PPO loss (ml/ppo.py:129-262), re-projection / LayerNorm renorm (:300-338),
# DiscreteActionDistributions (ml/dists.py:46-77) and the two-hot critic (:143-208): the
# KERNELS against fixtures produced by the reference's own source
# ------------------------------------------------------------------------------------------
BUCKETS = [4, 8, 5, 5, 2, 2]


def _loss_kernel(mlb, head, ld, actions, old_lp, scores, returns, old_values, adv_mr, vn_params, rows, M,
                 clip_coef, vcoef, ent_coef, flags, V):
    """One mlb_ppo_loss_f32 launch on explicit head outputs -> (PPOStats, d_head)."""
    import ctypes
    from madrona_learn_b200 import _lib
    from madrona_learn_b200._lib import c_float, c_int, c_ll, c_size_t, call, ptr
    from oracle import dists as odists
    A = len(BUCKETS)
    dhead = torch.zeros(rows, ld, device=DEV)
    dbias = torch.zeros(ld, device=DEV)
    stats = torch.zeros(ctypes.sizeof(_lib.PPOStats), dtype=torch.uint8, device=DEV)
    ws = torch.zeros(_lib.lib().mlb_ppo_loss_workspace(rows) + 16, dtype=torch.uint8, device=DEV)
    buckets_c = (ctypes.c_int32 * A)(*BUCKETS)
    obj_scale = (ctypes.c_float * A)(*[1.0 / (rows * A)] * A)
    ent_scale = (ctypes.c_float * A)(*[ent_coef / (rows * A)] * A)
    bins_c = (ctypes.c_float * V)(*odists.bins(V).tolist()) if V > 1 else None
    call('mlb_ppo_loss_f32', ptr(head), c_int(ld), ptr(actions), ptr(old_lp), ptr(scores), ptr(returns),
         ptr(old_values), ptr(None), ptr(adv_mr), ptr(vn_params), buckets_c, obj_scale, ent_scale, c_int(A),
         c_ll(rows), c_ll(M), c_float(clip_coef), c_float(vcoef), c_int(flags), ptr(dhead), ptr(dbias), ptr(stats),
         ptr(ws), c_size_t(ws.numel()), bins_c, c_int(V if V > 1 else 0))
    torch.cuda.synchronize()
    return _lib.PPOStats.from_buffer_copy(stats.cpu().numpy().tobytes()), dhead


@pytest.mark.parametrize('name', ['plain', 'clipv_huber', 'valuenorm', 'valuenorm_clipv', 'twohot', 'returns_only'])
def test_ppo_loss_kernel_vs_reference_golden(mlb, name):
    K = mlb.kernels
    z = np.load(os.path.join(G, 'ppo_loss.npz'))
    g = lambda k: z[f'{name}/{k}']
    clipv, huber, vn, twohot, use_adv = [int(x) for x in g('flags')]
    Tp, M, A = g('actions').shape
    rows = Tp * M
    V = g('critic').shape[-1]
    sumA = sum(BUCKETS)
    ld = (sumA + V + 3) // 4 * 4
    head = np.zeros((rows, ld), np.float32)
    head[:, :sumA] = g('logits').reshape(rows, sumA)
    head[:, sumA:sumA + V] = g('critic').reshape(rows, V)
    scores = g('advantages') if use_adv else g('returns')
    adv_mr = K.moments(_dev(scores.reshape(-1)), 1e-5)                      # the z-score statistics kernel
    vn_params = None
    if vn:                                                                  # the value-normaliser recurrence kernel
        st = torch.zeros(6, device=DEV)
        st[:5] = _dev(g('vn_before')[:5].astype(np.float32))
        st[5:].view(torch.int32).fill_(int(g('vn_before')[5]))
        mom = K.moments(_dev(g('returns').reshape(-1)), 0.0).view(1, 4)
        vn_params = K.ema_scan(st, mom, 0.99999)
        np.testing.assert_allclose(st[:5].cpu().numpy(), g('vn_after')[:5], rtol=3e-5, atol=1e-7)
        assert int(st[5:].view(torch.int32).item()) == int(g('vn_after')[5])
    flags = (1 if clipv else 0) | (2 if huber else 0)
    st_out, _ = _loss_kernel(mlb, _dev(head), ld, _dev(g('actions')), _dev(g('old_log_probs')), _dev(scores),
                             _dev(g('returns')), _dev(g('old_values')), adv_mr, vn_params, rows, M, 0.2, 0.5, 0.02,
                             flags, V)
    np.testing.assert_allclose(st_out.loss, float(g('loss')), rtol=1e-4, atol=1e-6)
    refs = [np.array([g('loss')]), g('action_obj'), g('value_loss'), g('value_errs'), g('entropy')]
    for i, x in enumerate(refs):
        x = np.asarray(x, np.float32)
        m = st_out.metrics[i]
        assert m.count == x.size
        np.testing.assert_allclose([m.mean, m.min, m.max], [x.mean(dtype=np.float64), x.min(), x.max()],
                                   rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(m.m2, np.sum((x.astype(np.float64) - x.mean(dtype=np.float64)) ** 2),
                                   rtol=2e-3, atol=1e-5)


def test_reprojection_kernels_vs_reference_golden(mlb):
    """mlb_optimizer_step_fused and mlb_renorm_segments on a zero gradient == the reference's
    normalize_params + normalize_layernorms (ml/ppo.py:300-338)."""
    from madrona_learn_b200.engine import PolicyProgram
    m = mlb
    z = np.load(os.path.join(G, 'ppo_loss.npz'))
    D, H = z['reproj/Dense_0_in'].shape
    for fused in (True, False):
        ac = m.ActorCritic(
            backbone=m.BackboneShared(prefix=None, encoder=m.BackboneEncoder(net=m.models.MLP(H, 2))),
            actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig([3, 2])),
            critic=m.models.DenseLayerCritic())
        prog = PolicyProgram(ac, D, {'act': m.DiscreteActionsConfig([3, 2])}, DEV, torch.float32)
        prog.init_params(0)
        with torch.no_grad():
            for i in range(2):
                k, s, b = prog.layer_views(prog.params, i)
                k.copy_(_dev(z[f'reproj/Dense_{i}_in']))
                s.copy_(_dev(z[f'reproj/LayerNorm_{i}_scale_in']))
                b.copy_(_dev(z[f'reproj/LayerNorm_{i}_bias_in']))
        prog.initial_weight_norms = {f'Dense_{i}': float(z['reproj/norms'][i]) for i in range(2)}
        prog.rebuild_segments()
        prog._fused_opt = prog._fused_opt and fused
        head0 = prog.head_views(prog.params)[0].clone()
        prog.grads.zero_()
        prog.optimizer_step(0.0, 0.5)                      # lr = 0: only the re-projection acts
        torch.cuda.synchronize()
        for i in range(2):
            k, s, b = prog.layer_views(prog.params, i)
            np.testing.assert_allclose(k.cpu().numpy(), z[f'reproj/Dense_{i}_out'], rtol=3e-6)
            np.testing.assert_allclose(s.cpu().numpy(), z[f'reproj/LayerNorm_{i}_scale_out'], rtol=3e-6)
            np.testing.assert_allclose(b.cpu().numpy(), z[f'reproj/LayerNorm_{i}_bias_out'], rtol=3e-6)
        assert torch.equal(head0, prog.head_views(prog.params)[0])      # actor / critic are not re-projected


def test_discrete_dists_kernels_vs_reference_golden(mlb):
    """action_stats / best (ml/dists.py:46-77): the sampling kernel's greedy mode and the loss
    kernel's log-prob / entropy arithmetic against the reference-generated dists.npz."""
    import ctypes
    from madrona_learn_b200._lib import c_int, c_ll, call, ptr
    z = np.load(os.path.join(G, 'dists.npz'))
    logits, acts = z['logits'], z['actions']
    rows, sumA = logits.shape
    A = len(BUCKETS)
    ld = 28
    head = np.zeros((rows, ld), np.float32)
    head[:, :sumA] = logits
    hd = _dev(head)
    best = torch.zeros(rows, A, dtype=torch.int32, device=DEV)
    vals = torch.zeros(rows, device=DEV)
    buckets_c = (ctypes.c_int32 * A)(*BUCKETS)
    call('mlb_sample_discrete_f32', ptr(hd), c_int(ld), ptr(None), buckets_c, c_int(A), c_ll(rows), c_int(0),
         c_int(1), ptr(best), ptr(None), ptr(vals), None, c_int(1))
    np.testing.assert_array_equal(best.cpu().numpy(), z['best'])
    # log-probs: with old_log_probs = 0, scores = 1 (no z-score) and an infinite clip range the per-
    # element action objective IS exp(new_log_prob); entropies are recorded as they are
    one = torch.ones(rows, 1, device=DEV)
    st, _ = _loss_kernel(mlb, hd, ld, _dev(acts), torch.zeros(rows, A, device=DEV), one, one, None, None, None,
                         rows, rows, 1e30, 0.5, 0.02, 0, 1)
    for m, x in ((st.metrics[1], np.exp(z['log_probs'].astype(np.float64))), (st.metrics[4], z['entropies'])):
        np.testing.assert_allclose([m.mean, m.min, m.max], [x.mean(), x.min(), x.max()], rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(m.m2, np.sum((x - x.mean()) ** 2), rtol=1e-3)


def test_twohot_kernels_vs_reference_golden(mlb):
    """SymExpTwoHotDistribution.mean / two_hot_cross_entropy_loss (ml/dists.py:143-208)."""
    import ctypes
    from madrona_learn_b200._lib import c_int, c_ll, call, ptr
    from oracle import dists as odists
    z = np.load(os.path.join(G, 'twohot.npz'))
    tl, tgt = z['logits'], z['targets']
    rows, V = tl.shape
    sumA, A = sum(BUCKETS), len(BUCKETS)
    ld = (sumA + V + 3) // 4 * 4
    head = np.zeros((rows, ld), np.float32)
    head[:, sumA:sumA + V] = tl
    hd = _dev(head)
    acts = torch.zeros(rows, A, dtype=torch.int32, device=DEV)
    vals = torch.zeros(rows, device=DEV)
    buckets_c = (ctypes.c_int32 * A)(*BUCKETS)
    bins_c = (ctypes.c_float * V)(*odists.bins(V).tolist())
    call('mlb_sample_discrete_f32', ptr(hd), c_int(ld), ptr(None), buckets_c, c_int(A), c_ll(rows), c_int(0),
         c_int(1), ptr(acts), ptr(None), ptr(vals), bins_c, c_int(V))
    np.testing.assert_allclose(vals.cpu().numpy(), z['mean'].reshape(-1), rtol=2e-4, atol=2e-4)
    one = torch.ones(rows, 1, device=DEV)
    st, _ = _loss_kernel(mlb, hd, ld, acts, torch.zeros(rows, A, device=DEV), one, _dev(tgt), None, None, None,
                         rows, rows, 0.2, 0.5, 0.02, 0, V)
    x = z['loss'].astype(np.float64)
    m = st.metrics[2]
    np.testing.assert_allclose([m.mean, m.min, m.max], [x.mean(), x.min(), x.max()], rtol=2e-5)
