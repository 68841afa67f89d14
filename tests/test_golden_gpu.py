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
