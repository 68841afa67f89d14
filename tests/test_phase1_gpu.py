"""GPU parity: GAE / returns / z-score / metrics / EMA / PRNG / gather vs the CPU oracle.

All calls go through the C-ABI (madrona_learn_b200.kernels -> libmlb200.so).
Tolerances (stated): GAE/returns bit-exact vs the op-by-op float32 oracle;
z-score/EMA rel 1e-5; metrics rel 1e-5 (m2 1e-4); indexing / PRNG / permutation exact.
"""
import numpy as np
import pytest
import torch

from oracle import algo_common as oac
from oracle import layouts, metrics as omet, moving_avg as oma, prng

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _mk(T, N, seed=0, p_done=0.02):
    rng = np.random.default_rng(seed)
    r = rng.standard_normal((T, N)).astype(np.float32)
    v = rng.standard_normal((T, N)).astype(np.float32)
    d = (rng.random((T, N)) < p_done)
    b = rng.standard_normal(N).astype(np.float32)
    return r, v, d, b


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


# sizes exercise VEC=1 (small / odd N), VEC=2, VEC=4 and T % U remainders
GAE_SHAPES = [(1, 1), (3, 5), (32, 32), (33, 1001), (7, 8192), (16, 151552 + 2), (5, 606208),
              (13, 75776 * 4 + 4)]


@pytest.mark.parametrize('T,N', GAE_SHAPES)
def test_gae_bit_exact(mlb, T, N):
    K = mlb.kernels
    r, v, d, b = _mk(T, N, seed=T * 7 + N)
    adv, ret = K.gae(_dev(r), _dev(v), _dev(d), _dev(b), 0.99, 0.95)
    ref = oac.compute_advantages(0.99, 0.95, r, v, d, b)
    np.testing.assert_array_equal(adv.cpu().numpy(), ref)
    np.testing.assert_array_equal(ret.cpu().numpy(), (ref + v).astype(np.float32))


@pytest.mark.parametrize('T,N', GAE_SHAPES)
def test_returns_bit_exact(mlb, T, N):
    K = mlb.kernels
    r, v, d, b = _mk(T, N, seed=T + N)
    ret = K.discounted_returns(_dev(r), _dev(d), _dev(b), 0.99)
    np.testing.assert_array_equal(ret.cpu().numpy(), oac.compute_returns(0.99, r, d, b))


def test_gae_all_done_and_none_done(mlb):
    K = mlb.kernels
    r, v, d, b = _mk(9, 640)
    for dd in (np.zeros_like(d), np.ones_like(d)):
        adv, _ = K.gae(_dev(r), _dev(v), _dev(dd), _dev(b), 0.9, 0.8)
        np.testing.assert_array_equal(adv.cpu().numpy(),
                                      oac.compute_advantages(0.9, 0.8, r, v, dd, b))


def test_gae_value_denorm_and_metrics(mlb):
    K = mlb.kernels
    T, N = 32, 4096
    r, v, d, b = _mk(T, N, seed=3)
    mu, sigma = np.float32(0.7), np.float32(2.5)
    vn = _dev(np.array([mu, 1 / sigma, sigma, 0, 0, 0], np.float32))
    mbuf = torch.zeros(80, dtype=torch.uint8, device=DEV)
    adv, ret = K.gae(_dev(r), _dev(v), _dev(d), _dev(b), 0.99, 0.95, vn_mu_sigma=vn, metrics=mbuf)
    v_un = (v * sigma + mu).astype(np.float32)
    b_un = (b * sigma + mu).astype(np.float32)
    ref = oac.compute_advantages(0.99, 0.95, r, v_un, d, b_un)
    np.testing.assert_array_equal(adv.cpu().numpy(), ref)
    rets = (ref + v_un).astype(np.float32)
    np.testing.assert_array_equal(ret.cpu().numpy(), rets)
    got = K.metrics_to_host(mbuf, 4)
    for g, x in zip(got, (r, v_un, rets, ref)):
        e = omet.metric_from_data(x)
        assert g['count'] == e['count']
        np.testing.assert_allclose(g['mean'], e['mean'], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(g['m2'], e['m2'], rtol=1e-4)
        assert g['min'] == e['min'] and g['max'] == e['max']


def test_gae_linearity_full_size(mlb):
    """Size-independent property at a sweep size the oracle is too slow for:
    GAE is linear in (rewards, values, bootstrap) for fixed dones."""
    K = mlb.kernels
    T, N = 64, 1 << 20
    g = torch.Generator(device=DEV).manual_seed(1)
    mk = lambda *s: torch.randn(*s, device=DEV, generator=g)
    r1, v1, b1, r2, v2, b2 = mk(T, N), mk(T, N), mk(N), mk(T, N), mk(T, N), mk(N)
    d = (torch.rand(T, N, device=DEV, generator=g) < 0.02)
    a1, _ = K.gae(r1, v1, d, b1, 0.99, 0.95)
    a2, _ = K.gae(r2, v2, d, b2, 0.99, 0.95)
    a12, ret12 = K.gae(r1 + r2, v1 + v2, d, b1 + b2, 0.99, 0.95)
    torch.testing.assert_close(a12, a1 + a2, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ret12, a12 + (v1 + v2), rtol=0, atol=0)
    # one column against the oracle
    cols = slice(12345, 12349)
    ref = oac.compute_advantages(0.99, 0.95, r1[:, cols].cpu().numpy(), v1[:, cols].cpu().numpy(),
                                 d[:, cols].cpu().numpy(), b1[cols].cpu().numpy())
    np.testing.assert_array_equal(a1[:, cols].cpu().numpy(), ref)


@pytest.mark.parametrize('n', [1, 5, 1000, 65536, 1 << 20, (1 << 20) + 3])
def test_zscore(mlb, n):
    K = mlb.kernels
    x = (np.random.default_rng(n).standard_normal(n) * 3 + 1.5).astype(np.float32)
    out = K.zscore(_dev(x))
    np.testing.assert_allclose(out.cpu().numpy(), oac.zscore_data(x), rtol=1e-5, atol=1e-6)
    m4 = K.moments(_dev(x)).cpu().numpy()
    mean, rstd = oac.zscore_stats(x)
    np.testing.assert_allclose(m4[:2], [mean, rstd], rtol=1e-5, atol=1e-7)


def test_zscore_constant_input_uses_var_floor(mlb):
    K = mlb.kernels
    x = np.full(4096, 2.0, np.float32)
    out = K.zscore(_dev(x)).cpu().numpy()
    np.testing.assert_allclose(out, oac.zscore_data(x), atol=1e-6)


def test_metric(mlb):
    K = mlb.kernels
    x = (np.random.default_rng(5).standard_normal((37, 1001)) * 4 - 2).astype(np.float32)
    g = K.metrics_to_host(K.metric(_dev(x)), 1)[0]
    e = omet.metric_from_data(x)
    assert g['count'] == e['count'] and g['min'] == e['min'] and g['max'] == e['max']
    np.testing.assert_allclose([g['mean'], g['m2']], [e['mean'], e['m2']], rtol=1e-5)


@pytest.mark.parametrize('C,Tp,N,M,E', [(1, 32, 512, 128, 4), (4, 8, 96, 64, 2), (2, 5, 10, 4, 3)])
def test_minibatch_moments(mlb, C, Tp, N, M, E):
    K = mlb.kernels
    rng = np.random.default_rng(C + N)
    T = C * Tp
    x = (rng.standard_normal((T, N)) * 2 + 0.3).astype(np.float32)
    J = C * N
    perm = np.stack([rng.permutation(J) for _ in range(E)]).astype(np.int32)
    tm = K.traj_moments(_dev(x), C)
    out = K.mb_moments(tm, _dev(perm), M, Tp).cpu().numpy()
    store = x.reshape(C, Tp, 1, N, 1)
    data = layouts.reorder_seq_data(store)[0]          # [J, T', 1]
    k = 0
    for e in range(E):
        for i in range(J // M):
            mb = data[perm[e, i * M:(i + 1) * M]]
            mean, rstd = oac.zscore_stats(mb)
            np.testing.assert_allclose(out[k, :2], [mean, rstd], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(out[k, 2], mb.astype(np.float64).var(), rtol=1e-5)
            assert out[k, 3] == M * Tp
            k += 1


def test_ema_matches_oracle_and_naive_f64(mlb):
    """tests/test_ema.py recipe (decay .999, 100 iters, last batch mean -20 / sigma .01)."""
    K = mlb.kernels
    decay, iters, batch, dims = 0.999, 100, 1024, 2
    rng = np.random.default_rng(5)
    means = rng.random((iters, dims)) * 100 - 5
    stds = rng.random((iters, dims)) * 2000 + 2
    means[-1] = -20
    stds[-1] = 0.01
    vals = (rng.standard_normal((iters, batch, dims)) * stds[:, None] + means[:, None]).astype(np.float32)
    norm = oma.EMANormalizer(decay)
    est = norm.init_estimates(dims)
    st = K.ema_state_init(dims, DEV)
    nx = np.zeros(dims)
    nxx = np.zeros(dims)
    for i in range(iters):
        stats = norm.update_input_stats(norm.init_input_stats(est), 0, vals[i])
        est = norm.update_estimates(est, stats)
        K.ema_update(st, dims, _dev(stats[0]), _dev(stats[1]), decay)
        nx = decay * nx + (1 - decay) * vals[i].astype(np.float64).mean(0)
        nxx = decay * nxx + (1 - decay) * np.square(vals[i].astype(np.float64)).mean(0)
    h = st.cpu().numpy()
    got = dict(mu=h[0:2], inv_sigma=h[2:4], sigma=h[4:6], mu_biased=h[6:8], sigma_sq_biased=h[8:10])
    for k in got:
        np.testing.assert_allclose(got[k], est[k], rtol=2e-5, err_msg=k)
    assert st[10:].view(torch.int32).item() == iters
    bc = -np.expm1(iters * np.log(decay))
    np.testing.assert_allclose(got['mu'], nx / bc, rtol=1e-3)
    np.testing.assert_allclose(got['sigma'], np.sqrt(nxx / bc - (nx / bc) ** 2), rtol=1e-3)
    x = vals[3]
    np.testing.assert_allclose(K.ema_normalize(st, dims, _dev(x)).cpu().numpy(),
                               norm.normalize(est, x), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(K.ema_invert(st, dims, _dev(x)).cpu().numpy(),
                               norm.invert(est, x), rtol=1e-5)


def test_ema_scan_equals_sequential_updates(mlb):
    K = mlb.kernels
    rng = np.random.default_rng(11)
    Kmb = 16
    mbm = np.zeros((Kmb, 4), np.float32)
    mbm[:, 0] = rng.standard_normal(Kmb) * 3
    mbm[:, 2] = rng.random(Kmb) * 5 + 0.1
    norm = oma.EMANormalizer(0.99999)
    est = norm.init_estimates(1)
    st = K.ema_state_init(1, DEV)
    out = K.ema_scan(st, _dev(mbm), 0.99999).cpu().numpy()
    for k in range(Kmb):
        old = est
        est = norm.update_estimates(est, (mbm[k, 0:1], mbm[k, 2:3]))
        np.testing.assert_allclose(out[k], [old['mu'][0], old['sigma'][0], est['mu'][0],
                                            est['inv_sigma'][0]], rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize('part', [False, True])
def test_threefry_split_bits(mlb, part):
    K = mlb.kernels
    key = prng.key(42)
    kd = torch.from_numpy(key.view(np.int32)).to(DEV)
    for num in (1, 2, 3, 7):
        got = K.threefry_split(kd, num, part).cpu().numpy().view(np.uint32)
        np.testing.assert_array_equal(got, prng.split(key, num, part))
    for n in (1, 2, 5, 1000, 8192):
        got = K.threefry_bits(kd, n, part).cpu().numpy().view(np.uint32)
        np.testing.assert_array_equal(got, prng.random_bits(key, (n,), part))


def test_threefry_known_answers(mlb):
    """split(PRNGKey(0)) printed in the public JAX PRNG tutorial (non-partitionable)."""
    K = mlb.kernels
    kd = torch.zeros(2, dtype=torch.int32, device=DEV)
    got = K.threefry_split(kd, 2).cpu().numpy().view(np.uint32)
    np.testing.assert_array_equal(got, [[4146024105, 967050713], [2718843009, 1272950319]])


@pytest.mark.parametrize('J,E', [(1, 2), (2, 1), (7, 3), (1625, 2), (1626, 2), (8192, 4), (10000, 2),
                                 (65536, 2)])
@pytest.mark.parametrize('part', [False, True])
def test_ppo_permutations_bit_exact(mlb, J, E, part):
    K = mlb.kernels
    key = prng.key(1234 + J)
    kd = torch.from_numpy(key.view(np.int32).copy()).to(DEV)
    perm = K.ppo_permutations(kd, E, J, part).cpu().numpy()
    rnds, final = prng.update_epoch_keys(key, E, part)
    for e in range(E):
        np.testing.assert_array_equal(perm[e], prng.permutation(rnds[e], J, part))
        assert sorted(perm[e].tolist()) == list(range(J))
    np.testing.assert_array_equal(kd.cpu().numpy().view(np.uint32), final)


@pytest.mark.parametrize('C,Tp,B,M,leaf,dtype', [
    (1, 32, 64, 16, (8,), np.float32), (4, 8, 33, 12, (6,), np.int32), (2, 5, 17, 17, (1,), np.uint8),
    (3, 4, 10, 30, (), np.float32), (2, 3, 9, 5, (3,), np.float16), (1, 2, 5, 0, (4,), np.float32)])
def test_minibatch_gather_bit_exact(mlb, C, Tp, B, M, leaf, dtype):
    K = mlb.kernels
    rng = np.random.default_rng(C * 100 + B)
    store = rng.integers(0, 250, size=(C, Tp, 1, B, *leaf)).astype(dtype)
    J = C * B
    idx = rng.permutation(J)[:M].astype(np.int32) if M else np.zeros(0, np.int32)
    sd = _dev(store[:, :, 0])
    if M == 0:
        out = K.mb_gather(sd, torch.zeros(0, dtype=torch.int32, device=DEV), C, Tp, B)
        assert out.shape[1] == 0
        return
    out = K.mb_gather(sd, _dev(idx), C, Tp, B).cpu().numpy()
    ref = layouts.minibatch({'x': layouts.reorder_seq_data(store)[0]}, idx)['x']
    np.testing.assert_array_equal(out, ref)
    np.testing.assert_array_equal(out, layouts.minibatch_from_store(store, idx))
    rnn = rng.integers(0, 250, size=(C, 1, B, 7)).astype(np.float32)
    out = K.mb_gather_rnn(_dev(rnn[:, 0]), _dev(idx), C, B).cpu().numpy()
    np.testing.assert_array_equal(out, layouts.reorder_rnn_data(rnn)[0][idx])


@pytest.mark.parametrize('world,B,M,C', [(2, 64, 16, 1), (4, 96, 24, 2), (8, 2048, 512, 1), (8, 32, 8, 3), (3, 50, 25, 2)])
def test_dp_assign_minibatches_vs_oracle(mlb, world, B, M, C):
    """mlb_dp_assign_minibatches (owner-affine split of the global minibatches, index-exact data-parallel mode)
    for every rank of an emulated world against oracle/layouts.dp_assign_minibatch: exact id lists, the ranks'
    lists partition every global minibatch, own trajectories stay local."""
    from madrona_learn_b200._lib import c_int, c_ll, call, ptr
    from oracle import layouts
    rng = np.random.default_rng(world * 1000 + M)
    E, Jp = 3, C * world * B
    nmb = Jp // (world * M)
    perms = np.stack([rng.permutation(Jp) for _ in range(E)]).astype(np.int32)
    perms[1] = np.sort(perms[1])                       # worst case: every minibatch owned by one or two ranks
    perm_d = torch.from_numpy(perms).to(DEV)
    outs = []
    for r in range(world):
        out = torch.full((E, nmb, M), -1, dtype=torch.int32, device=DEV)
        call('mlb_dp_assign_minibatches', ptr(perm_d), c_ll(Jp), c_int(E), c_int(nmb), c_int(world), c_int(r),
             c_ll(B), c_ll(M), ptr(out))
        outs.append(out.cpu().numpy())
    remote = 0
    for e in range(E):
        for k in range(nmb):
            ids = perms[e, k * world * M:(k + 1) * world * M]
            ref = layouts.dp_assign_minibatch(ids, world, B, M)
            for r in range(world):
                np.testing.assert_array_equal(outs[r][e, k], ref[r])
                if e != 1:
                    remote += int(np.sum((ref[r] % (world * B)) // B != r))
            assert np.array_equal(np.sort(np.concatenate([outs[r][e, k] for r in range(world)])), np.sort(ids))
    if M >= 512:                                       # random permutations: only the binomial imbalance moves
        assert remote / (2 * nmb * world * M) < 0.06
