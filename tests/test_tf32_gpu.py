"""GPU parity of the tensor-core path for compute_dtype=float32 (mlb_gemm_tf32_tc: tcgen05.mma.kind::tf32 on the
fp32 activations / master weights, ml/cfg.py:96 + XLA:GPU's default f32 dot precision).

Tolerances: rel-L2 <= 1e-3 against the exact (fp64) product / oracle for the
TF32 path; <= 2e-6 against the product of the operands with their low 13 mantissa bits dropped (what the tensor
core reads), which pins the operand layouts, the transposes, the K tails and the split-K reduction exactly."""
import numpy as np
import pytest
import torch

from oracle import nn as onn, ppo as oppo

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30)


def _trunc_tf32(x):
    return (np.ascontiguousarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _rna_tf32(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


SHAPES = [
    # M, N, K, ta, tb, accumulate, bias
    (1000, 256, 256, 0, 0, 0, False),      # forward  Z = X W          (B stored [K, N]: MN-major)
    (1000, 256, 64, 0, 0, 0, False),       # first layer, K = obs_dim
    (777, 28, 256, 0, 0, 0, True),         # head forward: narrow N with a tail, bias, ragged M
    (1000, 256, 256, 0, 1, 0, False),      # dX = dZ W^T               (B stored [N, K]: K-major)
    (515, 256, 28, 0, 1, 0, False),        # dfeat = dhead W_h^T: K tail of 28 (< one k-block)
    (256, 256, 5000, 1, 0, 1, False),      # dW = X^T dZ: both MN-major, split-K reduction over rows
    (64, 256, 4099, 1, 0, 1, False),       # dW of the first layer: M = 64 < tile, ragged K
    (256, 28, 3000, 1, 0, 1, False),       # head dW
    (132, 100, 72, 1, 1, 0, True),         # transA with K-major B, ragged tiles
    (128, 1024, 256, 0, 1, 1, False),      # LSTM gate pre-activations (+= h W_h^T), N = 4 RH
    # more output tiles than SMs -> the persistent kernel (double-buffered TMEM accumulator, several tiles per CTA)
    (19277, 256, 256, 0, 0, 0, False),     # 151 row tiles, ragged last tile
    (19277, 256, 64, 0, 1, 0, False),
    (40000, 28, 256, 0, 0, 0, True),       # head forward at minibatch size: 313 tiles of BN = 32, bias
    (5000, 1024, 200, 0, 1, 0, True),      # 40 x 4 tiles: n-tiles of a row block back to back, K tail
    (19280, 100, 72, 1, 1, 0, True),       # MN-major A in the persistent kernel
]


@pytest.mark.parametrize('M,N,K,ta,tb,acc,with_bias', SHAPES)
def test_gemm_tf32_vs_exact_product(mlb, M, N, K, ta, tb, acc, with_bias):
    from madrona_learn_b200._lib import c_int, call, ptr
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32) if with_bias else None
    opA = lambda a: (a.T if ta else a).astype(np.float64)
    opB = lambda b: (b.T if tb else b).astype(np.float64)
    extra = (C0.astype(np.float64) if acc else 0.0) + (bias.astype(np.float64) if with_bias else 0.0)
    exact = opA(A) @ opB(B) + extra
    dropped = opA(_trunc_tf32(A)) @ opB(_trunc_tf32(B)) + extra
    rounded = opA(_rna_tf32(A)) @ opB(_rna_tf32(B)) + extra
    d = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x)).to(DEV)
    Ad, Bd, Cd, bd = d(A), d(B), d(C0 if acc else np.full((M, N), np.nan, np.float32)), d(bias)
    call('mlb_gemm_tf32_tc', ptr(Ad), ptr(Bd), ptr(Cd), ptr(bd), c_int(M), c_int(N), c_int(K),
         c_int(A.shape[1]), c_int(B.shape[1]), c_int(N), c_int(ta), c_int(tb), c_int(acc), c_int(1))
    torch.cuda.synchronize()
    out = Cd.cpu().numpy()
    assert np.isfinite(out).all()
    import os
    # MLB_TF32_ROUND: 0 = the tensor core's own conversion (low 13 bits dropped), 1 (default) = the activation
    # operand of the store products rounded to nearest in the kernel (A; B too when A is MN-major), reductions
    # (accumulate: the dW-type products) left to the hardware, 2 = both operands always
    mode = int(os.environ.get('MLB_TF32_ROUND', '1'))
    cvA = _rna_tf32 if (mode >= 2 or (mode == 1 and not acc)) else _trunc_tf32
    cvB = _rna_tf32 if (mode >= 2 or (mode == 1 and ta and not acc)) else _trunc_tf32
    model = opA(cvA(A)) @ opB(cvB(B)) + extra
    print('tf32 gemm rel-L2: exact %.3g dropped-bits %.3g rounded %.3g kernel-model %.3g' % (
        _rel(out, exact), _rel(out, dropped), _rel(out, rounded), _rel(out, model)))
    assert _rel(out, exact) < 1e-3, _rel(out, exact)
    assert _rel(out, model) < 2e-6, (_rel(out, dropped), _rel(out, rounded), _rel(out, model))


def test_gemm_tf32_contract(mlb):
    """Same refusals as mlb_gemm_f32 plus the TMA constraints (no fallback inside the entry point)."""
    from madrona_learn_b200 import _lib
    from madrona_learn_b200._lib import c_int, ptr
    L = _lib.lib()
    a = torch.zeros(64, 64, device=DEV)
    args = lambda lda=64, n=64, acc=0, sk=1: (None, ptr(a), ptr(a), ptr(a), ptr(None), c_int(64), c_int(n), c_int(64),
                                              c_int(lda), c_int(64), c_int(64), c_int(0), c_int(0), c_int(acc), c_int(sk))
    assert L.mlb_gemm_tf32_tc(*args()) == 0
    assert L.mlb_gemm_tf32_tc(*args(lda=66)) != 0             # row stride not a multiple of 16 bytes
    assert L.mlb_gemm_tf32_tc(*args(n=30)) != 0               # N % 4
    assert L.mlb_gemm_tf32_tc(*args(sk=4)) != 0               # split-K without accumulate
    assert L.mlb_gemm_tf32_ok(c_int(64), c_int(64), c_int(64), c_int(64), c_int(64), c_int(64), ptr(a), ptr(a), ptr(a)) == 1
    assert L.mlb_gemm_tf32_ok(c_int(64), c_int(64), c_int(64), c_int(66), c_int(64), c_int(64), ptr(a), ptr(a), ptr(a)) == 0
    torch.cuda.synchronize()


@pytest.fixture
def tf32(mlb):
    prev = mlb.matmul_precision()
    mlb.set_matmul_precision('tf32')
    yield mlb
    mlb.set_matmul_precision(prev)


BUCKETS = [4, 8, 5, 5, 2, 2]


@pytest.mark.parametrize('H,L,Tp,M', [(64, 2, 4, 256), (256, 3, 8, 512)])
def test_tf32_loss_and_grads_vs_oracle(tf32, H, L, Tp, M):
    """compute_dtype=float32 with matmul precision 'tf32': forward / PPO loss / backward of the whole network
    against the fp64 oracle (no operand rounding in the oracle): rel-L2 <= 1e-3 on the head; gradients: whole-vector
    cosine >= 0.9999, per-leaf rel-L2 <= 3e-2 (measured in the test output)."""
    import ctypes
    from madrona_learn_b200 import _lib
    from madrona_learn_b200._lib import c_float, c_int, c_ll, c_size_t, call, ptr
    from madrona_learn_b200.engine import PolicyProgram
    from oracle import algo_common as oac
    m = tf32
    D = 64
    rows, A = Tp * M, len(BUCKETS)
    rng = np.random.default_rng(21)
    p = onn.init_params(rng, D, H, L, BUCKETS)
    p['actor']['kernel'] = (rng.standard_normal(p['actor']['kernel'].shape) * 0.2).astype(np.float32)
    pol = m.Policy(actor_critic=m.ActorCritic(
        backbone=m.BackboneShared(prefix=None, encoder=m.BackboneEncoder(net=m.models.MLP(H, L))),
        actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(BUCKETS)), critic=m.models.DenseLayerCritic()))
    prog = PolicyProgram(pol.actor_critic, D, {'act': m.DiscreteActionsConfig(BUCKETS)}, DEV, torch.float32)
    prog.load_oracle_params(p)
    cfg = oppo.PPOCfg(BUCKETS, entropy_coef=0.02)
    mb = dict(obs=rng.standard_normal((Tp, M, D)).astype(np.float32),
              actions=np.stack([rng.integers(0, b, (Tp, M)) for b in BUCKETS], -1).astype(np.int32),
              advantages=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              returns=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              values=rng.standard_normal((Tp, M, 1)).astype(np.float32), mb_weights=np.ones((M, 1), np.float32))
    mb['log_probs'] = (-np.abs(rng.standard_normal((Tp, M, A))) - 0.5).astype(np.float32)
    ref = oppo.ppo_loss(p, mb, cfg, None, dtype=np.float64)
    dv = {k: torch.from_numpy(v).to(DEV) for k, v in mb.items()}
    obs_d = dv['obs'].view(rows, D)
    head = prog.forward_train(obs_d, rows)
    h = head.cpu().numpy()
    nA = sum(BUCKETS)
    assert _rel(h[:, :nA], ref['logits']) < 1e-3
    assert _rel(h[:, nA:nA + 1], ref['critic']) < 1e-3
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
    np.testing.assert_allclose(stt.loss, ref['loss'], rtol=2e-3, atol=1e-5)
    prog.backward(obs_d, rows)
    g = prog.to_oracle_params(prog.grads)
    # gradients: the same conditioning that makes the bf16 path 4e-2 (ReLU masks / LayerNorm-backward cancellation
    # amplify the forward's 1e-4..1e-3) at a mantissa 8x finer: measured 2e-2 on the worst leaf, 5e-3 typical
    worst = []
    onn.tree_map(lambda a, b: worst.append(_rel(a, b)), g, ref['grads'])
    ga = np.concatenate([np.ravel(x) for x in onn.tree_leaves(g)]) if hasattr(onn, 'tree_leaves') else None
    print('tf32 grads rel-L2 per leaf: max %.3g median %.3g' % (max(worst), float(np.median(worst))))
    assert max(worst) < 3e-2 and np.median(worst) < 8e-3, worst
    flat_a, flat_b = [], []
    onn.tree_map(lambda a, b: (flat_a.append(np.ravel(a)), flat_b.append(np.ravel(b))), g, ref['grads'])
    fa, fb = np.concatenate(flat_a).astype(np.float64), np.concatenate(flat_b).astype(np.float64)
    cos = float(fa @ fb / (np.linalg.norm(fa) * np.linalg.norm(fb)))
    print('tf32 whole-gradient cosine %.6f rel-L2 %.3g' % (cos, _rel(fa, fb)))
    assert cos > 0.9999, cos


def test_tf32_update_iter_tracks_exact_fp32(mlb):
    """update_iter (rollout, GAE, 2 epochs x 2 minibatches, CUDA graph replay) with 'tf32' against 'highest':
    same seeds, the parameters after three updates agree to TF32 precision and the loss is finite."""
    m = mlb
    N, T, D = 512, 8, 64
    out = {}
    prev = m.matmul_precision()
    try:
        for prec in ('highest', 'tf32'):
            m.set_matmul_precision(prec)
            env = m.SyntheticVectorEnv(N, D, len(BUCKETS), seed=2, device=DEV)
            pol = m.Policy(actor_critic=m.ActorCritic(
                backbone=m.BackboneShared(prefix=None, encoder=m.BackboneEncoder(net=m.models.MLP(128, 2))),
                actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(BUCKETS)),
                critic=m.models.DenseLayerCritic()))
            cfg = m.TrainConfig(
                num_worlds=N, num_agents_per_world=1, num_updates=4, actions={'act': m.DiscreteActionsConfig(BUCKETS)},
                steps_per_update=T, lr=3e-4,
                algo=m.PPOConfig(num_epochs=2, minibatch_size=N // 2, clip_coef=0.2, value_loss_coef=0.5,
                                 entropy_coef={'act': 0.01}, max_grad_norm=0.5),
                num_bptt_chunks=1, gamma=0.99, seed=3, metrics_buffer_size=4, gae_lambda=0.95,
                dreamer_v3_critic=False)
            mgr = m.init_training(DEV, cfg, env.sim_fns(), pol, None, verbose=False)
            p0 = mgr.state.policy_states.program.params.clone()
            for _ in range(3):
                mgr.update_iter()
            torch.cuda.synchronize()
            out[prec] = (mgr.state.policy_states.program.params.double().cpu().numpy() - p0.double().cpu().numpy(),
                         mgr.metrics.latest()['Loss'].mean)
    finally:
        m.set_matmul_precision(prev)
    assert np.isfinite(out['tf32'][1])
    # sampled actions may differ where two logits are within TF32 error of each other, so compare the update
    # direction, not bits: cosine of the parameter deltas
    a, b = out['highest'][0], out['tf32'][0]
    cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))
    assert cos > 0.98, cos
    np.testing.assert_allclose(out['tf32'][1], out['highest'][1], rtol=0.05, atol=5e-3)


@pytest.mark.parametrize('rows,K,H,with_z', [(1000, 64, 256, True), (40000, 256, 256, True), (8192, 256, 256, False),
                                              (19277, 128, 128, True), (777, 64, 64, True)])
def test_dense_ln_relu_fwd_tf32_vs_unfused(mlb, rows, K, H, with_z):
    """mlb_dense_ln_relu_fwd_tf32 (LayerNorm + ReLU in the TF32 GEMM's epilogue) against the exact fp64 layer and
    against its own pieces: z equals mlb_gemm_tf32_tc bit for bit (same MMA schedule), y / stats equal
    mlb_ln_relu_fwd_f32 applied to that z up to the summation order of the row statistics."""
    from madrona_learn_b200._lib import c_int, c_ll, call, ptr
    rng = np.random.default_rng(rows + H)
    x = rng.standard_normal((rows, K)).astype(np.float32)
    w = (rng.standard_normal((K, H)) / np.sqrt(K)).astype(np.float32)
    sc = (1 + 0.1 * rng.standard_normal(H)).astype(np.float32)
    bi = (0.1 * rng.standard_normal(H)).astype(np.float32)
    d = lambda a: torch.from_numpy(a).to(DEV)
    xd, wd, scd, bid = d(x), d(w), d(sc), d(bi)
    z = torch.full((rows, H), float('nan'), device=DEV) if with_z else None
    y = torch.full((rows, H), float('nan'), device=DEV)
    st = torch.full((rows, 2), float('nan'), device=DEV)
    call('mlb_dense_ln_relu_fwd_tf32', ptr(xd), ptr(wd), ptr(scd), ptr(bid), ptr(z), ptr(y), ptr(st), c_ll(rows),
         c_int(K), c_int(H), c_int(K))
    z2 = torch.empty(rows, H, device=DEV)
    y2 = torch.empty(rows, H, device=DEV)
    st2 = torch.empty(rows, 2, device=DEV)
    call('mlb_gemm_tf32_tc', ptr(xd), ptr(wd), ptr(z2), ptr(None), c_int(rows), c_int(H), c_int(K), c_int(K), c_int(H),
         c_int(H), c_int(0), c_int(0), c_int(0), c_int(1))
    call('mlb_ln_relu_fwd_f32', ptr(z2), ptr(scd), ptr(bid), ptr(y2), ptr(st2), c_ll(rows), c_int(H))
    torch.cuda.synchronize()
    if with_z:
        assert torch.equal(z, z2)
    assert bool(torch.isfinite(y).all()) and bool(torch.isfinite(st).all())
    np.testing.assert_allclose(y.cpu().numpy(), y2.cpu().numpy(), rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(st.cpu().numpy(), st2.cpu().numpy(), rtol=2e-5, atol=2e-6)
    z64 = x.astype(np.float64) @ w.astype(np.float64)
    mean = z64.mean(-1, keepdims=True)
    var = np.maximum(0, (z64 * z64).mean(-1, keepdims=True) - mean * mean)
    yref = np.maximum(0, (z64 - mean) / np.sqrt(var + 1e-6) * sc + bi)
    assert _rel(y.cpu().numpy(), yref) < 1e-3


def test_tf32_recurrent_update_tracks_exact_fp32(mlb):
    """compute_dtype=float32 with an LSTM encoder under the default 'tf32' precision: the recurrent / input / BPTT
    products (accumulating GEMMs with N = 4 RH, MN-major reductions over T' x M rows) run on mlb_gemm_tf32_tc too.
    Same seeds as an exact-fp32 run: the parameter update points the same way and the loss agrees."""
    m = mlb
    N, T, D, C = 256, 16, 32, 2
    out = {}
    prev = m.matmul_precision()
    try:
        for prec in ('highest', 'tf32'):
            m.set_matmul_precision(prec)
            env = m.SyntheticVectorEnv(N, D, len(BUCKETS), seed=2, p_done=1 / 16, device=DEV)
            pol = m.Policy(actor_critic=m.ActorCritic(
                backbone=m.BackboneShared(prefix=None, encoder=m.RecurrentBackboneEncoder(
                    net=m.models.MLP(64, 1), rnn=m.rnn.LSTM(64, 2))),
                actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(BUCKETS)),
                critic=m.models.DenseLayerCritic()))
            cfg = m.TrainConfig(
                num_worlds=N, num_agents_per_world=1, num_updates=4, actions={'act': m.DiscreteActionsConfig(BUCKETS)},
                steps_per_update=T, lr=3e-4,
                algo=m.PPOConfig(num_epochs=2, minibatch_size=N * C // 2, clip_coef=0.2, value_loss_coef=0.5,
                                 entropy_coef={'act': 0.01}, max_grad_norm=0.5),
                num_bptt_chunks=C, gamma=0.99, seed=3, metrics_buffer_size=4, gae_lambda=0.95,
                dreamer_v3_critic=False, normalize_values=True)
            mgr = m.init_training(DEV, cfg, env.sim_fns(), pol, None, verbose=False)
            p0 = mgr.state.policy_states.program.params.clone()
            mgr.update_iter()
            torch.cuda.synchronize()
            out[prec] = (mgr.state.policy_states.program.params.double().cpu().numpy() - p0.double().cpu().numpy(),
                         mgr.metrics.latest()['Loss'].mean)
    finally:
        m.set_matmul_precision(prev)
    a, b = out['highest'][0], out['tf32'][0]
    cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))
    assert np.isfinite(out['tf32'][1]) and cos > 0.98, (cos, out)
    np.testing.assert_allclose(out['tf32'][1], out['highest'][1], rtol=0.05, atol=5e-3)
