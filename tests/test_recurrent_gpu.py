"""GPU parity of the recurrent (LSTM) actor-critic path, BASELINE config 4 family:
RecurrentBackboneEncoder(MLP, LSTM) rollout step, BPTT minibatch forward/backward with
sequence-end resets and cached chunk start states, value-normaliser EMA -- vs the oracle
(whose LSTM BPTT is itself validated against torch.autograd).  fp32 path, rel-L2 1e-4."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import algo_common as oac
from oracle import env as oenv
from oracle import layouts, nn as onn, ppo as oppo
from oracle.moving_avg import EMANormalizer as OEMA

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
BUCKETS = [4, 8, 5, 5, 2, 2]


def _rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30)


def _policy(m, H, L, RH, RL=1):
    return m.Policy(actor_critic=m.ActorCritic(
        backbone=m.BackboneShared(prefix=None, encoder=m.RecurrentBackboneEncoder(
            net=m.models.MLP(H, L), rnn=m.rnn.LSTM(RH, RL))),
        actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(BUCKETS)),
        critic=m.models.DenseLayerCritic()))


def test_lstm_forward_backward_vs_oracle(mlb):
    from madrona_learn_b200._lib import c_float, c_int, c_ll, c_size_t, call, ptr
    from madrona_learn_b200.engine import PolicyProgram
    m = mlb
    D, H, L, RH, Tp, M = 16, 32, 2, 24, 7, 40
    rows, A = Tp * M, len(BUCKETS)
    rng = np.random.default_rng(0)
    p = onn.init_params(rng, D, H, L, BUCKETS, lstm_hidden=RH, lstm_layers=1)
    p['actor']['kernel'] = (rng.standard_normal(p['actor']['kernel'].shape) * 0.3).astype(np.float32)
    p['lstm'][0]['bh'] = (0.1 * rng.standard_normal(4 * RH)).astype(np.float32)
    prog = PolicyProgram(_policy(m, H, L, RH).actor_critic, D, {'act': m.DiscreteActionsConfig(BUCKETS)}, DEV)
    prog.load_oracle_params(p)
    back = prog.to_oracle_params()
    onn.tree_map(lambda a, b: np.testing.assert_array_equal(a, b), back['lstm'], p['lstm'])
    cfg = oppo.PPOCfg(BUCKETS, entropy_coef=0.02)
    c0 = rng.standard_normal((M, RH)).astype(np.float32)
    h0 = rng.standard_normal((M, RH)).astype(np.float32) * 0.5
    mb = dict(obs=rng.standard_normal((Tp, M, D)).astype(np.float32),
              actions=np.stack([rng.integers(0, b, (Tp, M)) for b in BUCKETS], -1).astype(np.int32),
              advantages=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              returns=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              values=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              dones=(rng.random((Tp, M, 1)) < 0.2), mb_weights=np.ones((M, 1), np.float32),
              rnn_start_states=([c0], [h0]))
    mb['log_probs'] = (-np.abs(rng.standard_normal((Tp, M, A))) - 0.5).astype(np.float32)
    ref = oppo.ppo_loss(p, mb, cfg, None, dtype=np.float64)

    dv = {k: torch.from_numpy(np.ascontiguousarray(v)).to(DEV) for k, v in mb.items() if k != 'rnn_start_states'}
    seq = dict(Tp=Tp, M=M, ends=dv['dones'].view(torch.uint8).view(Tp, M), c0=torch.from_numpy(c0).to(DEV),
               h0=torch.from_numpy(h0).to(DEV))
    obs_d = dv['obs'].view(rows, D)
    head = prog.forward_train(obs_d, rows, seq)
    h = head.cpu().numpy()
    assert _rel(h[:, :26], ref['logits']) < 1e-4 and _rel(h[:, 26:27], ref['critic']) < 1e-4
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
         ptr(tw['stats_out']), ptr(tw['loss_ws']), c_size_t(tw['loss_ws'].numel()), None, c_int(0))
    prog.backward(obs_d, rows, seq)
    g = prog.to_oracle_params(prog.grads)
    onn.tree_map(lambda a, b: np.testing.assert_array_less(_rel(a, b), 2e-4), g, ref['grads'])
    # rollout-mode single step == first step of the sequence (no reset inside the step)
    states = ([torch.from_numpy(c0).to(DEV).clone()], [torch.from_numpy(h0).to(DEV).clone()])
    h1 = prog.forward_infer(dv['obs'][0].contiguous(), M, states).cpu().numpy()
    np.testing.assert_allclose(h1[:, :27], h[:M, :27], rtol=1e-4, atol=1e-5)
    cs, hs, out, _ = onn.lstm_step([c0.astype(np.float64)], [h0.astype(np.float64)],
                                   onn.mlp_fwd(mb['obs'][0].astype(np.float64), onn.cast_tree(p['mlp'], np.float64))[0],
                                   onn.cast_tree(p['lstm'], np.float64))
    np.testing.assert_allclose(states[0][0].cpu().numpy(), cs[0], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(states[1][0].cpu().numpy(), hs[0], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('D,H,L,RH,Tp,M', [(32, 128, 2, 64, 6, 300), (64, 256, 2, 256, 8, 2048)])
def test_lstm_tc_forward_backward_vs_oracle(mlb, D, H, L, RH, Tp, M):
    """compute_dtype = bfloat16: the LSTM's products run on tcgen05 (bf16 operands, fp32 accumulation and
    cell state).  Bounds vs the EXACT fp64 oracle: the SURVEY 8c bf16 tolerance (heads rel-L2 <= 2e-2, gradient
    cosine >= 0.999 / rel-L2 <= 5e-2 -- the recurrence compounds the bf16 rounding of h over T' steps)."""
    from madrona_learn_b200._lib import c_float, c_int, c_ll, c_size_t, call, ptr
    from madrona_learn_b200.engine import PolicyProgram
    m = mlb
    rows, A = Tp * M, len(BUCKETS)
    rng = np.random.default_rng(RH)
    p = onn.init_params(rng, D, H, L, BUCKETS, lstm_hidden=RH, lstm_layers=1)
    p['actor']['kernel'] = (rng.standard_normal(p['actor']['kernel'].shape) * 0.2).astype(np.float32)
    p['lstm'][0]['bh'] = (0.1 * rng.standard_normal(4 * RH)).astype(np.float32)
    prog = PolicyProgram(_policy(m, H, L, RH).actor_critic, D, {'act': m.DiscreteActionsConfig(BUCKETS)}, DEV,
                         torch.bfloat16)
    assert prog.tc and prog.lstm.tc
    prog.load_oracle_params(p)
    cfg = oppo.PPOCfg(BUCKETS, entropy_coef=0.02)
    c0 = rng.standard_normal((M, RH)).astype(np.float32)
    h0 = rng.standard_normal((M, RH)).astype(np.float32) * 0.5
    mb = dict(obs=rng.standard_normal((Tp, M, D)).astype(np.float32),
              actions=np.stack([rng.integers(0, b, (Tp, M)) for b in BUCKETS], -1).astype(np.int32),
              advantages=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              returns=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              values=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              dones=(rng.random((Tp, M, 1)) < 0.2), mb_weights=np.ones((M, 1), np.float32),
              rnn_start_states=([c0], [h0]))
    p64 = onn.cast_tree(p, np.float64)
    mb['log_probs'] = np.zeros((Tp, M, A), np.float32)
    ref0 = oppo.ppo_loss(p, mb, cfg, None, dtype=np.float64)
    lp0, _ = onn.action_stats(ref0['logits'], mb['actions'].reshape(rows, A), BUCKETS)
    mb['log_probs'] = lp0.reshape(Tp, M, A).astype(np.float32)      # old policy = current policy (first minibatch)
    ref = oppo.ppo_loss(p, mb, cfg, None, dtype=np.float64)
    dv = {k: torch.from_numpy(np.ascontiguousarray(v)).to(DEV) for k, v in mb.items() if k != 'rnn_start_states'}
    seq = dict(Tp=Tp, M=M, ends=dv['dones'].view(torch.uint8).view(Tp, M), c0=torch.from_numpy(c0).to(DEV),
               h0=torch.from_numpy(h0).to(DEV))
    obs_d = dv['obs'].view(rows, D)
    head = prog.forward_train(obs_d, rows, seq)
    h = head.cpu().numpy()
    assert _rel(h[:, :26], ref['logits']) < 2e-2 and _rel(h[:, 26:27], ref['critic']) < 2e-2
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
         ptr(tw['stats_out']), ptr(tw['loss_ws']), c_size_t(tw['loss_ws'].numel()), None, c_int(0))
    prog.backward(obs_d, rows, seq)
    g = prog.to_oracle_params(prog.grads)
    flat = lambda t: np.concatenate([np.asarray(x, np.float64).reshape(-1) for k in ('mlp', 'lstm', 'actor', 'critic')
                                     for x in onn.tree_leaves(t[k])])
    a, b = flat(g), flat(ref['grads'])
    cos = a @ b / (np.linalg.norm(a) * np.linalg.norm(b))
    rel = np.linalg.norm(a - b) / np.linalg.norm(b)
    print('PARITY lstm_tc', dict(RH=RH, rows=rows, grad_cos=float(cos), grad_rel=float(rel),
                                 logits_rel=float(_rel(h[:, :26], ref['logits']))))
    assert cos >= 0.999 and rel <= 5e-2, (cos, rel)
    for name in ('wi', 'wh', 'bh'):
        assert _rel(g['lstm'][0][name], ref['grads']['lstm'][0][name]) < 6e-2, name
    # rollout-mode single step == first step of the sequence
    states = ([torch.from_numpy(c0).to(DEV).clone()], [torch.from_numpy(h0).to(DEV).clone()])
    h1 = prog.forward_infer(dv['obs'][0].contiguous(), M, states).cpu().numpy()
    np.testing.assert_allclose(h1[:, :27], h[:M, :27], rtol=2e-2, atol=2e-3)
    cs, hs, out, _ = onn.lstm_step([c0.astype(np.float64)], [h0.astype(np.float64)],
                                   onn.mlp_fwd(mb['obs'][0].astype(np.float64), p64['mlp'])[0], p64['lstm'])
    assert _rel(states[0][0].cpu().numpy(), cs[0]) < 1e-2 and _rel(states[1][0].cpu().numpy(), hs[0]) < 1e-2
    # the optimiser refreshes the LSTM's bf16 operand copies
    prog.optimizer_step(3e-4, 0.5)
    wi, wh, _ = prog.lstm.views(prog.params)
    assert torch.equal(prog.lstm.wi_c, wi.to(torch.bfloat16)) and torch.equal(prog.lstm.wi_t, wi.t().contiguous().to(torch.bfloat16))
    assert torch.equal(prog.lstm.wh_c, wh.to(torch.bfloat16))


def test_recurrent_tc_update_iter_tracks_fp32(mlb, monkeypatch):
    """config-4 family through the public API on the tensor-core path: graph capture + replay, parameters
    track the fp32 run of the same seeds to bf16 precision."""
    m = mlb
    N, T, C, M, E, D, H, L, RH = 512, 16, 2, 256, 2, 32, 128, 2, 64
    out = {}
    for dt in (torch.float32, torch.bfloat16):
        env = m.SyntheticVectorEnv(N, D, len(BUCKETS), seed=11, p_done=0.1, device=DEV)
        cfg = m.TrainConfig(
            num_worlds=N, num_agents_per_world=1, num_updates=10, actions={'act': m.DiscreteActionsConfig(BUCKETS)},
            steps_per_update=T, lr=3e-4,
            algo=m.PPOConfig(num_epochs=E, minibatch_size=M, clip_coef=0.2, value_loss_coef=0.5,
                             entropy_coef={'act': 0.01}, max_grad_norm=0.5),
            num_bptt_chunks=C, gamma=0.99, seed=5, metrics_buffer_size=4, gae_lambda=0.95,
            dreamer_v3_critic=False, normalize_values=True, compute_dtype=dt)
        mgr = m.init_training(DEV, cfg, env.sim_fns(), _policy(m, H, L, RH), None, verbose=False)
        prog = mgr.state.policy_states.program
        flat = lambda: np.concatenate([np.asarray(x, np.float64).reshape(-1) for k in ('mlp', 'lstm', 'actor', 'critic')
                                       for x in onn.tree_leaves(prog.to_oracle_params()[k])])
        p0 = flat()                          # (the arenas differ: the tensor-core path pads the head to 64 columns)
        for _ in range(3):
            mgr.update_iter()
        torch.cuda.synchronize()
        out[dt] = flat() - p0, mgr.metrics.latest()
        assert np.isfinite(out[dt][1]['Loss'].mean)
    d32, dbf = out[torch.float32][0], out[torch.bfloat16][0]
    cos = float(d32 @ dbf / (np.linalg.norm(d32) * np.linalg.norm(dbf)))
    print('PARITY lstm_tc_update', dict(delta_cos=cos))
    assert cos > 0.9, cos
    assert abs(out[torch.bfloat16][1]['Entropy'].mean - out[torch.float32][1]['Entropy'].mean) < 0.05


@pytest.mark.parametrize('dtype,D,H,L,RH,RL,Tp,M', [(torch.float32, 16, 32, 2, 24, 2, 7, 40),
                                                    (torch.float32, 16, 32, 1, 16, 3, 5, 33),
                                                    (torch.bfloat16, 32, 128, 2, 64, 2, 6, 300)])
def test_multilayer_lstm_forward_backward_vs_oracle(mlb, dtype, D, H, L, RH, RL, Tp, M):
    """LSTM(num_layers > 1) (ml/rnn.py:10-45): layer l is fed by the hidden state of layer l - 1, the heads read
    the concatenation of every layer's hidden state.  fp32 path rel-L2 2e-4 per tensor; tensor-core path the
    bf16 tolerance of the single-layer test."""
    from madrona_learn_b200._lib import c_float, c_int, c_ll, c_size_t, call, ptr
    from madrona_learn_b200.engine import PolicyProgram
    m = mlb
    rows, A = Tp * M, len(BUCKETS)
    rng = np.random.default_rng(RL * 100 + RH)
    p = onn.init_params(rng, D, H, L, BUCKETS, lstm_hidden=RH, lstm_layers=RL)
    p['actor']['kernel'] = (rng.standard_normal(p['actor']['kernel'].shape) * 0.3).astype(np.float32)
    for lyr in p['lstm']:
        lyr['bh'] = (0.1 * rng.standard_normal(4 * RH)).astype(np.float32)
    prog = PolicyProgram(_policy(m, H, L, RH, RL).actor_critic, D, {'act': m.DiscreteActionsConfig(BUCKETS)}, DEV, dtype)
    assert prog.lstm.RL == RL and prog.feat == RL * RH
    prog.load_oracle_params(p)
    back = prog.to_oracle_params()
    onn.tree_map(lambda a, b: np.testing.assert_array_equal(a, b), back['lstm'], p['lstm'])
    cfg = oppo.PPOCfg(BUCKETS, entropy_coef=0.02)
    c0 = [rng.standard_normal((M, RH)).astype(np.float32) for _ in range(RL)]
    h0 = [(rng.standard_normal((M, RH)) * 0.5).astype(np.float32) for _ in range(RL)]
    mb = dict(obs=rng.standard_normal((Tp, M, D)).astype(np.float32),
              actions=np.stack([rng.integers(0, b, (Tp, M)) for b in BUCKETS], -1).astype(np.int32),
              advantages=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              returns=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              values=rng.standard_normal((Tp, M, 1)).astype(np.float32),
              dones=(rng.random((Tp, M, 1)) < 0.2), mb_weights=np.ones((M, 1), np.float32),
              rnn_start_states=(c0, h0))
    mb['log_probs'] = np.zeros((Tp, M, A), np.float32)
    ref0 = oppo.ppo_loss(p, mb, cfg, None, dtype=np.float64)
    lp0, _ = onn.action_stats(ref0['logits'], mb['actions'].reshape(rows, A), BUCKETS)
    mb['log_probs'] = (lp0.reshape(Tp, M, A) + (0.0 if dtype == torch.bfloat16 else 0.3) *
                       rng.standard_normal((Tp, M, A))).astype(np.float32)
    ref = oppo.ppo_loss(p, mb, cfg, None, dtype=np.float64)
    dv = {k: torch.from_numpy(np.ascontiguousarray(v)).to(DEV) for k, v in mb.items() if k != 'rnn_start_states'}
    td = lambda x: torch.from_numpy(x).to(DEV)
    # chunk-start states as the PPO loop hands them over: per-layer feature slices of one [M, RL*RH] buffer
    cc, hh = td(np.concatenate(c0, -1)), td(np.concatenate(h0, -1))
    seq = dict(Tp=Tp, M=M, ends=dv['dones'].view(torch.uint8).view(Tp, M),
               c0=[cc[:, l * RH:(l + 1) * RH] for l in range(RL)], h0=[hh[:, l * RH:(l + 1) * RH] for l in range(RL)])
    obs_d = dv['obs'].view(rows, D)
    head = prog.forward_train(obs_d, rows, seq)
    h = head.cpu().numpy()
    tc = dtype == torch.bfloat16
    assert _rel(h[:, :26], ref['logits']) < (2e-2 if tc else 1e-4) and _rel(h[:, 26:27], ref['critic']) < (2e-2 if tc else 1e-4)
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
         ptr(tw['stats_out']), ptr(tw['loss_ws']), c_size_t(tw['loss_ws'].numel()), None, c_int(0))
    prog.backward(obs_d, rows, seq)
    g = prog.to_oracle_params(prog.grads)
    if tc:
        flat = lambda t: np.concatenate([np.asarray(x, np.float64).reshape(-1) for k in ('mlp', 'lstm', 'actor', 'critic')
                                         for x in onn.tree_leaves(t[k])])
        a, b = flat(g), flat(ref['grads'])
        cos, rel = a @ b / (np.linalg.norm(a) * np.linalg.norm(b)), np.linalg.norm(a - b) / np.linalg.norm(b)
        print('PARITY lstm_tc_multilayer', dict(RL=RL, grad_cos=float(cos), grad_rel=float(rel)))
        assert cos >= 0.999 and rel <= 5e-2, (cos, rel)
    else:
        onn.tree_map(lambda a, b: np.testing.assert_array_less(_rel(a, b), 2e-4), g, ref['grads'])
    # rollout-mode single step == first step of the sequence, for every layer's state
    states = ([td(x).clone() for x in c0], [td(x).clone() for x in h0])
    h1 = prog.forward_infer(dv['obs'][0].contiguous(), M, states).cpu().numpy()
    np.testing.assert_allclose(h1[:, :27], h[:M, :27], rtol=2e-2 if tc else 1e-4, atol=2e-3 if tc else 1e-5)
    p64 = onn.cast_tree(p, np.float64)
    cs, hs, out, _ = onn.lstm_step([x.astype(np.float64) for x in c0], [x.astype(np.float64) for x in h0],
                                   onn.mlp_fwd(mb['obs'][0].astype(np.float64), p64['mlp'])[0], p64['lstm'])
    for l in range(RL):
        assert _rel(states[0][l].cpu().numpy(), cs[l]) < (1e-2 if tc else 1e-4)
        assert _rel(states[1][l].cpu().numpy(), hs[l]) < (1e-2 if tc else 1e-4)


def test_multilayer_recurrent_update_iter(mlb):
    """Two-layer LSTM through the public API: per-layer chunk-start states cached in the store, gathered per
    minibatch, BPTT, per-gate re-projection of both layers; graph capture and replay; both compute dtypes."""
    m = mlb
    N, T, C, M, D, H, L, RH, RL = 256, 16, 2, 128, 32, 64, 1, 64, 2
    for dt in (torch.float32, torch.bfloat16):
        env = m.SyntheticVectorEnv(N, D, len(BUCKETS), seed=11, p_done=0.1, device=DEV)
        cfg = m.TrainConfig(
            num_worlds=N, num_agents_per_world=1, num_updates=10, actions={'act': m.DiscreteActionsConfig(BUCKETS)},
            steps_per_update=T, lr=3e-4,
            algo=m.PPOConfig(num_epochs=2, minibatch_size=M, clip_coef=0.2, value_loss_coef=0.5,
                             entropy_coef={'act': 0.01}, max_grad_norm=0.5),
            num_bptt_chunks=C, gamma=0.99, seed=5, metrics_buffer_size=4, gae_lambda=0.95,
            dreamer_v3_critic=False, normalize_values=True, compute_dtype=dt)
        mgr = m.init_training(DEV, cfg, env.sim_fns(), _policy(m, H, L, RH, RL), None, verbose=False)
        prog = mgr.state.policy_states.program
        n0 = {k: v for k, v in prog.initial_weight_norms.items() if k.startswith('lstm')}
        assert len(n0) == 8 * RL
        for _ in range(3):
            mgr.update_iter()
        torch.cuda.synchronize()
        assert np.isfinite(mgr.metrics.latest()['Loss'].mean)
        assert mgr.rollout_mgr.store['rnn_start_c'].shape[-1] == RL * RH
        assert float(mgr.rollout_mgr.store['rnn_start_h'][1].abs().sum()) > 0      # second chunk starts from a live carry
        got = prog.to_oracle_params()
        for li in range(RL):                      # every gate kernel of every layer keeps its own initial norm
            for kk, pre in (('wi', 'i'), ('wh', 'h')):
                for g, gate in enumerate('ifgo'):
                    blk = got['lstm'][li][kk][:, g * RH:(g + 1) * RH]
                    np.testing.assert_allclose(np.linalg.norm(blk), n0[f'lstm{li}/{pre}{gate}'], rtol=2e-5)


def test_recurrent_update_iter_matches_oracle(mlb, monkeypatch):
    """config-4 family end to end: 3 BPTT chunks, LSTM, normalize_values=True."""
    monkeypatch.setenv('MLB_CUDA_GRAPH', '0')
    m = mlb
    N, T, C, M, E, D, H, L, RH = 24, 12, 3, 18, 2, 8, 32, 1, 16
    env = m.SyntheticVectorEnv(N, D, len(BUCKETS), seed=11, p_done=0.15, device=DEV)
    cfg = m.TrainConfig(
        num_worlds=N, num_agents_per_world=1, num_updates=10, actions={'act': m.DiscreteActionsConfig(BUCKETS)},
        steps_per_update=T, lr=3e-4,
        algo=m.PPOConfig(num_epochs=E, minibatch_size=M, clip_coef=0.2, value_loss_coef=0.5,
                         entropy_coef={'act': 0.01}, max_grad_norm=0.5),
        num_bptt_chunks=C, gamma=0.99, seed=5, metrics_buffer_size=4, gae_lambda=0.95,
        dreamer_v3_critic=False, normalize_values=True)
    mgr = m.init_training(DEV, cfg, env.sim_fns(), _policy(m, H, L, RH), None, verbose=False)
    prog = mgr.state.policy_states.program
    Tp, A = T // C, len(BUCKETS)
    p0 = prog.to_oracle_params()
    key0 = mgr.state.train_states.update_prng_key.cpu().numpy().view(np.uint32).copy()
    h = mgr.state.train_states.value_normalizer_state.cpu().numpy()
    vn0 = dict(mu=h[0:1].copy(), inv_sigma=h[1:2].copy(), sigma=h[2:3].copy(), mu_biased=h[3:4].copy(),
               sigma_sq_biased=h[4:5].copy(), N=np.int32(0))
    mgr.update_iter()
    torch.cuda.synchronize()
    st = {k: v.cpu().numpy() for k, v in mgr.rollout_mgr.store.items()}
    boot = mgr.rollout_mgr.bootstrap.cpu().numpy()
    # replay the recurrent rollout with the oracle: per-step values/log-probs, chunk start states
    p64 = onn.cast_tree(p0, np.float64)
    obs_seq, act_seq = st['obs'].reshape(T, N, D), st['actions'].reshape(T, N, A)
    dones = st['dones'].reshape(T, N)
    cs, hs = [np.zeros((N, RH))], [np.zeros((N, RH))]
    for t in range(T):
        if t % Tp == 0:
            np.testing.assert_allclose(st['rnn_start_c'][t // Tp, 0], cs[0], rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(st['rnn_start_h'][t // Tp, 0], hs[0], rtol=1e-4, atol=1e-5)
        feat, _ = onn.mlp_fwd(obs_seq[t].astype(np.float64), p64['mlp'])
        cs, hs, out, _ = onn.lstm_step(cs, hs, feat, p64['lstm'])
        logits, critic = onn.heads_fwd(out, p64)
        lp, _ = onn.action_stats(logits, act_seq[t], BUCKETS)
        np.testing.assert_allclose(st['log_probs'].reshape(T, N, A)[t], lp, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(st['values'].reshape(T, N)[t], critic[:, 0], rtol=1e-4, atol=1e-5)
        keep = (~dones[t])[:, None]
        cs, hs = [cs[0] * keep], [hs[0] * keep]
    feat, _ = onn.mlp_fwd(mgr.rollout.cur_obs['obs'].cpu().numpy().astype(np.float64), p64['mlp'])
    _, _, out, _ = onn.lstm_step(cs, hs, feat, p64['lstm'])
    np.testing.assert_allclose(boot.reshape(N), onn.heads_fwd(out, p64)[1][:, 0], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(mgr.rollout.rnn_states[0][0].cpu().numpy(), cs[0], rtol=1e-4, atol=1e-5)
    # GAE on the stored buffers (bit-exact), with the value normaliser inverted on load
    on = OEMA(cfg.value_normalizer_decay)
    vals, bvals = on.invert(vn0, st['values']), on.invert(vn0, boot)
    adv = oac.compute_advantages(cfg.gamma, cfg.gae_lambda, st['rewards'], vals, st['dones'], bvals)
    np.testing.assert_array_equal(st['advantages'], adv)
    # the PPO update (BPTT minibatches) re-run by the oracle
    ocfg = oppo.PPOCfg(BUCKETS, num_epochs=E, minibatch_size=M, entropy_coef=0.01, lr=cfg.lr,
                       normalize_values=True)
    roll = {k: layouts.reorder_seq_data(st[k])[0] for k in
            ('obs', 'actions', 'log_probs', 'advantages', 'returns', 'values', 'dones')}
    roll['rnn_start_states'] = ([layouts.reorder_rnn_data(st['rnn_start_c'])[0]],
                                [layouts.reorder_rnn_data(st['rnn_start_h'])[0]])
    opt, norms = oppo.adam_init(p0), oppo.initial_weight_norms(p0)
    p1, opt1, key1, vn1, last, perms = oppo.ppo_update(p0, opt, norms, roll, ocfg, key0, vn0, dtype=np.float32)
    np.testing.assert_array_equal(mgr.ppo_ws.perm.cpu().numpy(), perms)
    got = prog.to_oracle_params()
    num = sum(float(np.sum(np.square(a.astype(np.float64) - b))) for a, b in
              zip(onn.tree_leaves(got), onn.tree_leaves(p1)))
    den = sum(float(np.sum(np.square(a.astype(np.float64) - b))) for a, b in
              zip(onn.tree_leaves(p1), onn.tree_leaves(p0)))
    assert np.sqrt(num / den) < 2e-2, np.sqrt(num / den)
    onn.tree_map(lambda a, b: np.testing.assert_allclose(a, b, atol=4 * cfg.lr), got, p1)
    # every LSTM gate kernel keeps its own initial norm (ml/ppo.py:303-310 per `kernel` leaf)
    for kk in ('wi', 'wh'):
        for g in range(4):
            blk = got['lstm'][0][kk][:, g * RH:(g + 1) * RH]
            np.testing.assert_allclose(np.linalg.norm(blk), norms['lstm'][0][kk][g], rtol=1e-5)
    # graph-captured replay keeps working with the recurrent state
    monkeypatch.setenv('MLB_CUDA_GRAPH', '1')
    mgr2 = m.init_training(DEV, cfg, m.SyntheticVectorEnv(N, D, A, seed=11, p_done=0.15, device=DEV).sim_fns(),
                           _policy(m, H, L, RH), None, verbose=False)
    for _ in range(3):
        mgr2.update_iter()
    torch.cuda.synchronize()
    assert mgr2._graph is not None and np.isfinite(mgr2.metrics.latest()['Loss'].mean)
