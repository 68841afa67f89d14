"""CPU: the oracle's manual backward (MLP+LN+ReLU, heads, PPO loss, LSTM BPTT) equals
torch.autograd on an independent float64 torch restatement of the same forward."""
import numpy as np
import pytest
import torch

from oracle import nn, ppo


def _torch_forward_loss(tp, mb, cfg, vn=None):
    T, M = mb['obs'].shape[:2]
    rows = T * M
    x = torch.from_numpy(mb['obs'].reshape(rows, -1)).double()
    for lyr in tp['mlp']:
        z = x @ lyr['kernel']
        mean = z.mean(-1, keepdim=True)
        var = torch.clamp((z * z).mean(-1, keepdim=True) - mean * mean, min=0)
        x = torch.relu((z - mean) * torch.rsqrt(var + 1e-6) * lyr['scale'] + lyr['bias'])
    logits = x @ tp['actor']['kernel'] + tp['actor']['bias']
    v = x @ tp['critic']['kernel'] + tp['critic']['bias']
    acts = torch.from_numpy(mb['actions'].reshape(rows, -1)).long()
    old = torch.from_numpy(mb['log_probs'].reshape(rows, -1)).double()
    adv = torch.from_numpy(mb['advantages'].reshape(rows, 1)).double()
    if cfg.normalize_advantages:
        adv = (adv - adv.mean()) * torch.rsqrt(torch.clamp(adv.var(unbiased=False), min=1e-5))
    off, obj, ent = 0, [], []
    for i, nb in enumerate(cfg.buckets):
        lp = torch.log_softmax(logits[:, off:off + nb], -1)
        ent.append(-(lp.exp() * lp).sum(-1))
        new = lp.gather(1, acts[:, i:i + 1])[:, 0]
        ratio = (new - old[:, i]).exp()
        a = adv[:, 0]
        obj.append(torch.minimum(a * ratio, a * ratio.clamp(1 - cfg.clip_coef, 1 + cfg.clip_coef)))
        off += nb
    obj, ent = torch.stack(obj, 1), torch.stack(ent, 1)
    ret = torch.from_numpy(mb['returns'].reshape(rows, 1)).double()
    if vn is not None:
        ret = (ret - vn[0]) * vn[1]
    vv = v
    if cfg.clip_value_loss:
        old_v = torch.from_numpy(mb['values'].reshape(rows, 1)).double()
        vv = torch.maximum(torch.minimum(v, old_v + cfg.clip_coef), old_v - cfg.clip_coef)
    d = vv - ret
    vl = torch.nn.functional.huber_loss(vv, ret, reduction='none') if cfg.huber_value_loss else 0.5 * d * d
    return -obj.mean() + cfg.value_loss_coef * vl.mean() - cfg.entropy_coef * ent.mean()


def _mk_mb(rng, T, M, D, buckets):
    A = len(buckets)
    return dict(
        obs=rng.standard_normal((T, M, D)).astype(np.float32),
        actions=np.stack([rng.integers(0, b, (T, M)) for b in buckets], -1).astype(np.int32),
        log_probs=(-np.abs(rng.standard_normal((T, M, A))) - 0.5).astype(np.float32),
        advantages=rng.standard_normal((T, M, 1)).astype(np.float32),
        returns=rng.standard_normal((T, M, 1)).astype(np.float32),
        values=rng.standard_normal((T, M, 1)).astype(np.float32),
        mb_weights=np.ones((M, 1), np.float32))


@pytest.mark.parametrize('clipv,huber', [(False, False), (True, False), (False, True), (True, True)])
def test_ppo_loss_grads_match_autograd(clipv, huber):
    rng = np.random.default_rng(0)
    buckets = [4, 8, 5, 5, 2, 2]
    cfg = ppo.PPOCfg(buckets, clip_value_loss=clipv, huber_value_loss=huber, entropy_coef=0.02)
    params = nn.init_params(rng, 12, 32, 3, buckets, dtype=np.float64)
    # make the actor non-degenerate so ratios leave the clip range
    params['actor']['kernel'] = rng.standard_normal(params['actor']['kernel'].shape) * 0.5
    for l in params['mlp']:
        l['scale'] = 1 + 0.1 * rng.standard_normal(l['scale'].shape)
        l['bias'] = 0.1 * rng.standard_normal(l['bias'].shape)
    mb = _mk_mb(rng, 5, 16, 12, buckets)
    out = ppo.ppo_loss(params, mb, cfg, dtype=np.float64)
    tp = nn.tree_map(lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True), params)
    loss = _torch_forward_loss(tp, mb, cfg)
    loss.backward()
    np.testing.assert_allclose(out['loss'], loss.item(), rtol=1e-6)
    tg = nn.tree_map(lambda t: t.grad.numpy(), tp)
    nn.tree_map(lambda a, b: np.testing.assert_allclose(a, b, rtol=2e-4, atol=1e-8),
                out['grads'], tg)


def test_lstm_bptt_matches_autograd():
    rng = np.random.default_rng(1)
    T, M, D, H, L = 6, 5, 7, 8, 2
    p = nn.init_params(rng, D, D, 1, [3], dtype=np.float64, lstm_hidden=H, lstm_layers=L)['lstm']
    for l in p:
        l['bh'] = 0.1 * rng.standard_normal(l['bh'].shape)
    xs = rng.standard_normal((T, M, D))
    ends = rng.random((T, M)) < 0.3
    c0 = [rng.standard_normal((M, H)) for _ in range(L)]
    h0 = [rng.standard_normal((M, H)) for _ in range(L)]
    outs, caches = nn.lstm_sequence_fwd(c0, h0, xs, ends, p)
    dout = rng.standard_normal(outs.shape)
    dxs, grads = nn.lstm_sequence_bwd(dout, caches, p, H)

    tp = nn.tree_map(lambda a: torch.tensor(a, requires_grad=True), p)
    tx = torch.tensor(xs, requires_grad=True)
    cs = [torch.tensor(c) for c in c0]
    hs = [torch.tensor(h) for h in h0]
    touts = []
    for t in range(T):
        x = tx[t]
        ncs, nhs, o = [], [], []
        for l in range(L):
            z = x @ tp[l]['wi'] + hs[l] @ tp[l]['wh'] + tp[l]['bh']
            i, f, g, oo = torch.sigmoid(z[:, :H]), torch.sigmoid(z[:, H:2*H]), torch.tanh(z[:, 2*H:3*H]), torch.sigmoid(z[:, 3*H:])
            c = f * cs[l] + i * g
            h = oo * torch.tanh(c)
            x = h
            ncs.append(c), nhs.append(h), o.append(h)
        keep = torch.tensor(~ends[t]).double().reshape(-1, 1)
        cs = [c * keep for c in ncs]
        hs = [h * keep for h in nhs]
        touts.append(torch.cat(o, -1))
    tout = torch.stack(touts)
    np.testing.assert_allclose(outs, tout.detach().numpy(), rtol=1e-10)
    (tout * torch.tensor(dout)).sum().backward()
    np.testing.assert_allclose(dxs, tx.grad.numpy(), rtol=1e-8, atol=1e-12)
    for l in range(L):
        for k in ('wi', 'wh', 'bh'):
            np.testing.assert_allclose(grads[l][k], tp[l][k].grad.numpy(), rtol=1e-8, atol=1e-12)


def test_optimizer_step_properties():
    rng = np.random.default_rng(2)
    buckets = [3, 2]
    cfg = ppo.PPOCfg(buckets)
    params = nn.init_params(rng, 6, 16, 2, buckets)
    norms = ppo.initial_weight_norms(params)
    grads = nn.tree_map(lambda a: rng.standard_normal(a.shape).astype(np.float32), params)
    opt = ppo.adam_init(params)
    new_p, opt2, gn = ppo.optimizer_step(params, grads, opt, cfg, norms)
    assert opt2['t'] == 1 and gn > cfg.max_grad_norm
    for i, l in enumerate(new_p['mlp']):
        np.testing.assert_allclose(np.linalg.norm(l['kernel']), norms['mlp'][i], rtol=1e-5)
        np.testing.assert_allclose(np.dot(l['scale'], l['scale']) + np.dot(l['bias'], l['bias']),
                                   l['scale'].size, rtol=1e-5)
    # first adam step moves every (unclipped-direction) weight by ~lr
    d = new_p['actor']['kernel'] - params['actor']['kernel']
    np.testing.assert_allclose(np.abs(d), cfg.lr, rtol=1e-3)


def test_separate_backbone_oracle_gradient_matches_finite_differences():
    """BackboneSeparate restatement (ml/actor_critic.py:247-303): analytic gradient of the PPO loss through the
    two towers vs central differences in fp64; the actor tower gets no gradient from the value loss."""
    from oracle import nn as onn, ppo as oppo
    rng = np.random.default_rng(0)
    B = [3, 4]
    D, H, L, Tp, M = 8, 16, 2, 2, 8
    p = onn.init_params(rng, D, H, L, B, separate=True, dtype=np.float64)
    p['actor']['kernel'] = rng.standard_normal(p['actor']['kernel'].shape) * 0.3
    cfg = oppo.PPOCfg(B, entropy_coef=0.02)
    mb = dict(obs=rng.standard_normal((Tp, M, D)),
              actions=np.stack([rng.integers(0, b, (Tp, M)) for b in B], -1).astype(np.int32),
              advantages=rng.standard_normal((Tp, M, 1)), returns=rng.standard_normal((Tp, M, 1)),
              values=rng.standard_normal((Tp, M, 1)), mb_weights=np.ones((M, 1)),
              log_probs=-np.abs(rng.standard_normal((Tp, M, 2))) - 0.5)
    ref = oppo.ppo_loss(p, mb, cfg, None, dtype=np.float64)
    for path in (('mlp', 0, 'kernel', (1, 2)), ('mlp_critic', 1, 'kernel', (3, 4)), ('mlp_critic', 0, 'scale', (2,)),
                 ('critic', 'kernel', (5, 0)), ('actor', 'kernel', (2, 3))):
        def get(t):
            for k in path[:-1]:
                t = t[k]
            return t
        a, idx, e = get(p), path[-1], 1e-6
        old = a[idx]
        a[idx] = old + e
        lp = oppo.ppo_loss(p, mb, cfg, None, dtype=np.float64, want_grads=False)['loss']
        a[idx] = old - e
        lm = oppo.ppo_loss(p, mb, cfg, None, dtype=np.float64, want_grads=False)['loss']
        a[idx] = old
        np.testing.assert_allclose((lp - lm) / (2 * e), get(ref['grads'])[idx], rtol=1e-5, atol=1e-9)
    # value-loss-only: the actor tower is untouched
    cfg0 = oppo.PPOCfg(B, entropy_coef=0.0, clip_coef=0.2)
    mb0 = dict(mb, advantages=np.zeros((Tp, M, 1)))
    g = oppo.ppo_loss(p, mb0, cfg0, None, dtype=np.float64, adv_stats=(0.0, 1.0))['grads']
    assert all(np.abs(l['kernel']).max() == 0 for l in g['mlp'])
    assert any(np.abs(l['kernel']).max() > 0 for l in g['mlp_critic'])
