"""Oracle (test infrastructure): policy/value network forward AND manual backward.

Restates, in NumPy (dtype selectable: float32 for the timed CPU baseline, float64 as the
tight numerical reference), the model pieces on the hot path:
  MLP            ml/models.py:99-119   L x [Dense(no bias) -> LayerNorm -> ReLU]
  LayerNorm      ml/models.py:46-56 -> flax 0.8.1 nn.LayerNorm (PARITY UNPINNED, third party):
                 eps 1e-6, fast variance var = max(0, E[x^2] - E[x]^2), scale+bias
  heads          ml/models.py:122-154  Dense(+bias) -> sum(buckets) logits ; Dense(+bias) -> 1
  distributions  ml/dists.py:12-96     per-component log-softmax / entropy / sampling
  LSTM           ml/rnn.py:10-111 -> flax nn.OptimizedLSTMCell (PARITY UNPINNED)
The manual backward is validated against torch.autograd in tests/test_oracle_nn.py.

Parameter tree (flax naming): a dict
  {'mlp': [{'kernel': [in,H], 'scale': [H], 'bias': [H]}, ...],
   'actor': {'kernel': [F, sumA], 'bias': [sumA]},
   'critic': {'kernel': [F, V], 'bias': [V]},
   'lstm': [{'wi': [in,4H], 'wh': [H,4H], 'bh': [4H]}, ...]   (optional; gate order i,f,g,o)}
"""
import numpy as np

LN_EPS = 1e-6


# ---------------------------------------------------------------------------------------
# MLP
# ---------------------------------------------------------------------------------------
def layernorm_relu_fwd(z, scale, bias):
    mean = z.mean(axis=-1, keepdims=True)
    mean2 = (z * z).mean(axis=-1, keepdims=True)
    var = np.maximum(0, mean2 - mean * mean)
    rstd = 1.0 / np.sqrt(var + z.dtype.type(LN_EPS))
    xhat = (z - mean) * rstd
    u = xhat * scale + bias
    return np.maximum(u, 0), (xhat, rstd, u > 0)


def layernorm_relu_bwd(dy, cache, scale):
    xhat, rstd, mask = cache
    du = dy * mask
    dscale = (du * xhat).sum(axis=0)
    dbias = du.sum(axis=0)
    dxhat = du * scale
    dz = rstd * (dxhat - dxhat.mean(axis=-1, keepdims=True)
                 - xhat * (dxhat * xhat).mean(axis=-1, keepdims=True))
    return dz, dscale, dbias


def bf16_round(x):
    """Round-to-nearest-even to bfloat16 precision (value returned in x's dtype).  Used to
    emulate the quantisation points of the tensor-core path (compute_dtype=bfloat16): weights,
    layer inputs/outputs and the back-propagated dY / dZ are bf16 in HBM; every accumulation,
    LayerNorm statistic and the loss stay fp32."""
    x32 = np.ascontiguousarray(x, np.float32)
    u = x32.view(np.uint32)
    r = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    out = ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)
    return out.astype(np.asarray(x).dtype)


def _id(x):
    return x


def mlp_fwd(x, layers, q=None):
    quant = q is not None
    q = q or _id
    caches = []
    x = q(x)
    for lyr in layers:
        z = x @ q(lyr['kernel'])
        y, c = layernorm_relu_fwd(z, lyr['scale'], lyr['bias'])
        if quant:
            # the tensor-core path stashes xhat in bf16 and re-derives the ReLU mask from it
            xh = q(c[0])
            c = (xh, c[1], (xh * lyr['scale'] + lyr['bias']) > 0)
        caches.append((x, c))
        x = q(y)
    return x, caches


def mlp_bwd(dy, caches, layers, need_dx=False, q=None):
    q = q or _id
    grads = [None] * len(layers)
    for i in range(len(layers) - 1, -1, -1):
        x, c = caches[i]
        dz, ds, db = layernorm_relu_bwd(dy, c, layers[i]['scale'])
        dz = q(dz)
        grads[i] = {'kernel': x.T @ dz, 'scale': ds, 'bias': db}
        if i > 0 or need_dx:
            dy = q(dz @ q(layers[i]['kernel']).T)
    return (dy if need_dx else None), grads


# ---------------------------------------------------------------------------------------
# distributions (ml/dists.py:12-96)
# ---------------------------------------------------------------------------------------
def _lse(l):
    m = l.max(axis=-1, keepdims=True)
    return np.log(np.exp(l - m).sum(axis=-1, keepdims=True)) + m


def action_stats(logits, actions, buckets):
    """-> log_probs [rows, A], entropies [rows, A]  (ml/dists.py:54-77)."""
    lps, ents = [], []
    off = 0
    for i, nb in enumerate(buckets):
        l = logits[:, off:off + nb]
        logp = l - _lse(l)
        p = np.exp(logp)
        ents.append(-(p * logp).sum(axis=-1))
        lps.append(np.take_along_axis(logp, actions[:, i:i + 1].astype(np.int64), axis=-1)[:, 0])
        off += nb
    return np.stack(lps, axis=1), np.stack(ents, axis=1)


def action_stats_bwd(logits, actions, buckets, dlogp, dent):
    """Gradient of sum(dlogp*log_probs + dent*entropies) w.r.t. logits."""
    dl = np.zeros_like(logits)
    off = 0
    rows = np.arange(logits.shape[0])
    for i, nb in enumerate(buckets):
        l = logits[:, off:off + nb]
        logp = l - _lse(l)
        p = np.exp(logp)
        H = -(p * logp).sum(axis=-1, keepdims=True)
        g = -p * dlogp[:, i:i + 1]
        g[rows, actions[:, i]] += dlogp[:, i]
        g += dent[:, i:i + 1] * (-p * (logp + H))
        dl[:, off:off + nb] = g
        off += nb
    return dl


def sample_actions(logits, key, buckets, partitionable=False):
    """DiscreteActionDistributions.sample (ml/dists.py:26-44): keys = split(key, A); per
    component categorical(key_i, logits_i) (Gumbel-max), log_prob = logit[a] - logsumexp."""
    from . import prng
    keys = prng.split(key, len(buckets), partitionable)
    acts, lps = [], []
    off = 0
    for i, nb in enumerate(buckets):
        l = logits[:, off:off + nb].astype(np.float32)
        a = prng.categorical(keys[i], l, partitionable)
        lp = np.take_along_axis(l, a[:, None].astype(np.int64), axis=-1) - _lse(l)
        acts.append(a)
        lps.append(lp[:, 0])
        off += nb
    return np.stack(acts, axis=1).astype(np.int32), np.stack(lps, axis=1).astype(np.float32)


def best_actions(logits, buckets):
    acts = []
    off = 0
    for nb in buckets:
        acts.append(np.argmax(logits[:, off:off + nb], axis=-1))
        off += nb
    return np.stack(acts, axis=1).astype(np.int32)


# ---------------------------------------------------------------------------------------
# LSTM (flax OptimizedLSTMCell semantics; gates i, f, g, o; hidden bias only)
# ---------------------------------------------------------------------------------------
def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_cell_fwd(c, h, x, lyr):
    H = c.shape[-1]
    z = x @ lyr['wi'] + h @ lyr['wh'] + lyr['bh']
    i, f, g, o = _sigmoid(z[:, :H]), _sigmoid(z[:, H:2 * H]), np.tanh(z[:, 2 * H:3 * H]), _sigmoid(z[:, 3 * H:])
    c2 = f * c + i * g
    tc = np.tanh(c2)
    h2 = o * tc
    return c2, h2, (x, c, h, i, f, g, o, tc)


def lstm_cell_bwd(dc2, dh2, cache, lyr):
    x, c, h, i, f, g, o, tc = cache
    do = dh2 * tc
    dc = dc2 + dh2 * o * (1 - tc * tc)
    di, df, dg = dc * g, dc * c, dc * i
    dz = np.concatenate([di * i * (1 - i), df * f * (1 - f), dg * (1 - g * g), do * o * (1 - o)], axis=1)
    grads = {'wi': x.T @ dz, 'wh': h.T @ dz, 'bh': dz.sum(axis=0)}
    return dc * f, dz @ lyr['wh'].T, dz @ lyr['wi'].T, grads


def lstm_step(cs, hs, x, lstm):
    """MultiLayerLSTMCell (ml/rnn.py:10-45): output = concat of every layer's new h."""
    ncs, nhs, outs, caches = [], [], [], []
    for l, lyr in enumerate(lstm):
        c2, h2, cache = lstm_cell_fwd(cs[l], hs[l], x, lyr)
        x = h2
        ncs.append(c2), nhs.append(h2), outs.append(h2), caches.append(cache)
    return ncs, nhs, np.concatenate(outs, axis=-1), caches


def lstm_sequence_fwd(cs, hs, xs, ends, lstm):
    """LSTM.sequence (ml/rnn.py:91-111): per step cell then zero the carry where end[t]."""
    T = xs.shape[0]
    outs, caches = [], []
    for t in range(T):
        cs, hs, out, cache = lstm_step(cs, hs, xs[t], lstm)
        keep = (~ends[t].astype(bool)).astype(xs.dtype).reshape(-1, 1)
        cs = [c * keep for c in cs]
        hs = [h * keep for h in hs]
        outs.append(out)
        caches.append((cache, keep))
    return np.stack(outs, axis=0), caches


def lstm_sequence_bwd(douts, caches, lstm, H):
    """BPTT through lstm_sequence_fwd.  douts [T, M, L*H].  Returns (dxs [T, M, in], grads)."""
    T = douts.shape[0]
    L = len(lstm)
    M = douts.shape[1]
    dcs = [np.zeros((M, H), douts.dtype) for _ in range(L)]
    dhs = [np.zeros((M, H), douts.dtype) for _ in range(L)]
    grads = [{k: np.zeros_like(v) for k, v in lyr.items()} for lyr in lstm]
    dxs = []
    for t in range(T - 1, -1, -1):
        cache, keep = caches[t]
        dcs = [d * keep for d in dcs]
        dhs = [d * keep for d in dhs]
        dx_above = None
        for l in range(L - 1, -1, -1):
            dh = dhs[l] + douts[t][:, l * H:(l + 1) * H]
            if dx_above is not None:
                dh = dh + dx_above
            dc_prev, dh_prev, dx, g = lstm_cell_bwd(dcs[l], dh, cache[l], lstm[l])
            dcs[l], dhs[l], dx_above = dc_prev, dh_prev, dx
            for k in g:
                grads[l][k] += g[k]
        dxs.append(dx_above)
    return np.stack(dxs[::-1], axis=0), grads


# ---------------------------------------------------------------------------------------
# full actor-critic
# ---------------------------------------------------------------------------------------
def heads_fwd(feat, params, q=None):
    q = q or _id
    logits = feat @ q(params['actor']['kernel']) + params['actor']['bias']
    critic = feat @ q(params['critic']['kernel']) + params['critic']['bias']
    return logits, critic


def actor_critic_fwd(params, obs, q=None, seq=None):
    """BackboneShared([Recurrent]BackboneEncoder(MLP[, LSTM])) forward (ml/actor_critic.py).
    q: optional quantiser (bf16_round) applied at the tensor-core path's storage points.
    seq (recurrent, ActorCritic.update -> RecurrentBackboneEncoder.sequence :179-199):
    dict(Tp, M, ends [T', M], c0 [list of M x RH], h0) -- obs rows are in [T', M] order."""
    feat, caches = mlp_fwd(obs, params['mlp'], q)
    if 'mlp_critic' in params:
        # BackboneSeparate (ml/actor_critic.py:247-303): a second encoder on the same observations feeds the
        # critic head; the actor head sees only the actor encoder's features
        qq = q or _id
        cfeat, ccaches = mlp_fwd(obs, params['mlp_critic'], q)
        logits = feat @ qq(params['actor']['kernel']) + params['actor']['bias']
        critic = cfeat @ qq(params['critic']['kernel']) + params['critic']['bias']
        return logits, critic, (feat, caches, cfeat, ccaches, 'separate')
    if 'lstm' in params:
        Tp, M = seq['Tp'], seq['M']
        xs = feat.reshape(Tp, M, -1)
        outs, lcaches = lstm_sequence_fwd([np.asarray(c, feat.dtype) for c in seq['c0']],
                                          [np.asarray(h, feat.dtype) for h in seq['h0']],
                                          xs, seq['ends'], params['lstm'])
        rfeat = outs.reshape(Tp * M, -1)
        logits, critic = heads_fwd(rfeat, params, q)
        return logits, critic, (rfeat, caches, lcaches, (Tp, M))
    logits, critic = heads_fwd(feat, params, q)
    return logits, critic, (feat, caches)


def actor_critic_bwd(params, cache, dlogits, dcritic, q=None):
    if len(cache) == 5:                       # BackboneSeparate: two independent towers
        feat, caches, cfeat, ccaches, _ = cache
        qq = q or _id
        dl, dc = qq(dlogits), qq(dcritic)
        g = {'actor': {'kernel': feat.T @ dl, 'bias': dlogits.sum(axis=0)},
             'critic': {'kernel': cfeat.T @ dc, 'bias': dcritic.sum(axis=0)}}
        _, g['mlp'] = mlp_bwd(qq(dl @ qq(params['actor']['kernel']).T), caches, params['mlp'], q=q)
        _, g['mlp_critic'] = mlp_bwd(qq(dc @ qq(params['critic']['kernel']).T), ccaches, params['mlp_critic'], q=q)
        return g
    if len(cache) == 4:                       # recurrent
        rfeat, caches, lcaches, (Tp, M) = cache
        g = {'actor': {'kernel': rfeat.T @ dlogits, 'bias': dlogits.sum(axis=0)},
             'critic': {'kernel': rfeat.T @ dcritic, 'bias': dcritic.sum(axis=0)}}
        dr = dlogits @ params['actor']['kernel'].T + dcritic @ params['critic']['kernel'].T
        H = params['lstm'][0]['wh'].shape[0]
        dxs, g['lstm'] = lstm_sequence_bwd(dr.reshape(Tp, M, -1), lcaches, params['lstm'], H)
        _, g['mlp'] = mlp_bwd(dxs.reshape(Tp * M, -1), caches, params['mlp'])
        return g
    feat, caches = cache
    qq = q or _id
    dl, dc = qq(dlogits), qq(dcritic)
    g = {'actor': {'kernel': feat.T @ dl, 'bias': dlogits.sum(axis=0)},
         'critic': {'kernel': feat.T @ dc, 'bias': dcritic.sum(axis=0)}}
    dfeat = qq(dl @ qq(params['actor']['kernel']).T + dc @ qq(params['critic']['kernel']).T)
    _, g['mlp'] = mlp_bwd(dfeat, caches, params['mlp'], q=q)
    return g


def init_params(rng, obs_dim, hidden, num_layers, buckets, critic_dim=1, dtype=np.float32,
                lstm_hidden=0, lstm_layers=0, separate=False):
    """Orthogonal-ish init (QR of a Gaussian, sign-fixed) with the reference's scales
    (ml/models.py:103,125,144).  Init parity is NOT required (tests load identical weights)."""
    def orth(shape, scale):
        a = rng.standard_normal((max(shape), min(shape)))
        q, r = np.linalg.qr(a)
        q = q * np.sign(np.diag(r))
        if shape[0] < shape[1]:
            q = q.T
        return (scale * q[:shape[0], :shape[1]]).astype(dtype)
    layers = []
    d = obs_dim
    for _ in range(num_layers):
        layers.append({'kernel': orth((d, hidden), np.sqrt(2)), 'scale': np.ones(hidden, dtype),
                       'bias': np.zeros(hidden, dtype)})
        d = hidden
    p = {'mlp': layers}
    if separate:                              # BackboneSeparate: the critic encoder's own stack
        p['mlp_critic'] = []
        d = obs_dim
        for _ in range(num_layers):
            p['mlp_critic'].append({'kernel': orth((d, hidden), np.sqrt(2)), 'scale': np.ones(hidden, dtype),
                                    'bias': np.zeros(hidden, dtype)})
            d = hidden
    feat = hidden
    if lstm_layers:
        lst = []
        din = hidden
        for _ in range(lstm_layers):
            lst.append({'wi': np.concatenate([orth((din, lstm_hidden), 1.0) for _ in range(4)], axis=1),
                        'wh': np.concatenate([orth((lstm_hidden, lstm_hidden), 1.0) for _ in range(4)], axis=1),
                        'bh': np.zeros(4 * lstm_hidden, dtype)})
            din = lstm_hidden
        p['lstm'] = lst
        feat = lstm_hidden * lstm_layers
    nA = int(sum(buckets))
    p['actor'] = {'kernel': orth((feat, nA), 0.01), 'bias': np.zeros(nA, dtype)}
    p['critic'] = {'kernel': orth((feat, critic_dim), 1.0), 'bias': np.zeros(critic_dim, dtype)}
    return p


def tree_map(fn, *trees):
    t0 = trees[0]
    if isinstance(t0, dict):
        return {k: tree_map(fn, *[t[k] for t in trees]) for k in t0}
    if isinstance(t0, (list, tuple)):
        return [tree_map(fn, *[t[i] for t in trees]) for i in range(len(t0))]
    return fn(*trees)


def tree_leaves(t):
    out = []
    tree_map(lambda x: out.append(x), t)
    return out


def cast_tree(t, dtype):
    return tree_map(lambda x: np.asarray(x, dtype), t)
