"""Oracle (test infrastructure): PPO minibatch loss / gradients / optimiser / update loop.

Restates /root/reference/src/madrona_learn/ppo.py:109-488 (default branch: no
filter_advantages, no importance sampling, plain critic = DenseLayerCritic, P = 1) on top
of oracle/nn.py.  optax pieces (PARITY UNPINNED, optax 0.1.9 is not under /root/reference):
clip_by_global_norm, adam(b1 .9, b2 .999, eps 1e-8), l2_loss = 0.5 d^2, huber_loss(delta 1).
"""
import numpy as np

from . import algo_common, layouts, nn, prng
from .moving_avg import EMANormalizer


class PPOCfg:
    """The hyper-parameters the loss/update read (ml/ppo.py:24-46, ml/cfg.py:68-96)."""

    def __init__(self, buckets, num_epochs=4, minibatch_size=2048, clip_coef=0.2,
                 value_loss_coef=0.5, entropy_coef=0.01, max_grad_norm=0.5, lr=3e-4,
                 clip_value_loss=False, huber_value_loss=False, normalize_advantages=True,
                 normalize_values=False, value_normalizer_decay=0.99999, gamma=0.99,
                 gae_lambda=0.95, partitionable=False, dreamer_v3_critic=False, compute_advantages=True,
                 normalize_returns=True, hlgauss=None, continuous=None):
        """hlgauss: (centers, bounds, smoothness) of an HLGaussCritic (cfg.hlgauss_critic), else None.
        continuous: (stddev_min, stddev_max) of a ContinuousActionsConfig (then `buckets` = [None] * num_dims
        and the actor columns are raw means | raw stds), else None."""
        self.__dict__.update(locals())
        del self.__dict__['self']


def ppo_loss(params, mb, cfg, vn_state=None, dtype=np.float64, want_grads=True, adv_stats=None,
             new_vn_state=None, quant=None):
    """_ppo_update's loss_fn (ml/ppo.py:129-262) and its gradient w.r.t. params.

    mb: dict with obs [T', M, D], actions [T', M, A] i32, log_probs [T', M, A],
    advantages/returns/values [T', M, 1], mb_weights [M, 1] (ones on the default branch).
    Returns dict(loss, action_obj, value_loss, entropy, grads, new_vn_state, aux...).
    """
    f = dtype
    Tp, M = mb['obs'].shape[:2]
    rows = Tp * M
    p = nn.cast_tree(params, f)
    obs = mb['obs'].reshape(rows, -1).astype(f)
    seq = None
    if 'lstm' in p:
        rs = mb['rnn_start_states']
        seq = dict(Tp=Tp, M=M, ends=np.asarray(mb['dones']).reshape(Tp, M).astype(bool), c0=rs[0], h0=rs[1])
    logits, critic, cache = nn.actor_critic_fwd(p, obs, quant, seq)
    out = ppo_loss_heads(logits, critic, mb, cfg, vn_state, f, want_grads, adv_stats, new_vn_state)
    if want_grads:
        out['grads'] = nn.actor_critic_bwd(p, cache, out['dlogits'], out['dcritic'], quant)
    return out


def _action_bwd(logits, acts, cfg, dlogp, dent, A, f):
    cont = getattr(cfg, 'continuous', None)
    if cont is None:
        return nn.action_stats_bwd(logits, acts, cfg.buckets, dlogp, dent)
    from . import dists as _d
    dm, ds = _d.continuous_action_stats_bwd(logits[:, :A], logits[:, A:2 * A], acts, cont[0], cont[1], dlogp, dent, f)
    return np.concatenate([dm, ds], axis=1)


def ppo_loss_heads(logits, critic, mb, cfg, vn_state=None, dtype=np.float64, want_grads=True,
                   adv_stats=None, new_vn_state=None):
    """The part of loss_fn downstream of the policy forward (ml/ppo.py:131-262): everything the
    fused loss kernel computes from the head outputs `logits` [rows, sumA] / `critic` [rows, V],
    plus d loss / d(logits, critic).  Pinned against the reference's own `_ppo_update` run under
    oracle/jax_shim (tests/golden/ppo_loss.npz)."""
    f = dtype
    Tp, M = mb['actions'].shape[:2]
    rows = Tp * M
    A = len(cfg.buckets)
    logits, critic = np.asarray(logits, f), np.asarray(critic, f)
    acts = mb['actions'].reshape(rows, A)
    old_lp = mb['log_probs'].reshape(rows, A).astype(f)
    w = np.broadcast_to(mb['mb_weights'].reshape(1, M, 1), (Tp, M, 1)).reshape(rows, 1).astype(f)
    cont = getattr(cfg, 'continuous', None)
    if cont is not None:          # ContinuousActionDistributions.action_stats (ml/dists.py:260-284)
        from . import dists as _d
        acts = np.asarray(acts).view(np.float32) if acts.dtype == np.int32 else acts.astype(np.float32)
        new_lp, ent = _d.continuous_action_stats(logits[:, :A], logits[:, A:2 * A], acts, cont[0], cont[1], dtype=f)
    else:
        new_lp, ent = nn.action_stats(logits, acts, cfg.buckets)

    # advantages: per-MINIBATCH z-score (ml/ppo.py:134-137 -> ml/algo_common.py:133-140)
    if getattr(cfg, 'compute_advantages', True):
        adv = mb['advantages'].reshape(rows, 1).astype(np.float32)
        do_norm = cfg.normalize_advantages
    else:       # :138-143: the returns play the role of the scores
        adv = mb['returns'].reshape(rows, 1).astype(np.float32)
        do_norm = getattr(cfg, 'normalize_returns', True)
    if do_norm:
        if adv_stats is None:
            adv = algo_common.zscore_data(adv)
        else:       # data-parallel: (mean, rstd) of the GLOBAL minibatch, computed elsewhere
            adv = ((adv - np.float32(adv_stats[0])) * np.float32(adv_stats[1])).astype(np.float32)
    adv = adv.astype(f)

    ratio = np.exp(new_lp - old_lp)                                       # :146-147
    surr1 = adv * ratio                                                   # :155
    clipped = np.clip(ratio, 1.0 - cfg.clip_coef, 1.0 + cfg.clip_coef)    # :157-159
    surr2 = adv * clipped
    obj = np.minimum(surr1, surr2)                                        # :162

    returns = mb['returns'].reshape(rows, 1).astype(np.float32)
    if cfg.dreamer_v3_critic or getattr(cfg, 'hlgauss', None) is not None:
        from . import dists
        clog = critic.reshape(rows, -1)
        V = clog.shape[1]
        if cfg.dreamer_v3_critic:
            # two-hot cross-entropy on the critic's bin logits (:169-177, ml/dists.py:172-208)
            b = dists.bins(V)
            lo = np.clip((b <= returns).astype(np.int32).sum(-1) - 1, 0, V - 1)
            hi = np.clip(V - (b > returns).astype(np.int32).sum(-1), 0, V - 1)
            same = lo == hi
            dl = np.where(same, 1, np.abs(b[lo] - returns[:, 0]))
            du = np.where(same, 1, np.abs(b[hi] - returns[:, 0]))
            wl, wu = dl / (dl + du), du / (dl + du)
            two_hot = np.zeros((rows, V), f)
            np.add.at(two_hot, (np.arange(rows), lo), wl)
            np.add.at(two_hot, (np.arange(rows), hi), wu)
            vmean = dists.twohot_mean(clog.astype(np.float32))
        else:
            # HL-Gauss: cross-entropy against the Gaussian histogram (:178-185, ml/models.py:212-250);
            # `two_hot` is the target distribution c (it sums to 1, so d loss / d logits = softmax - c)
            centers, bounds, smooth = cfg.hlgauss
            two_hot = dists.hlgauss_target(returns, centers, bounds, smooth, dtype=f)
            vmean = dists.hlgauss_mean(clog.astype(np.float32), centers)
        m_ = clog.max(-1, keepdims=True)
        logp = clog - (np.log(np.exp(clog - m_).sum(-1, keepdims=True)) + m_)
        vloss = -(two_hot * logp).sum(-1, keepdims=True)
        value_errs = vmean.astype(f) - returns.astype(f)
        action_obj_avg = np.mean(w * obj)
        value_loss = cfg.value_loss_coef * np.mean(w * vloss)
        entropy_avg = cfg.entropy_coef * np.mean(w * ent)
        loss = -action_obj_avg + value_loss - entropy_avg
        out = dict(loss=f(loss), action_obj=obj, value_losses=vloss, entropies=ent, value_errs=value_errs,
                   new_vn_state=None, logits=logits, critic=critic, new_log_probs=new_lp)
        if not want_grads:
            return out
        pick1 = (surr1 <= surr2)
        inside = (ratio >= 1.0 - cfg.clip_coef) & (ratio <= 1.0 + cfg.clip_coef)
        dobj_dratio = np.where(pick1 | inside, adv, 0.0)
        dlogp = -(w * dobj_dratio * ratio) / (rows * A)
        dent = -(cfg.entropy_coef * w) / (rows * A) * np.ones_like(ent)
        dlogits = _action_bwd(logits, acts, cfg, dlogp, dent, A, f)
        dcritic = cfg.value_loss_coef * w * (np.exp(logp) - two_hot) / rows
        out['dlogits'], out['dcritic'] = dlogits, dcritic
        return out
    # value loss, plain critic branch (:186-218)
    v_new = critic.reshape(rows, 1)
    norm = EMANormalizer(cfg.value_normalizer_decay) if cfg.normalize_values else None
    if norm is None:
        value_errs = v_new - returns.astype(f)
        norm_returns = returns.astype(f)
        new_vn = None
    else:
        value_errs = (v_new * f(vn_state['sigma'][0]) + f(vn_state['mu'][0])) - returns.astype(f)
        if new_vn_state is None:
            new_vn, nr = norm.normalize_and_update_estimates(vn_state, returns)
        else:       # data-parallel: the EMA was advanced with the GLOBAL minibatch statistics
            new_vn, nr = new_vn_state, norm.normalize(new_vn_state, returns)
        norm_returns = nr.astype(f)
    v_used = v_new
    vclip_mask = np.ones_like(v_new)
    if cfg.clip_value_loss:
        old_v = mb['values'].reshape(rows, 1).astype(f)
        lo, hi = old_v - cfg.clip_coef, old_v + cfg.clip_coef
        v_used = np.clip(v_new, lo, hi)
        vclip_mask = ((v_new >= lo) & (v_new <= hi)).astype(f)
    d = v_used - norm_returns
    if cfg.huber_value_loss:
        q = np.minimum(np.abs(d), 1.0)
        vloss = 0.5 * q * q + (np.abs(d) - q)
        dvloss = np.clip(d, -1.0, 1.0)
    else:
        vloss = 0.5 * d * d
        dvloss = d

    action_obj_avg = np.mean(w * obj)                                     # :220-224 (one group)
    value_loss = cfg.value_loss_coef * np.mean(w * vloss)                 # :227-228, :247
    entropy_avg = cfg.entropy_coef * np.mean(w * ent)                     # :231-239
    loss = -action_obj_avg + value_loss - entropy_avg                     # :245-252

    out = dict(loss=f(loss), action_obj=obj, value_losses=vloss, entropies=ent,
               value_errs=value_errs, new_vn_state=new_vn, logits=logits, critic=critic,
               new_log_probs=new_lp)
    if not want_grads:
        return out

    # d loss / d new_lp: min() picks surr1 unless the clipped branch is strictly smaller
    pick1 = (surr1 <= surr2)
    inside = (ratio >= 1.0 - cfg.clip_coef) & (ratio <= 1.0 + cfg.clip_coef)
    dobj_dratio = np.where(pick1 | inside, adv, 0.0)
    dlogp = -(w * dobj_dratio * ratio) / (rows * A)
    dent = -(cfg.entropy_coef * w) / (rows * A) * np.ones_like(ent)
    dlogits = _action_bwd(logits, acts, cfg, dlogp, dent, A, f)
    dcritic = cfg.value_loss_coef * w * dvloss * vclip_mask / rows
    out['dlogits'] = dlogits
    out['dcritic'] = dcritic
    return out


# ---------------------------------------------------------------------------------------
# optimiser (ml/ppo.py:84-90, 283-338)
# ---------------------------------------------------------------------------------------
def adam_init(params):
    z = lambda x: np.zeros_like(x)
    return dict(m=nn.tree_map(z, params), v=nn.tree_map(z, params), t=0)


def global_norm(grads):
    return np.sqrt(sum(float(np.sum(np.square(g.astype(np.float64)))) for g in nn.tree_leaves(grads)))


def optimizer_step(params, grads, opt, cfg, init_norms, dtype=np.float32):
    """clip_by_global_norm -> adam -> apply_updates -> kernel re-projection -> LN renorm."""
    f = dtype
    gn = global_norm(grads)
    scale = 1.0 if gn < cfg.max_grad_norm else cfg.max_grad_norm / gn
    t = opt['t'] + 1
    b1, b2, eps = 0.9, 0.999, 1e-8
    bc1, bc2 = 1.0 - b1 ** t, 1.0 - b2 ** t

    def upd(p, g, m, v):
        g = g.astype(f) * f(scale)
        m2 = f(b1) * m + f(1 - b1) * g
        v2 = f(b2) * v + f(1 - b2) * g * g
        u = -f(cfg.lr) * (m2 / f(bc1)) / (np.sqrt(v2 / f(bc2)) + f(eps))
        return (p + u).astype(f), m2.astype(f), v2.astype(f)

    trip = nn.tree_map(upd, params, grads, opt['m'], opt['v'])

    def pick(i):
        def rec(t3):
            if isinstance(t3, dict):
                return {k: rec(v) for k, v in t3.items()}
            if isinstance(t3, list):
                return [rec(v) for v in t3]
            return t3[i]
        return rec(trip)
    new_p, new_m, new_v = pick(0), pick(1), pick(2)

    # kernel re-projection to the initial L2 norm for every backbone 'kernel' (:303-310);
    # top-level actor / critic are excluded (ml/train_state.py:422-423)
    for name in ('mlp', 'mlp_critic'):        # 'mlp_critic': the critic encoder of a BackboneSeparate
        for i, lyr in enumerate(new_p.get(name, [])):
            k = lyr['kernel']
            lyr['kernel'] = (f(init_norms[name][i]) * k / np.sqrt(np.sum(np.square(k), dtype=np.float64))).astype(f)
            # LayerNorm renorm (:312-338): factor = sqrt(F / (b.b + s.s))
            s, b = lyr['scale'], lyr['bias']
            fac = np.sqrt(s.shape[-1] / (np.dot(b.astype(np.float64), b) + np.dot(s.astype(np.float64), s)))
            lyr['scale'] = (fac * s).astype(f)
            lyr['bias'] = (fac * b).astype(f)
    # flax keeps one kernel leaf PER GATE (ii, if, ig, io, hi, hf, hg, ho): each column block of the
    # stacked [in, 4H] matrices is re-projected to its own initial norm
    for i, lyr in enumerate(new_p.get('lstm', [])):
        for kk in ('wi', 'wh'):
            k = lyr[kk].copy()
            Hh = k.shape[1] // 4
            for g in range(4):
                blk = k[:, g * Hh:(g + 1) * Hh]
                k[:, g * Hh:(g + 1) * Hh] = f(init_norms['lstm'][i][kk][g]) * blk / np.sqrt(
                    np.sum(np.square(blk), dtype=np.float64))
            lyr[kk] = k.astype(f)
    return new_p, dict(m=new_m, v=new_v, t=t), gn


def initial_weight_norms(params):
    """ml/train_state.py:413-423."""
    n = {'mlp': [float(np.sqrt(np.sum(np.square(l['kernel'].astype(np.float64))))) for l in params['mlp']]}
    if 'mlp_critic' in params:
        n['mlp_critic'] = [float(np.sqrt(np.sum(np.square(l['kernel'].astype(np.float64))))) for l in params['mlp_critic']]
    if 'lstm' in params:
        def gate_norms(k):
            Hh = k.shape[1] // 4
            return [float(np.sqrt(np.sum(np.square(k[:, g * Hh:(g + 1) * Hh].astype(np.float64))))) for g in range(4)]
        n['lstm'] = [{k: gate_norms(l[k]) for k in ('wi', 'wh')} for l in params['lstm']]
    return n


# ---------------------------------------------------------------------------------------
# the update loop (ml/ppo.py:366-488, default branch)
# ---------------------------------------------------------------------------------------
def flatten_time(rollout):
    """RolloutData.flatten_time (ml/rollouts.py:331-334): [J, T', ...] -> [J*T', 1, ...]."""
    return {k: v.reshape(-1, 1, *v.shape[2:]) for k, v in rollout.items() if k != 'rnn_start_states'}


def select_filter_advantages(advantages, est_state, decay, M):
    """filter_advantages branch of _ppo (ml/ppo.py:374-405).  advantages [J, T', 1] (training layout).
    -> (valid_inds int32 [J*T'], num_minibatches, new max_advantage_est_state)."""
    from .moving_avg import EMAEstimate
    flat = np.abs(np.asarray(advantages, np.float32)).reshape(-1)          # flatten_time order f = j*T' + s
    est = EMAEstimate(decay).update_estimates(est_state, flat.max())
    order = np.argsort(-flat, kind='stable').astype(np.int32)              # argsort(descending=True)
    num_above = int(np.sum(flat >= np.float32(0.01) * est['mu'][0]))
    nmb = min((num_above + (M - 1)) // M, flat.size // M)
    valid = np.where(np.arange(flat.size) < nmb * M, order, -1).astype(np.int32)
    return valid, nmb, est


def select_importance(advantages, values, returns, key, num_minibatches, M, partitionable=False):
    """importance_sample_trajectories branch (ml/ppo.py:407-435): [J, T', 1] inputs ->
    (sampled trajectory ids int32 [num_minibatches*M], traj_weights f32 [J, 1], probs).
    jax.random.choice(replace=False, p) is the Gumbel top-k trick (PARITY UNPINNED: jax)."""
    f = np.float32
    a = np.abs(np.asarray(advantages, f)).mean(axis=1)
    ve = np.abs(np.asarray(values, f) - np.asarray(returns, f)).mean(axis=1)
    scores = (a + ve).astype(f)                                             # [J, 1]
    e = np.exp(scores - scores.max(axis=0, keepdims=True))
    probs = (e / e.sum(axis=0, keepdims=True)).astype(f)
    J = scores.shape[0]
    weights = ((f(1.0) / f(J)) / probs).astype(f)
    g = prng.gumbel_from_bits(prng.random_bits(key, (J,), partitionable)) + np.log(probs.reshape(-1))
    ind = np.argsort(-g, kind='stable')[:num_minibatches * M].astype(np.int32)
    return ind, weights, probs


def ppo_update(params, opt, init_norms, rollout, cfg, update_key, vn_state=None,
               dtype=np.float32, perms=None, quant=None, valid_inds=None, traj_weights=None,
               num_minibatches=None):
    """rollout: dict name -> [J, T', ...] training layout (ml/rollouts.py:788-804), P=1.
    valid_inds / traj_weights / num_minibatches: the alternate selections of ml/ppo.py:374-435
    (None: the default branch :437-443).
    Returns (params, opt, new_key, vn_state, last_minibatch_outputs, perms)."""
    J = rollout['dones'].shape[0]
    M = cfg.minibatch_size
    if valid_inds is None:
        assert J % M == 0                                                 # :439
        valid_inds = np.arange(J, dtype=np.int32)                         # :441
        num_minibatches = J // M
    traj_w = np.ones((J, 1), np.float32) if traj_weights is None else traj_weights   # :443
    key = np.asarray(update_key, np.uint32)
    used = []
    last = None
    for e in range(cfg.num_epochs):
        if perms is None:
            ks = prng.split(key, 2, cfg.partitionable)                    # gen_update_rnd
            rnd, key = ks[0], ks[1]
            inds = prng.permutation(rnd, valid_inds, cfg.partitionable)   # :451
            inds = inds[np.argsort(np.where(inds == -1, 1, 0), kind='stable')]    # :453-458
        else:
            inds = perms[e]
        used.append(inds)
        for i in range(num_minibatches):
            mb_inds = inds[i * M:(i + 1) * M]                             # :464-466
            mb = layouts.minibatch(rollout, mb_inds)
            mb['mb_weights'] = traj_w[mb_inds]
            out = ppo_loss(params, mb, cfg, vn_state, dtype=dtype, quant=quant)
            if out['new_vn_state'] is not None:
                vn_state = out['new_vn_state']
            grads = nn.cast_tree(out['grads'], dtype)
            params, opt, gn = optimizer_step(params, grads, opt, cfg, init_norms, dtype)
            out['grad_norm'] = gn
            last = out
    return params, opt, key, vn_state, last, np.stack(used)
