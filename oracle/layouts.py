"""Oracle (test infrastructure): rollout-buffer index algebra (bit-exact).

Restates the non-PBT branch of /root/reference/src/madrona_learn/rollouts.py:
  store            [C, T', P, B, *leaf]   (:356-367, :460-478)
  training layout  [P, C*B, T', *leaf]    (:788-804, reorder_seq_data / reorder_rnn_data)
  minibatch        take(axis 0) + swapaxes(0,1) -> [T', M, *leaf]   (:319-329)
and the PBT reorder-chunk computation (:1107-1190) used by the reference's only runnable
known-answer test (tests/test_rollouts.py:36-81).
"""
import numpy as np


def sim_to_train(x, P):
    """[N, ...] -> [P, N/P, ...]  (ml/rollouts.py:579-587, non-PBT reshape)."""
    return x.reshape(P, -1, *x.shape[1:])


def store_step(store, c, s, value):
    """store.at[(c, s)].set(value)  (ml/rollouts.py:356-367)."""
    store[c, s] = value
    return store


def reorder_seq_data(x):
    """[C, T', P, B, ...] -> [P, C*B, T', ...]; trajectory j = c*B + b  (:791-793)."""
    t = np.transpose(x, (2, 0, 3, 1, *range(4, x.ndim)))
    return t.reshape(t.shape[0], -1, *t.shape[3:])


def reorder_rnn_data(x):
    """[C, P, B, ...] -> [P, C*B, ...]  (:800-802)."""
    t = np.transpose(x, (1, 0, 2, *range(3, x.ndim)))
    return t.reshape(t.shape[0], -1, *t.shape[3:])


def minibatch(data, indices):
    """RolloutData.minibatch for ONE policy (inside the per-policy vmap, ml/ppo.py:463-469).

    data: dict name -> array [J, T', ...] (and 'rnn_start_states' -> [J, ...]);
    returns dict name -> [T', M, ...], rnn_start_states -> [M, ...]  (:319-329).
    """
    out = {}
    for k, v in data.items():
        if k == 'rnn_start_states':
            out[k] = _tree_map(lambda a: np.take(a, indices, axis=0), v)
        else:
            out[k] = _tree_map(lambda a: np.swapaxes(np.take(a, indices, axis=0), 0, 1), v)
    return out


def _tree_map(fn, t):
    if isinstance(t, dict):
        return {k: _tree_map(fn, v) for k, v in t.items()}
    if isinstance(t, (list, tuple)):
        return type(t)(_tree_map(fn, v) for v in t)
    return fn(t)


def minibatch_from_store(store_leaf, indices, p=0):
    """The fused path: gather [T', M, ...] straight from the [C, T', P, B, ...] store.

    With j = c*B + b:  mb[s, m] = store[j_m // B, s, p, j_m % B]  (SURVEY Appendix A).
    Must equal minibatch(reorder_seq_data(store)[p], indices) bit-exactly.
    """
    C, Tp, P, B = store_leaf.shape[:4]
    idx = np.asarray(indices)
    c = idx // B
    b = idx % B
    # result [T', M, ...]
    return np.stack([store_leaf[c, s, p, b] for s in range(Tp)], axis=0)


def compute_reorder_chunks(assignments, P, C, B):
    """PBT policy-batch reorder.  ml/rollouts.py:1107-1190.

    assignments int32 [S] with values in [0, P); C = policy chunk size; B = #chunks.
    Returns (to_policy_idxs [B, C], to_sim_idxs [S]).  Agents are stably sorted by policy;
    each policy's run is cut into full chunks of C (packed first, in policy order) and one
    partial chunk placed at partial_base + p*C.
    """
    a = np.asarray(assignments, np.int32)
    S = a.size
    sort_idxs = np.argsort(a, kind='stable').astype(np.int32)
    sorted_a = a[sort_idxs]
    counts = np.bincount(sorted_a, minlength=P).astype(np.int32)[:P]
    starts = np.full(P, S, np.int32)
    present = counts > 0
    first = np.searchsorted(sorted_a, np.arange(P), side='left').astype(np.int32)
    starts[present] = first[present]
    num_full = counts // C
    full_counts = (num_full * C).astype(np.int32)
    full_cumsum = np.cumsum(full_counts)
    partial_base = full_cumsum[-1]
    full_starts = full_cumsum - full_counts
    offs = np.arange(S, dtype=np.int32) - starts[sorted_a]
    full_idx = full_starts[sorted_a] + offs
    partial_starts = partial_base + np.arange(0, P * C, C) - full_counts
    partial_idx = partial_starts[sorted_a] + offs
    pos = np.where(offs < full_counts[sorted_a], full_idx, partial_idx).astype(np.int32)
    to_policy = np.full(B * C, S, np.int32)
    to_policy[pos] = sort_idxs
    to_policy = to_policy.reshape(B, C)
    to_policy = np.where(to_policy != S, to_policy, to_policy[:, 0:1])
    to_sim = np.empty(S, np.int32)
    to_sim[sort_idxs] = pos
    return to_policy, to_sim


def dp_assign_minibatch(ids, world, B, M):
    """Owner-affine split of one GLOBAL minibatch `ids` (world*M global trajectory ids j = c*(world*B) + r*B + b,
    i.e. a slice of the reference's permutation, ml/ppo.py:445-463 on the concatenated rollout) into `world`
    lists of M ids (the data-parallel extension; not in the single-device reference): every rank
    keeps the ids it owns, in order, at most M; the surplus of the over-represented ranks (rank order, then
    list order) fills the deficits of the others (rank order).  The union is `ids` exactly."""
    import numpy as np
    ids = np.asarray(ids)
    owner = (ids % (world * B)) // B
    own = [ids[owner == r] for r in range(world)]
    keep = [o[:M] for o in own]
    surplus = np.concatenate([o[M:] for o in own]) if world else ids[:0]
    out, q = [], 0
    for r in range(world):
        need = M - len(keep[r])
        out.append(np.concatenate([keep[r], surplus[q:q + need]]).astype(np.int32))
        q += need
    assert q == len(surplus)
    return out
