"""Oracle (test infrastructure): GAE / discounted returns / z-score.

Restates /root/reference/src/madrona_learn/algo_common.py:45-140 in NumPy float32, keeping
the reference's structure: a reverse loop over T, vectorised over the N columns, every
intermediate rounded to float32 exactly where XLA would round it (no FMA contraction).
"""
import numpy as np

F32 = np.float32


def _as_tn(x, T, N):
    return np.ascontiguousarray(x).reshape(T, N)


def compute_advantages(gamma, gae_lambda, rewards, values, dones, bootstrap_values):
    """GAE(lambda).  ml/algo_common.py:84-130.

    rewards, values: float32 [C, T', P, B, 1] (or anything reshapeable to [T, N]);
    dones: bool/uint8 same shape; bootstrap_values: float32 [P, B, 1] -> [N].
    Returns advantages with the shape of ``rewards``.

    Carry is (next_advantage=0, next_values=bootstrap) (:126-128); at step i (reverse):
      nv = where(done_i, 0, nv); na = where(done_i, 0, na)              (:112-113)
      td = r_i + gamma*nv - v_i                                          (:116)
      A_i = td + (gamma*lambda)*na                                       (:120)
      carry <- (A_i, v_i)                                                (:124)
    gamma and gamma*lambda are Python floats (ml/cfg.py:78,83): the product gamma*lambda is
    formed in double and only then rounded to f32 by weak-type promotion.
    """
    shape = rewards.shape
    dones_arr = np.asarray(dones)
    T = int(np.prod(shape[:2])) if len(shape) >= 4 else shape[0]
    N = int(np.prod(shape)) // T
    r = _as_tn(np.asarray(rewards, F32), T, N)
    v = _as_tn(np.asarray(values, F32), T, N)
    d = _as_tn(dones_arr, T, N).astype(bool)
    g = F32(gamma)
    gl = F32(gamma * gae_lambda)
    nv = np.asarray(bootstrap_values, F32).reshape(N).copy()
    na = np.zeros(N, F32)
    adv = np.empty((T, N), F32)
    zero = F32(0)
    for i in range(T - 1, -1, -1):
        nv = np.where(d[i], zero, nv)
        na = np.where(d[i], zero, na)
        td = (r[i] + g * nv) - v[i]
        a = td + gl * na
        adv[i] = a
        na = a
        nv = v[i]
    return adv.reshape(shape)


def compute_returns(gamma, rewards, dones, bootstrap_values):
    """Discounted returns.  ml/algo_common.py:45-81.

    nr = where(done_i, 0, nr); R_i = r_i + gamma*nr (:70-72); nr initialised to bootstrap.
    """
    shape = rewards.shape
    T = int(np.prod(shape[:2])) if len(shape) >= 4 else shape[0]
    N = int(np.prod(shape)) // T
    r = _as_tn(np.asarray(rewards, F32), T, N)
    d = _as_tn(np.asarray(dones), T, N).astype(bool)
    g = F32(gamma)
    nr = np.asarray(bootstrap_values, F32).reshape(N).copy()
    ret = np.empty((T, N), F32)
    zero = F32(0)
    for i in range(T - 1, -1, -1):
        nr = np.where(d[i], zero, nr)
        cur = r[i] + g * nr
        ret[i] = cur
        nr = cur
    return ret.reshape(shape)


def zscore_data(data):
    """(x - mean) * rsqrt(max(var, 1e-5)) over ALL elements.  ml/algo_common.py:133-140.

    mean/var use float32 accumulation in the reference (``dtype=jnp.float32``); XLA's
    reduction order is unspecified, so the oracle accumulates in float64 and rounds once --
    the stated tolerance (rel 1e-5) covers the difference.  var is the population variance
    mean((x-mean)^2) (jnp.var), not E[x^2]-E[x]^2.
    """
    x = np.asarray(data, F32)
    mean = F32(np.mean(x, dtype=np.float64))
    var = F32(np.mean(np.square(x.astype(np.float64) - np.float64(mean)), dtype=np.float64))
    rstd = F32(1.0) / np.sqrt(np.maximum(var, F32(1e-5)), dtype=F32)
    return ((x - mean) * F32(rstd)).astype(F32)


def zscore_stats(data):
    """The (mean, rstd) pair zscore_data applies; what the fused loss kernel consumes."""
    x = np.asarray(data, F32)
    mean = F32(np.mean(x, dtype=np.float64))
    var = F32(np.mean(np.square(x.astype(np.float64) - np.float64(mean)), dtype=np.float64))
    rstd = F32(1.0) / np.sqrt(np.maximum(var, F32(1e-5)), dtype=F32)
    return mean, F32(rstd)
