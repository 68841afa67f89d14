"""Oracle (test infrastructure): SymExpTwoHotDistribution (the reference's DEFAULT critic,
ml/cfg.py:85) -- /root/reference/src/madrona_learn/dists.py:119-208.  Pinned by
tests/golden/twohot.npz; the CUDA lowering is a "next" row."""
import numpy as np

F32 = np.float32


def symexp(x):                       # ml/utils.py:39-40
    return np.sign(x) * np.expm1(np.abs(x))


def bins(num_bins):                  # :127-141
    half = symexp(np.linspace(-14, 0, num_bins // 2 + 1, dtype=F32)).astype(F32)
    return np.concatenate([half, -half[:-1][::-1]]).astype(F32)


def _softmax(x):
    e = np.exp(x - x.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


def twohot_mean(logits):             # :143-170 (symmetric summation)
    b = bins(logits.shape[-1])
    mid = (b.size - 1) // 2
    p = _softmax(logits.astype(F32))
    p1, p2, p3 = p[..., :mid], p[..., mid:mid + 1], p[..., mid + 1:]
    b1, b2, b3 = b[:mid], b[mid:mid + 1], b[mid + 1:]
    return ((p2 * b2).sum(-1, keepdims=True) +
            ((p1 * b1)[..., ::-1] + (p3 * b3)).sum(-1, keepdims=True)).astype(F32)


def twohot_loss(logits, targets):    # :172-208
    b = bins(logits.shape[-1])
    n = b.size
    lo = np.clip((b <= targets).astype(np.int32).sum(-1) - 1, 0, n - 1)
    hi = np.clip(n - (b > targets).astype(np.int32).sum(-1), 0, n - 1)
    same = lo == hi
    dl = np.where(same[..., None], 1, np.abs(b[lo, None] - targets))
    du = np.where(same[..., None], 1, np.abs(b[hi, None] - targets))
    tot = dl + du
    wl, wu = dl / tot, du / tot
    oh = lambda i: (i[..., None] == np.arange(n)).astype(F32)
    two_hot = oh(lo) * wl + oh(hi) * wu
    l = logits.astype(F32)
    m = l.max(-1, keepdims=True)
    logp = l - (np.log(np.exp(l - m).sum(-1, keepdims=True)) + m)
    return (-(two_hot * logp).sum(-1, keepdims=True)).astype(F32)


# ---------------------------------------------------------------------------------------------
# HL-Gauss critic -- /root/reference/src/madrona_learn/models.py:177-306.  Pinned by
# tests/golden/hlgauss.npz (the reference's own HLGaussDist / HLGaussCritic.create executed under the shim).
# ---------------------------------------------------------------------------------------------
def hlgauss_bins(num_bins=127, min_bound=-100, max_bound=100):       # HLGaussCritic.create :271-283
    half = np.linspace(min_bound, 0, num_bins // 2 + 1)
    b = np.concatenate([half, -half[:-1][::-1]], axis=0)
    width = b[1] - b[0]
    bounds = b - 0.5 * width
    bounds = np.concatenate([bounds, np.asarray([bounds[-1] + width])], axis=0)
    return b.astype(F32), bounds.astype(F32)


def hlgauss_mean(logits, centers):                                    # HLGaussDist.mean :185-210
    mid = (centers.size - 1) // 2
    p = _softmax(logits.astype(F32))
    p1, p2, p3 = p[..., :mid], p[..., mid:mid + 1], p[..., mid + 1:]
    c1, c2, c3 = centers[:mid], centers[mid:mid + 1], centers[mid + 1:]
    return ((p2 * c2).sum(-1, keepdims=True) +
            ((p1 * c1)[..., ::-1] + (p3 * c3)).sum(-1, keepdims=True)).astype(F32)


def hlgauss_target(targets, centers, bounds, smoothness, dtype=F32):  # HLGaussDist.loss :212-247 (the histogram c)
    from scipy.special import erf
    f = dtype
    t = np.clip(targets.astype(f), centers[0], centers[-1])           # [..., 1]
    lo = (bounds <= t).astype(np.int32).sum(-1) - 1
    hi = lo + 1
    lo = np.clip(lo, 0, bounds.size - 2)
    hi = np.clip(hi, 1, bounds.size - 1)
    sig = (f(smoothness) * (bounds[hi] - bounds[lo]).astype(f))[..., None]
    cdfs = erf(((bounds.astype(f) - t) / (np.sqrt(f(2)) * sig)).astype(f)).astype(f)
    z = (cdfs[..., -1] - cdfs[..., 0])[..., None]
    return (1 / z * (cdfs[..., 1:] - cdfs[..., :-1])).astype(f)


def hlgauss_loss(logits, targets, centers, bounds, smoothness):       # :233-250
    c = hlgauss_target(targets, centers, bounds, smoothness)
    l = logits.astype(F32)
    m = l.max(-1, keepdims=True)
    logp = l - (np.log(np.exp(l - m).sum(-1, keepdims=True)) + m)
    return (-(c * logp).sum(-1, keepdims=True)).astype(F32)


# ---------------------------------------------------------------------------------------------
# ContinuousActionDistributions -- /root/reference/src/madrona_learn/dists.py:211-284.  Pinned by
# tests/golden/continuous.npz (the reference's own class executed under the shim).
# logits layout used by the lowering: raw means [rows, n] | raw stds [rows, n].
# ---------------------------------------------------------------------------------------------
def continuous_params(raw_means, raw_stds, stddev_min, stddev_max, dtype=F32):
    f = dtype
    mean = np.tanh(raw_means.astype(f))                                          # :230, :253, :272
    sig = 1 / (1 + np.exp(-(raw_stds.astype(f) + f(2.0))))
    std = (f(stddev_max) - f(stddev_min)) * sig + f(stddev_min)                  # :231, :273
    return mean, std


def continuous_action_stats(raw_means, raw_stds, actions, stddev_min, stddev_max, dtype=F32):   # :260-284
    f = dtype
    mean, std = continuous_params(raw_means, raw_stds, stddev_min, stddev_max, f)
    z = (actions.astype(f) - mean) / std
    log_probs = -0.5 * z * z - np.log(std) - f(0.5 * np.log(2 * np.pi))           # jax.scipy.stats.norm.logpdf
    entropies = 0.5 * np.log(2 * f(np.pi) * np.square(std)) + f(0.5)
    return log_probs.astype(f), entropies.astype(f)


def continuous_action_stats_bwd(raw_means, raw_stds, actions, stddev_min, stddev_max, dlogp, dent, dtype=np.float64):
    """Gradient of sum(dlogp * log_probs + dent * entropies) w.r.t. (raw_means, raw_stds)."""
    f = dtype
    mean, std = continuous_params(raw_means, raw_stds, stddev_min, stddev_max, f)
    sig = (std - f(stddev_min)) / (f(stddev_max) - f(stddev_min)) if stddev_max > stddev_min else np.zeros_like(std)
    d = actions.astype(f) - mean
    dmu = dlogp * d / (std * std)
    dsd = dlogp * (d * d / std ** 3 - 1 / std) + dent / std
    return dmu * (1 - mean * mean), dsd * (f(stddev_max) - f(stddev_min)) * sig * (1 - sig)


def erfinv_xla(x):
    """XLA's fp32 erf_inv (Giles' polynomial), evaluated by jax.random.normal on a uniform in (-1, 1)."""
    x = np.asarray(x, F32)
    w = (-np.log1p((-x * x).astype(F32))).astype(F32)
    lt = w < 5
    w1, w2 = (w - F32(2.5)).astype(F32), (np.sqrt(np.maximum(w, 0)).astype(F32) - F32(3)).astype(F32)
    c1 = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503,
          -0.00417768164, 0.246640727, 1.50140941]
    c2 = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613,
          0.00943887047, 1.00167406, 2.83297682]
    p1, p2 = np.full_like(x, c1[0]), np.full_like(x, c2[0])
    for a, b in zip(c1[1:], c2[1:]):
        p1 = (p1 * w1 + F32(a)).astype(F32)
        p2 = (p2 * w2 + F32(b)).astype(F32)
    return (np.where(lt, p1, p2) * x).astype(F32)


def continuous_sample(raw_means, raw_stds, key, stddev_min, stddev_max, partitionable=False):   # :216-246
    """sample_keys = split(prng_key, 1); actions = normal(sample_key) * std + mean; jax.random.normal =
    sqrt(2) * erf_inv(uniform(minval=nextafter(-1, 0), maxval=1)) on threefry bits (parity UNPINNED: jax is
    not installable here; restated from jax/_src/random.py `_normal_real`)."""
    from . import prng
    mean, std = continuous_params(raw_means, raw_stds, stddev_min, stddev_max)
    k = prng.split(key, 1, partitionable)[0]
    bits = prng.random_bits(k, mean.shape, partitionable)
    fl = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(F32) - F32(1.0)
    mn = np.nextafter(F32(-1.0), F32(0.0))
    u = np.maximum(mn, (fl * (F32(1.0) - mn) + mn).astype(F32))
    a = (F32(np.sqrt(2)) * erfinv_xla(u) * std + mean).astype(F32)
    lp, _ = continuous_action_stats(raw_means, raw_stds, a, stddev_min, stddev_max)
    return a, lp
