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
