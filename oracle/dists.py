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
