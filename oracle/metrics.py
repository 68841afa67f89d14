"""Oracle (test infrastructure): Welford metric records.

Restates /root/reference/src/madrona_learn/metrics.py:12-98.
A metric is a dict(mean, m2, min, max: float32; count: int32).
"""
import numpy as np

F32 = np.float32
FMAX = np.finfo(np.float32).max
FMIN = np.finfo(np.float32).min


def metric_init():
    # :20-29
    return dict(mean=F32(0), m2=F32(0), min=F32(FMAX), max=F32(FMIN), count=np.int32(0))


def metric_from_data(data):
    """ml/metrics.py:31-48.  (init_from_data_masked, :50-67, ignores its mask: identical.)"""
    x = np.asarray(data)
    mean = F32(np.mean(x, dtype=np.float64))
    deltas = x.astype(np.float64) - np.float64(mean)
    return dict(mean=mean, m2=F32(np.sum(deltas * deltas)),
                min=F32(x.min()), max=F32(x.max()), count=np.int32(x.size))


def metric_merge(a, b):
    """Chan merge.  ml/metrics.py:79-98."""
    new_count = np.int32(a['count'] + b['count'])
    delta = F32(b['mean'] - a['mean'])
    safe_denom = F32(1) / np.maximum(F32(new_count), F32(1))
    mean = F32(a['mean'] + delta * F32(b['count']) * safe_denom)
    m2 = F32(a['m2'] + b['m2'] + delta * delta * F32(a['count']) * F32(b['count']) * safe_denom)
    return dict(mean=mean, m2=m2, min=F32(min(a['min'], b['min'])),
                max=F32(max(a['max'], b['max'])), count=new_count)
