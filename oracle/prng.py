"""Oracle (test infrastructure): JAX threefry PRNG -- key/split/fold_in/random_bits/permutation.

PARITY UNPINNED against a live jax (jax is not installable in this image and is not under
/root/reference; the reference leaves its jax version unpinned, pyproject.toml:9-15).
Restated from the published algorithm of jax/_src/prng.py and jax/_src/random.py
(Threefry-2x32, 20 rounds, Random123).  Call sites on the path: ml/train.py:284-289,
ml/train_state.py:134-136,288,465-477, ml/rollouts.py:878-880, ml/ppo.py:451,
ml/dists.py:30-33.

Known-answer vectors held in tests/golden/threefry_kat.json:
  * the three Random123 Threefry-2x32x20 KATs (also used by jax's own random_test.py);
  * split(PRNGKey(0)) == [[4146024105, 967050713], [2718843009, 1272950319]]
    (the value printed in the public JAX PRNG tutorial; non-partitionable layout).

Two counter layouts are implemented because jax changed its default in 0.5.0:
  partitionable=False  (jax < 0.5 default; OUR default)  -- counters iota(2n) split in halves
  partitionable=True   (jax >= 0.5 default)               -- 64-bit iota as (hi, lo) words
"""
import math
import numpy as np

U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds.  All args uint32 arrays/scalars; returns (y0, y1)."""
    with np.errstate(over='ignore'):
        k0 = np.asarray(k0, U32)
        k1 = np.asarray(k1, U32)
        x0 = np.array(x0, dtype=U32, copy=True)
        x1 = np.array(x1, dtype=U32, copy=True)
        ks = (k0, k1, k0 ^ k1 ^ U32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x0 ^ x1
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + U32(g + 1)
    return x0, x1


def key(seed):
    """random.key(seed) / PRNGKey(seed): [hi32, lo32]."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], U32)


def _threefry_2x32_flat(keypair, counts):
    """jax's threefry_2x32(keypair, count): flatten, pad to even, split in two halves."""
    counts = np.asarray(counts, U32).ravel()
    odd = counts.size % 2
    if odd:
        counts = np.concatenate([counts, np.zeros(1, U32)])
    half = counts.size // 2
    y0, y1 = threefry2x32(keypair[0], keypair[1], counts[:half], counts[half:])
    out = np.concatenate([y0, y1])
    return out[:-1] if odd else out


def _iota_2x32(n):
    i = np.arange(n, dtype=np.uint64)
    return (i >> np.uint64(32)).astype(U32), (i & np.uint64(0xFFFFFFFF)).astype(U32)


def split(k, num=2, partitionable=False):
    """random.split(key, num) -> [num, 2] uint32."""
    k = np.asarray(k, U32)
    if partitionable:
        c1, c2 = _iota_2x32(num)
        b1, b2 = threefry2x32(k[0], k[1], c1, c2)
        return np.stack([b1, b2], axis=1)
    return _threefry_2x32_flat(k, np.arange(2 * num, dtype=U32)).reshape(num, 2)


def fold_in(k, data):
    """random.fold_in(key, data): threefry_2x32(key, [0, data])."""
    k = np.asarray(k, U32)
    return _threefry_2x32_flat(k, np.array([0, int(data) & 0xFFFFFFFF], U32))


def random_bits(k, shape, partitionable=False):
    """random.bits(key, shape, uint32)."""
    k = np.asarray(k, U32)
    size = int(np.prod(shape)) if len(shape) else 1
    if partitionable:
        c1, c2 = _iota_2x32(size)
        b1, b2 = threefry2x32(k[0], k[1], c1, c2)
        return (b1 ^ b2).reshape(shape)
    return _threefry_2x32_flat(k, np.arange(size, dtype=U32)).reshape(shape)


def shuffle_rounds(size):
    """jax/_src/random.py::_shuffle -- ceil(3 ln(size) / ln(2^32-1))."""
    return int(np.ceil(3 * np.log(max(1, size)) / np.log(np.iinfo(np.uint32).max)))


def permutation(k, x, partitionable=False):
    """random.permutation(key, x) for 1-D x (or int -> arange): repeated stable sort by
    random 32-bit keys (ml/ppo.py:451)."""
    x = np.arange(x, dtype=np.int32) if np.isscalar(x) else np.array(x, copy=True)
    k = np.asarray(k, U32)
    for _ in range(shuffle_rounds(x.size)):
        ks = split(k, 2, partitionable)
        k, sub = ks[0], ks[1]
        bits = random_bits(sub, x.shape, partitionable)
        x = x[np.argsort(bits, kind='stable')]
    return x


def uniform_from_bits(bits, minval=np.finfo(np.float32).tiny, maxval=1.0):
    """jax.random.uniform float32 from raw bits: 23 mantissa bits -> [1,2) - 1."""
    fb = (np.asarray(bits, U32) >> U32(9)) | U32(0x3F800000)
    f = fb.view(np.float32) - np.float32(1.0)
    lo, hi = np.float32(minval), np.float32(maxval)
    return np.maximum(lo, f * (hi - lo) + lo).astype(np.float32)


def gumbel_from_bits(bits):
    """jax.random.gumbel: -log(-log(uniform(tiny, 1)))."""
    u = uniform_from_bits(bits)
    return (-np.log(-np.log(u))).astype(np.float32)


def categorical(k, logits, partitionable=False):
    """random.categorical(key, logits) = argmax(logits + gumbel) over the last axis
    (ml/dists.py:33).  Sampling parity with the CUDA path is distributional only."""
    logits = np.asarray(logits, np.float32)
    g = gumbel_from_bits(random_bits(k, logits.shape, partitionable))
    return np.argmax(g + logits, axis=-1).astype(np.int32)


def update_epoch_keys(update_prng_key, num_epochs, partitionable=False):
    """The key stream of ml/ppo.py:445-451 + ml/train_state.py:134-136: per epoch
    (rnd, next) = split(key); key <- next.  Returns ([rnd_e], final_key)."""
    k = np.asarray(update_prng_key, U32)
    out = []
    for _ in range(num_epochs):
        ks = split(k, 2, partitionable)
        out.append(ks[0])
        k = ks[1]
    return out, k
