"""Tensor-level wrappers over the C-ABI (one Python function per mlb_* entry point).

Inputs/outputs are CUDA torch tensors used purely as device buffers; each wrapper passes raw
pointers + the current stream to libmlb200 and never computes anything itself.
"""
import ctypes

import torch

from . import _lib
from ._lib import (Metric, c_float, c_int, c_ll, c_size_t, call, lib, ptr)

_METRIC_BYTES = ctypes.sizeof(Metric)   # 20


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def metrics_to_host(buf, n):
    """Read n mlb_metric records from a uint8 device buffer (host sync)."""
    raw = buf.cpu().numpy().tobytes()
    out = []
    for i in range(n):
        m = Metric.from_buffer_copy(raw[i * _METRIC_BYTES:(i + 1) * _METRIC_BYTES])
        out.append(dict(mean=m.mean, m2=m.m2, min=m.min, max=m.max, count=m.count))
    return out


# ---------------------------------------------------------------------------------------
# K1
# ---------------------------------------------------------------------------------------
def gae(rewards, values, dones, bootstrap, gamma, gae_lambda, advantages=None, returns=None,
        want_returns=True, vn_mu_sigma=None, metrics=None, ws=None):
    """mlb_gae_f32.  rewards/values f32 [T, N], dones u8/bool [T, N], bootstrap f32 [N].
    Returns (advantages, returns).  `metrics`: optional uint8 buffer of 4*20 bytes."""
    T, N = rewards.shape[0], rewards[0].numel()
    dev = rewards.device
    if advantages is None:
        advantages = torch.empty_like(rewards)
    if returns is None and want_returns:
        returns = torch.empty_like(rewards)
    d8 = dones.view(torch.uint8) if dones.dtype == torch.bool else dones
    wsb = 0
    if metrics is not None:
        need = lib().mlb_gae_workspace(T, N)
        if ws is None or ws.numel() < need:
            ws = _ws(need, dev)
        wsb = ws.numel()
    call('mlb_gae_f32', ptr(rewards), ptr(values), ptr(d8), ptr(bootstrap), ptr(advantages),
         ptr(returns), c_int(T), c_ll(N), c_float(gamma),
         c_float(float(gamma) * float(gae_lambda)), ptr(vn_mu_sigma), ptr(metrics),
         ptr(ws) if metrics is not None else ptr(None), c_size_t(wsb))
    return advantages, returns


def discounted_returns(rewards, dones, bootstrap, gamma, returns=None):
    """mlb_returns_f32."""
    T, N = rewards.shape[0], rewards[0].numel()
    if returns is None:
        returns = torch.empty_like(rewards)
    d8 = dones.view(torch.uint8) if dones.dtype == torch.bool else dones
    call('mlb_returns_f32', ptr(rewards), ptr(d8), ptr(bootstrap), ptr(returns), c_int(T),
         c_ll(N), c_float(gamma))
    return returns


# ---------------------------------------------------------------------------------------
# K2
# ---------------------------------------------------------------------------------------
def moments(x, var_floor=1e-5, out4=None, ws=None):
    n = x.numel()
    if out4 is None:
        out4 = torch.empty(4, dtype=torch.float32, device=x.device)
    need = lib().mlb_moments_workspace(n)
    if ws is None or ws.numel() < need:
        ws = _ws(need, x.device)
    call('mlb_moments_f32', ptr(x), c_ll(n), c_float(var_floor), ptr(out4), ptr(ws),
         c_size_t(ws.numel()))
    return out4


def zscore_apply(x, mean_rstd, out=None):
    if out is None:
        out = torch.empty_like(x)
    call('mlb_zscore_apply_f32', ptr(x), ptr(out), c_ll(x.numel()), ptr(mean_rstd))
    return out


def zscore(x, out=None, out4=None, ws=None):
    n = x.numel()
    if out is None:
        out = torch.empty_like(x)
    if out4 is None:
        out4 = torch.empty(4, dtype=torch.float32, device=x.device)
    need = lib().mlb_moments_workspace(n)
    if ws is None or ws.numel() < need:
        ws = _ws(need, x.device)
    call('mlb_zscore_f32', ptr(x), ptr(out), c_ll(n), ptr(out4), ptr(ws), c_size_t(ws.numel()))
    return out


def metric(x, out=None, ws=None):
    n = x.numel()
    if out is None:
        out = torch.empty(_METRIC_BYTES, dtype=torch.uint8, device=x.device)
    need = lib().mlb_moments_workspace(n)
    if ws is None or ws.numel() < need:
        ws = _ws(need, x.device)
    call('mlb_metric_f32', ptr(x), c_ll(n), ptr(out), ptr(ws), c_size_t(ws.numel()))
    return out


def traj_moments(x, C, out=None):
    T, N = x.shape[0], x[0].numel()
    if out is None:
        out = torch.empty((C * N, 2), dtype=torch.float64, device=x.device)
    call('mlb_traj_moments_f32', ptr(x), c_int(T), c_ll(N), c_int(C), ptr(out))
    return out


def mb_moments(tm, perm, M, Tp, var_floor=1e-5, out=None, raw_out=None):
    E, J = perm.shape
    if out is None and raw_out is None:
        out = torch.empty((E * (J // M), 4), dtype=torch.float32, device=perm.device)
    call('mlb_mb_moments_f32', ptr(tm), ptr(perm), c_int(E), c_ll(J), c_ll(M), c_int(Tp),
         c_float(var_floor), ptr(out), ptr(raw_out))
    return out


def moments_finalize(raw, count, var_floor, out):
    call('mlb_moments_finalize_f32', ptr(raw), c_int(raw.shape[0]), ctypes.c_double(count),
         c_float(var_floor), ptr(out))
    return out


# ---------------------------------------------------------------------------------------
# K3
# ---------------------------------------------------------------------------------------
def ema_state_init(dim, device):
    """[5*dim floats | int32 N]: mu=0, inv_sigma=1, sigma=1, mu_biased=0, sigma_sq_biased=0."""
    st = torch.zeros(5 * dim + 1, dtype=torch.float32, device=device)
    st[dim:3 * dim] = 1.0
    return st


def ema_update(state, dim, batch_mean, batch_var, decay, eps=1e-5):
    call('mlb_ema_update_f32', ptr(state), c_int(dim), ptr(batch_mean), ptr(batch_var),
         c_float(decay), c_float(eps))
    return state


def ema_scan(state, mbm, decay, eps=1e-5, out=None):
    K = mbm.shape[0]
    if out is None:
        out = torch.empty((K, 4), dtype=torch.float32, device=mbm.device)
    call('mlb_ema_scan_f32', ptr(state), ptr(mbm), c_int(K), c_float(decay), c_float(eps),
         ptr(out))
    return out


def ema_normalize(state, dim, x, out=None):
    if out is None:
        out = torch.empty_like(x)
    call('mlb_ema_normalize_f32', ptr(state), c_int(dim), ptr(x), ptr(out),
         c_ll(x.numel() // dim))
    return out


def ema_invert(state, dim, x, out=None):
    if out is None:
        out = torch.empty_like(x)
    call('mlb_ema_invert_f32', ptr(state), c_int(dim), ptr(x), ptr(out), c_ll(x.numel() // dim))
    return out


def obs_moments(obs, raw_out):
    """mlb_obs_moments_f32: obs f32 [..., D] -> raw f64 [D, 2] (sum, sumsq over all leading rows)."""
    D = obs.shape[-1]
    call('mlb_obs_moments_f32', ptr(obs), c_ll(obs.numel() // D), c_int(D), ptr(raw_out))
    return raw_out


def obs_stats_merge(raw, count, mean=None, var=None):
    """mlb_obs_stats_merge_f32: raw f64 [T, D, 2] -> (mean, var) f32 [D]."""
    T, D = raw.shape[0], raw.shape[1]
    if mean is None:
        mean = torch.empty(D, dtype=torch.float32, device=raw.device)
        var = torch.empty(D, dtype=torch.float32, device=raw.device)
    call('mlb_obs_stats_merge_f32', ptr(raw), c_int(T), ctypes.c_double(count), c_int(D), ptr(mean), ptr(var))
    return mean, var


def ema_estimate_init(device):
    """EMAEstimate state: [mu, mu_biased | int32 N]."""
    return torch.zeros(3, dtype=torch.float32, device=device)


def ema_estimate_update(state, x, decay):
    call('mlb_ema_estimate_update_f32', ptr(state), ptr(x), c_float(decay))
    return state


def sort_u64(keys):
    """In-place ascending sort of a power-of-two-length int64 tensor holding u64 keys."""
    call('mlb_sort_u64', ptr(keys), c_ll(keys.numel()))
    return keys


def sort_pad(n):
    return int(lib().mlb_sort_pad(c_ll(n)))


# ---------------------------------------------------------------------------------------
# PRNG
# ---------------------------------------------------------------------------------------
def threefry_split(key, num, partitionable=False):
    out = torch.empty((num, 2), dtype=torch.int32, device=key.device)
    call('mlb_threefry_split', ptr(key), ptr(out), c_int(num), c_int(int(partitionable)))
    return out


def threefry_bits(key, n, partitionable=False):
    out = torch.empty(n, dtype=torch.int32, device=key.device)
    call('mlb_threefry_bits', ptr(key), ptr(out), c_ll(n), c_int(int(partitionable)))
    return out


def ppo_permutations(key, E, J, partitionable=False, perm=None, ws=None):
    """Advances `key` (int32[2] device tensor holding the uint32 words) in place."""
    if perm is None:
        perm = torch.empty((E, J), dtype=torch.int32, device=key.device)
    need = lib().mlb_ppo_permutations_workspace(E, J)
    if ws is None or ws.numel() < need:
        ws = _ws(need, key.device)
    call('mlb_ppo_permutations', ptr(key), ptr(perm), c_int(E), c_ll(J),
         c_int(int(partitionable)), ptr(ws), c_size_t(ws.numel()))
    return perm


# ---------------------------------------------------------------------------------------
# K5
# ---------------------------------------------------------------------------------------
def mb_gather(store, idx, C, Tp, B, out=None):
    """store [C, T', B, *leaf] (P=1 squeezed) -> [T', M, *leaf]."""
    M = idx.numel()
    leaf = store.shape[3:] if store.dim() > 3 else ()
    row_elems = 1
    for d in leaf:
        row_elems *= d
    row_bytes = row_elems * store.element_size()
    if out is None:
        out = torch.empty((Tp, M, *leaf), dtype=store.dtype, device=store.device)
    call('mlb_mb_gather', ptr(store), ptr(idx), ptr(out), c_int(C), c_int(Tp), c_ll(B), c_ll(M),
         c_ll(row_bytes))
    return out


def mb_gather_multi(leaves, idx, C, Tp, B):
    """leaves: list of (store [C, T', B, *leaf], out [T', M, *leaf] or None, out_bf16 or None)."""
    arr = (_lib.GatherLeaf * len(leaves))()
    for i, (store, out, out_bf16) in enumerate(leaves):
        row_elems = 1
        for d in store.shape[3:]:
            row_elems *= d
        arr[i].store = store.data_ptr()
        arr[i].out = out.data_ptr() if out is not None else None
        arr[i].out_bf16 = out_bf16.data_ptr() if out_bf16 is not None else None
        arr[i].row_bytes = row_elems * store.element_size()
    call('mlb_mb_gather_multi', arr, c_int(len(leaves)), ptr(idx), c_int(C), c_int(Tp), c_ll(B),
         c_ll(idx.numel()))


def mb_gather_multi_peer(leaves, peer_table, world, idx, C, Tp, B):
    """mlb_mb_gather_multi_peer.  leaves: list of (local store [C, T', 1, B, *leaf] (shape/dtype only),
    out, out_bf16); peer_table: HOST void*[world][len(leaves)] (DistContext.peer_store_table)."""
    arr = (_lib.GatherLeaf * len(leaves))()
    for i, (store, out, out_bf16) in enumerate(leaves):
        row_elems = store.numel() // (C * Tp * B)       # P = 1: [C, T', 1, B, *leaf] (or [C, 1, B, *] with Tp = 1)
        arr[i].store = None
        arr[i].out = out.data_ptr() if out is not None else None
        arr[i].out_bf16 = out_bf16.data_ptr() if out_bf16 is not None else None
        arr[i].row_bytes = row_elems * store.element_size()
    call('mlb_mb_gather_multi_peer', arr, c_int(len(leaves)), peer_table, c_int(world), ptr(idx), c_int(C),
         c_int(Tp), c_ll(B), c_ll(idx.numel()))


def mb_gather_rnn(store, idx, C, B, out=None):
    M = idx.numel()
    leaf = store.shape[2:]
    row_elems = 1
    for d in leaf:
        row_elems *= d
    row_bytes = row_elems * store.element_size()
    if out is None:
        out = torch.empty((M, *leaf), dtype=store.dtype, device=store.device)
    call('mlb_mb_gather_rnn', ptr(store), ptr(idx), ptr(out), c_int(C), c_ll(B), c_ll(M),
         c_ll(row_bytes))
    return out
