"""PBT policy-batch reorder: the step in front of policy inference on the
multi-policy path.  Mirrors `PolicyBatchReorderState` (ml/rollouts.py:137-168) and
`_compute_reorder_chunks` (ml/rollouts.py:1107-1190); the index construction is
`mlb_reorder_chunks` (stable counting sort on the device, integer-only, bit-exact against the
reference's KAT vectors) and the two gathers are `mlb_gather_rows_clip`.
"""
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from ._lib import c_int, c_ll, c_size_t, call, lib, ptr


def _compute_reorder_chunks(assignments: torch.Tensor, P: int, C: int, B: int):
    """assignments int32 [S] (values in [0, P)) -> (to_policy_idxs int32 [B, C], to_sim_idxs int32 [S])."""
    assert assignments.dim() == 1 and assignments.dtype == torch.int32 and assignments.is_cuda
    S = assignments.numel()
    dev = assignments.device
    to_policy = torch.empty(B, C, dtype=torch.int32, device=dev)
    to_sim = torch.empty(S, dtype=torch.int32, device=dev)
    ws = torch.empty(lib().mlb_reorder_chunks_workspace(S, P) + 16, dtype=torch.uint8, device=dev)
    call('mlb_reorder_chunks', ptr(assignments), c_ll(S), c_int(P), c_int(C), c_ll(B), ptr(to_policy),
         ptr(to_sim), ptr(ws), c_size_t(ws.numel()))
    return to_policy, to_sim


def _gather_rows(x, idx, n_out_shape):
    x = x.contiguous()
    row = 1
    for d in x.shape[1:]:
        row *= d
    out = torch.empty(*n_out_shape, *x.shape[1:], dtype=x.dtype, device=x.device)
    call('mlb_gather_rows_clip', ptr(x.view(torch.uint8) if x.dtype == torch.bool else x), ptr(idx),
         ptr(out.view(torch.uint8) if out.dtype == torch.bool else out), c_ll(idx.numel()), c_ll(x.shape[0]),
         c_ll(row * x.element_size()))
    return out


@dataclass
class PolicyBatchReorderState:                       # ml/rollouts.py:137-168
    to_policy_idxs: Optional[torch.Tensor]
    to_sim_idxs: Optional[torch.Tensor]
    policy_dims: Tuple[int, ...]
    sim_dims: Tuple[int, ...]

    def to_policy(self, data):
        def txfm(x):
            if self.to_policy_idxs is None:
                return x.reshape(*self.policy_dims, *x.shape[1:])
            return _gather_rows(x, self.to_policy_idxs, tuple(self.to_policy_idxs.shape))   # mode='clip'
        return {k: txfm(v) for k, v in data.items()} if isinstance(data, dict) else txfm(data)

    def to_sim(self, data):
        def txfm(x):
            if self.to_sim_idxs is None:
                return x.reshape(*self.sim_dims, *x.shape[2:])
            flat = x.reshape(x.shape[0] * x.shape[1], *x.shape[2:])
            return _gather_rows(flat, self.to_sim_idxs, (self.to_sim_idxs.numel(),))
        return {k: txfm(v) for k, v in data.items()} if isinstance(data, dict) else txfm(data)


def reorder_state_for(assignments, P, C):
    """B = S // C + P chunks always suffice (at most S // C full chunks + one partial per policy);
    the reference's test sizing S // C + P - 1 (tests/test_rollouts.py:36-44) is one short when
    P == 1 and C does not divide S."""
    S = assignments.numel()
    B = S // C + P
    tp, ts = _compute_reorder_chunks(assignments, P, C, B)
    return PolicyBatchReorderState(tp, ts, (B, C), (S,))
