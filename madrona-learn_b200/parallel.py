"""Data-parallel context: one process per GPU, worlds sharded across ranks.

The reference is single-device (ml/train.py:131-146); its semantics extended to R ranks are
"equal to the 1-GPU reference on the concatenated data":
  * rollout, GAE, gather are rank-local (no communication);
  * per-minibatch z-score / value-normaliser statistics are over the GLOBAL minibatch: raw
    (sum, sumsq) of all minibatches of the update are all-reduced ONCE, right after GAE;
  * the loss kernel already divides by the global row count, so gradients are SUM-reduced:
    one NCCL all-reduce over the flat gradient arena per minibatch (captured in the update
    graph); clip_by_global_norm then sees the norm of the reduced gradient on every rank.
Permutation mode: "fast" -- every rank permutes its local trajectories (same key stream on
all ranks); the index-exact global permutation (peer row fetch) is a later row.

Works with any torch.distributed backend: NCCL over NVLink on the B200 box, gloo in the CPU
tests of the host-side logic (tests/test_parallel_cpu.py).
"""
import ctypes
import os

import torch
import torch.distributed as dist


class DistContext:
    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError('torch.distributed is not initialised')
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)
        self._raw = {}
        self.fused = False

    def rank_seed(self, seed):
        """Rollout / environment randomness differs per rank; parameters do not."""
        return int(seed) + 1000003 * self.rank

    def shard_worlds(self, num_worlds):
        if num_worlds % self.world_size:
            raise ValueError('num_worlds must divide evenly across ranks (ml/ppo.py:439 analogue)')
        per = num_worlds // self.world_size
        return per, self.rank * per

    def allreduce_sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_grads(self, flat_grads):
        return self.allreduce_sum(flat_grads)

    # -----------------------------------------------------------------------------------
    # fused NVLink all-reduce (+ global-norm reduction) of the gradient arena
    # -----------------------------------------------------------------------------------
    def enable_fused_allreduce(self, prog):
        """Move `prog`'s gradient arena into symmetric (peer-mapped) memory and exchange the peer
        pointers; afterwards allreduce_grads_fused() replaces the NCCL all-reduce + the norm
        reduction with one kernel (mlb_allreduce_sumsq_f32).  Returns False (and changes nothing)
        when the backend is not NCCL, symmetric memory is unavailable or MLB_FUSED_ALLREDUCE=0."""
        self.fused = False
        if os.environ.get('MLB_FUSED_ALLREDUCE', '1') == '0' or dist.get_backend(self.group) != 'nccl':
            return False
        if self.world_size > 16:
            return False
        try:
            import torch.distributed._symmetric_memory as symm_mem
            from . import _lib
            n = prog.num_params
            SIG = 64                                   # uint32 words of signal region (2 * MLB_MAX_PEERS used)
            arena = symm_mem.empty(((n + 3) // 4) * 4 + SIG, dtype=torch.float32, device=prog.device)
            arena.zero_()
            hdl = symm_mem.rendezvous(arena, self.group if self.group is not None else dist.group.WORLD)
            ptrs = list(hdl.buffer_ptrs)
        except Exception as e:                         # noqa: BLE001 -- optional fast path
            self.fused_error = repr(e)
            return False
        tab = _lib.PeerTable()
        tab.rank, tab.world = self.rank, self.world_size
        sig_off = (((n + 3) // 4) * 4) * 4
        for r in range(self.world_size):
            tab.grads[r] = ptrs[r]
            tab.signals[r] = ptrs[r] + sig_off
        self._symm = (arena, hdl)                      # keep the mapping alive
        self._peer_tab = tab
        self._reduced = torch.zeros(n, dtype=torch.float32, device=prog.device)
        self._ar_state = torch.zeros(2, dtype=torch.int32, device=prog.device)
        self._ar_ws = torch.zeros(_lib.lib().mlb_allreduce_workspace(), dtype=torch.uint8, device=prog.device)
        prog.adopt_grad_arena(arena[:n])
        torch.cuda.synchronize()
        dist.barrier(group=self.group)                 # every rank's arena + signals are zeroed
        self.fused = True
        return True

    def allreduce_grads_fused(self, prog):
        """-> reduced gradient tensor; prog.grad_sumsq holds sum(g^2) of it."""
        from ._lib import c_ll, c_size_t, call, ptr
        call('mlb_allreduce_sumsq_f32', ctypes.byref(self._peer_tab), ptr(self._reduced), c_ll(prog.num_params),
             ptr(prog.grad_sumsq), ptr(self._ar_state), ptr(self._ar_ws), c_size_t(self._ar_ws.numel()))
        return self._reduced

    def allreduce_raw_moments(self, raw):
        """raw f64 [K, 2] = per-minibatch (sum, sumsq) of this rank's shard."""
        return self.allreduce_sum(raw)

    def max_over_ranks(self, seconds, device):
        t = torch.tensor([seconds], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def barrier(self):
        dist.barrier(group=self.group)
