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
            n_pad = ((n + 3) // 4) * 4
            SIG = 64                                   # uint32 words of signal region (2 * MLB_MAX_PEERS used)
            # symmetric layout: [ gradient arena | reduced gradient (NVLS broadcast target) | signals ]
            arena = symm_mem.empty(2 * n_pad + SIG, dtype=torch.float32, device=prog.device)
            arena.zero_()
            hdl = symm_mem.rendezvous(arena, self.group if self.group is not None else dist.group.WORLD)
            ptrs = list(hdl.buffer_ptrs)
            mc = int(hdl.multicast_ptr or 0)
        except Exception as e:                         # noqa: BLE001 -- optional fast path
            self.fused_error = repr(e)
            return False
        tab = _lib.PeerTable()
        tab.rank, tab.world = self.rank, self.world_size
        for r in range(self.world_size):
            tab.grads[r] = ptrs[r]
            tab.signals[r] = ptrs[r] + 2 * n_pad * 4
        self._symm = (arena, hdl)                      # keep the mapping alive
        self._peer_tab = tab
        # MLB_ALLREDUCE = auto | p2p | nvls.  auto: the NVSwitch in-network reduction (multimem) when a
        # multicast mapping exists and world > 2 (measured per call: 8 GPUs 19.6 us NVLS / 28.0 us
        # one-shot peer loads / 28.9 us NCCL + norm; 2 GPUs 18.8 / 14.9 / 19.6), else peer loads.
        mode = os.environ.get('MLB_ALLREDUCE', 'auto')
        self.nvls = bool(mc) and mode != 'p2p' and (mode == 'nvls' or self.world_size > 2)
        self._mc_grads, self._mc_out = mc, mc + n_pad * 4
        self._reduced = arena[n_pad:n_pad + n] if self.nvls else torch.zeros(n, dtype=torch.float32, device=prog.device)
        self._ar_state = torch.zeros(4, dtype=torch.int32, device=prog.device)
        self._ar_ws = torch.zeros(_lib.lib().mlb_allreduce_workspace(), dtype=torch.uint8, device=prog.device)
        prog.adopt_grad_arena(arena[:n])
        torch.cuda.synchronize()
        dist.barrier(group=self.group)                 # every rank's arena + signals are zeroed
        self.fused = True
        return True

    def allreduce_grads_fused(self, prog):
        """-> reduced gradient tensor; prog.grad_sumsq holds sum(g^2) of it."""
        from ._lib import c_ll, c_size_t, call, ptr
        if self.nvls:
            call('mlb_allreduce_nvls_f32', ctypes.byref(self._peer_tab), ctypes.c_void_p(self._mc_grads),
                 ctypes.c_void_p(self._mc_out), ptr(self._reduced), c_ll(prog.num_params), ptr(prog.grad_sumsq),
                 ptr(self._ar_state), ptr(self._ar_ws), c_size_t(self._ar_ws.numel()))
        else:
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
