"""Data-parallel context: one process per GPU, worlds sharded across ranks.

The reference is single-device (ml/train.py:131-146); its semantics extended to R ranks are
"equal to the 1-GPU reference on the concatenated data":
  * rollout, GAE, gather are rank-local (no communication);
  * per-minibatch z-score / value-normaliser statistics are over the GLOBAL minibatch: raw
    (sum, sumsq) of all minibatches of the update are all-reduced ONCE, right after GAE;
  * the loss kernel already divides by the global row count, so gradients are SUM-reduced:
    one NCCL all-reduce over the flat gradient arena per minibatch (captured in the update
    graph); clip_by_global_norm then sees the norm of the reduced gradient on every rank.
Permutation modes:
  * "index_exact" (default on NCCL when NVLink symmetric memory is available): every rank
    generates the SAME permutation of the GLOBAL trajectory ids (ml/ppo.py:437-451 on the
    concatenated data) and trains on slice [r*M/R, (r+1)*M/R) of every global minibatch; rows it
    does not own are fetched from the owner's rollout store by NVLink peer loads inside the gather
    kernel (mlb_mb_gather_multi_peer).  The R-rank run then sees bit-identical minibatches to the
    1-GPU run on the concatenated worlds (tools/dp_update_check.py, tests/test_dp_gpu.py).
  * "fast" (MLB_DP_PERM=fast, and the fallback): every rank permutes its local trajectories
    (same key stream on all ranks); equal in distribution, not index-exact.

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
        self.perm_mode = 'fast'
        self.peer_stores = None

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


    # -----------------------------------------------------------------------------------
    # index-exact global permutation: rollout stores in NVLink symmetric memory
    # -----------------------------------------------------------------------------------
    def alloc_symmetric_stores(self, specs, device):
        """specs: {name: (shape, dtype)}.  Allocates every leaf inside ONE symmetric (peer-mapped)
        buffer, exchanges the mappings and records peer_stores[r][name] = device pointer of rank r's
        leaf in THIS process' address space.  Returns {name: tensor} or None when unavailable (not
        NCCL, no symmetric memory, MLB_DP_PERM=fast, more than 8 ranks): the caller then allocates
        ordinary tensors and the run uses the "fast" permutation mode."""
        if (os.environ.get('MLB_DP_PERM', 'index_exact') == 'fast' or dist.get_backend(self.group) != 'nccl' or
                self.world_size > 8):
            return None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            offs, total = {}, 0
            for name, (shape, dtype) in specs.items():
                n = 1
                for d in shape:
                    n *= int(d)
                offs[name] = total
                total += (n * torch.empty(0, dtype=dtype).element_size() + 255) // 256 * 256
            buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
            hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
            ptrs = list(hdl.buffer_ptrs)
        except Exception as e:                         # noqa: BLE001 -- optional fast path
            self.symm_store_error = repr(e)
            return None
        out = {}
        for name, (shape, dtype) in specs.items():
            n = 1
            for d in shape:
                n *= int(d)
            nbytes = n * torch.empty(0, dtype=dtype).element_size()
            out[name] = buf[offs[name]:offs[name] + nbytes].view(dtype).view(*shape)
        self._symm_store = (buf, hdl)                  # keep the mapping alive
        self.peer_stores = [{name: int(ptrs[r]) + offs[name] for name in specs} for r in range(self.world_size)]
        self.perm_mode = 'index_exact'
        return out

    def peer_store_table(self, names):
        """HOST array of device pointers [world][len(names)] for mlb_mb_gather_multi_peer."""
        arr = (ctypes.c_void_p * (self.world_size * len(names)))()
        for r in range(self.world_size):
            for i, n in enumerate(names):
                arr[r * len(names) + i] = self.peer_stores[r][n]
        return arr

    def allgather_traj_moments(self, tm_local, tm_global, C):
        """tm_local f64 [C*B, 2] (trajectory j = c*B + b of this rank) -> tm_global [C*world*B, 2] in
        GLOBAL trajectory order j = c*(world*B) + r*B + b: one all-gather per BPTT chunk.  Being a
        collective that every rank enters after its GAE, it is also the point after which peers'
        rollout stores may be read."""
        B = tm_local.shape[0] // C
        lg = tm_global.view(C, self.world_size * B, 2)
        ll = tm_local.view(C, B, 2)
        for c in range(C):
            dist.all_gather_into_tensor(lg[c], ll[c], group=self.group)
        return tm_global

    def stream_barrier(self, device):
        """Cross-rank ordering point on the current stream (graph-capturable): a 1-element all-reduce."""
        if getattr(self, '_bar', None) is None:
            self._bar = torch.zeros(1, dtype=torch.float32, device=device)
        dist.all_reduce(self._bar, group=self.group)

    def allreduce_raw_moments(self, raw):
        """raw f64 [K, 2] = per-minibatch (sum, sumsq) of this rank's shard."""
        return self.allreduce_sum(raw)

    def max_over_ranks(self, seconds, device):
        t = torch.tensor([seconds], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def barrier(self):
        dist.barrier(group=self.group)
