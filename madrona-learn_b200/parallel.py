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

    def allreduce_raw_moments(self, raw):
        """raw f64 [K, 2] = per-minibatch (sum, sumsq) of this rank's shard."""
        return self.allreduce_sum(raw)

    def max_over_ranks(self, seconds, device):
        t = torch.tensor([seconds], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def barrier(self):
        dist.barrier(group=self.group)
