"""CPU, world_size 2, gloo: the data-parallel host logic (madrona_learn_b200.parallel) --
statistics all-reduce once per update, gradient SUM all-reduce with globally scaled losses --
reproduces the single-process result on the concatenated data."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUCKETS = [4, 3, 2]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_mb(rng, T, M, D):
    A = len(BUCKETS)
    return dict(obs=rng.standard_normal((T, M, D)).astype(np.float32),
                actions=np.stack([rng.integers(0, b, (T, M)) for b in BUCKETS], -1).astype(np.int32),
                log_probs=(-np.abs(rng.standard_normal((T, M, A))) - 0.3).astype(np.float32),
                advantages=(rng.standard_normal((T, M, 1)) * 2 + 1).astype(np.float32),
                returns=(rng.standard_normal((T, M, 1)) * 3 - 1).astype(np.float32),
                values=rng.standard_normal((T, M, 1)).astype(np.float32),
                mb_weights=np.ones((M, 1), np.float32))


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import madrona_learn_b200  # noqa: F401  (package import must work without a GPU)
    from madrona_learn_b200.parallel import DistContext
    from oracle import nn as onn, ppo as oppo
    from oracle.moving_avg import EMANormalizer
    ctx = DistContext()
    assert ctx.world_size == world and ctx.rank == rank
    assert ctx.shard_worlds(64) == (32, 32 * rank)
    with pytest.raises(ValueError):
        ctx.shard_worlds(63)
    assert ctx.rank_seed(7) != DistContext.rank_seed(type('X', (), {'rank': rank + 1})(), 7)
    # the fused NVLink all-reduce is an NCCL-only fast path: under gloo it must decline cleanly
    # (nothing allocated, the program's gradient arena untouched) and leave the collective path
    assert ctx.enable_fused_allreduce(object()) is False and ctx.fused is False

    rng = np.random.default_rng(0)                       # same stream on both ranks
    T, M, D = 4, 16, 6
    full = _make_mb(rng, T, M, D)
    params = onn.init_params(np.random.default_rng(1), D, 16, 2, BUCKETS, dtype=np.float64)
    params['actor']['kernel'] = np.random.default_rng(2).standard_normal(params['actor']['kernel'].shape) * 0.4
    cfg = oppo.PPOCfg(BUCKETS, normalize_values=True, entropy_coef=0.02)
    norm = EMANormalizer(cfg.value_normalizer_decay)
    vn0 = norm.update_estimates(norm.init_estimates(1), (np.array([0.5], np.float32), np.array([3.0], np.float32)))
    ref = oppo.ppo_loss(params, full, cfg, vn0, dtype=np.float64)

    # this rank's shard of the minibatch (worlds split across ranks)
    sl = slice(rank * M // world, (rank + 1) * M // world)
    mine = {k: (v[:, sl] if k != 'mb_weights' else v[sl]) for k, v in full.items()}
    # (1) raw moments, ONE all-reduce for advantages + returns
    raw = torch.tensor([[mine['advantages'].astype(np.float64).sum(), np.square(mine['advantages'].astype(np.float64)).sum()],
                        [mine['returns'].astype(np.float64).sum(), np.square(mine['returns'].astype(np.float64)).sum()]],
                       dtype=torch.float64)
    ctx.allreduce_raw_moments(raw)
    n = T * M
    mean = raw[:, 0] / n
    var = raw[:, 1] / n - mean * mean
    adv_stats = (np.float32(mean[0]), np.float32(1.0 / np.sqrt(max(float(var[0]), 1e-5))))
    vn1 = norm.update_estimates(vn0, (np.array([mean[1]], np.float32), np.array([var[1]], np.float32)))
    for k in ('mu', 'inv_sigma'):
        np.testing.assert_allclose(vn1[k], ref['new_vn_state'][k], rtol=1e-5)
    # (2) local loss / grads with global statistics; local mean -> scale by local/global rows
    out = oppo.ppo_loss(params, mine, cfg, vn0, dtype=np.float64, adv_stats=adv_stats, new_vn_state=vn1)
    leaves = onn.tree_leaves(out['grads'])
    flat = torch.from_numpy(np.concatenate([g.reshape(-1) for g in leaves])) * (1.0 / world)
    ctx.allreduce_grads(flat)                            # SUM over ranks
    ref_flat = np.concatenate([g.reshape(-1) for g in onn.tree_leaves(ref['grads'])])
    np.testing.assert_allclose(flat.numpy(), ref_flat, rtol=2e-4, atol=1e-9)
    loss = torch.tensor([float(out['loss']) / world], dtype=torch.float64)
    ctx.allreduce_sum(loss)
    np.testing.assert_allclose(loss.item(), ref['loss'], rtol=1e-5)
    assert ctx.max_over_ranks(float(rank + 1), 'cpu') == float(world)
    ctx.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')


def test_data_parallel_equals_single_process(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f'ok{r}')) for r in range(world))


def test_dp_assign_oracle_partitions_the_global_minibatch():
    """oracle/layouts.dp_assign_minibatch (the checker of mlb_dp_assign_minibatches): the ranks' lists partition
    the global minibatch, every rank holds exactly M ids, a rank never fetches remotely while it still owns
    unassigned ids, and the remote fraction of a random permutation is the binomial imbalance."""
    import numpy as np
    from oracle import layouts
    rng = np.random.default_rng(0)
    for world, B, M, C in ((2, 64, 16, 1), (4, 96, 24, 2), (8, 2048, 512, 1), (8, 32, 8, 3)):
        Jp = C * world * B
        perm = rng.permutation(Jp)
        remote = total = 0
        for k in range(Jp // (world * M)):
            ids = perm[k * world * M:(k + 1) * world * M]
            lists = layouts.dp_assign_minibatch(ids, world, B, M)
            assert all(len(x) == M for x in lists)
            assert np.array_equal(np.sort(np.concatenate(lists)), np.sort(ids))
            owner = (ids % (world * B)) // B
            for r, x in enumerate(lists):
                mine = (x % (world * B)) // B == r
                n_own = int(np.sum(owner == r))
                assert int(mine.sum()) == min(n_own, M)            # keeps everything it owns, up to M
                remote += int((~mine).sum())
                total += M
        if M >= 512:
            assert remote / total < 0.06
