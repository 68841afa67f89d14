"""Multi-GPU parity of one whole data-parallel update (launched by torchrun, one rank per GPU):

  1. index-exact minibatches: the ranks' trajectory-id lists of global minibatch (e, k) partition
     perm[e, k*Mg:(k+1)*Mg] exactly, and the rows every rank gathers (from its own store or over NVLink from
     the owner's) equal BIT FOR BIT the 1-GPU gather of the same trajectories from the concatenated rollout
     store (ml/ppo.py:437-466 on the concatenated data);
  2. the R-rank `update_iter` (worlds sharded, per-minibatch gradient all-reduce through the fused
     NVLink / NVLS kernel, global z-score statistics) equals the 1-GPU update on the concatenated
     rollout: same permutations, parameters rel-L2 <= 1e-5 (fp32 path; only summation order differs).

Prints 'DP_UPDATE_OK {json}' on rank 0.  MLB_DP_DTYPE=bf16 runs the tensor-core path (tolerance 2e-3).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('MLB_CUDA_GRAPH', '0')
import numpy as np
import torch
import torch.distributed as dist

BUCKETS = [4, 8, 5, 5, 2, 2]


def make(m, dev, N, T, C, M, E, H, dtype, seed, dist_ctx, lstm=False, normalize_values=False):
    env = m.SyntheticVectorEnv(N, 32, len(BUCKETS), seed=seed, p_done=1 / 16, device=dev)
    enc = (m.RecurrentBackboneEncoder(net=m.models.MLP(H, 1), rnn=m.rnn.LSTM(H, 1)) if lstm
           else m.BackboneEncoder(net=m.models.MLP(H, 2)))
    policy = m.Policy(actor_critic=m.ActorCritic(
        backbone=m.BackboneShared(prefix=None, encoder=enc),
        actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(BUCKETS)),
        critic=m.models.DenseLayerCritic()))
    cfg = m.TrainConfig(
        num_worlds=N, num_agents_per_world=1, num_updates=10, actions={'act': m.DiscreteActionsConfig(BUCKETS)},
        steps_per_update=T, lr=3e-4,
        algo=m.PPOConfig(num_epochs=E, minibatch_size=M, clip_coef=0.2, value_loss_coef=0.5,
                         entropy_coef={'act': 0.01}, max_grad_norm=0.5),
        num_bptt_chunks=C, gamma=0.99, seed=5, metrics_buffer_size=4, gae_lambda=0.95,
        dreamer_v3_critic=False, normalize_values=normalize_values, compute_dtype=dtype)
    return m.init_training(dev, cfg, env.sim_fns(), policy, None, dist_ctx=dist_ctx, verbose=False), cfg


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    import madrona_learn_b200 as m
    from madrona_learn_b200 import kernels as K
    from madrona_learn_b200.parallel import DistContext
    from madrona_learn_b200.ppo import _ppo
    from madrona_learn_b200.rollouts import RolloutData
    dtype = torch.bfloat16 if os.environ.get('MLB_DP_DTYPE') == 'bf16' else torch.float32
    report = {}
    for case, kw in (('mlp', dict(N=256, T=16, C=1, M=64, E=2, H=128)),
                     ('lstm_chunks_vn', dict(N=64, T=16, C=2, M=32, E=1, H=64, lstm=True, normalize_values=True))):
        if kw.get('lstm') and dtype != torch.float32:
            continue
        ctx = DistContext()
        mgr, cfg = make(m, dev, dtype=dtype, seed=100 + rank, dist_ctx=ctx, **kw)
        assert ctx.perm_mode == 'index_exact', getattr(ctx, 'symm_store_error', 'index-exact mode unavailable')
        prog = mgr.state.policy_states.program
        p0 = prog.params.clone()
        key0 = mgr.state.train_states.update_prng_key.clone()
        vn0 = (mgr.state.train_states.value_normalizer_state.clone()
               if mgr.state.train_states.value_normalizer is not None else None)
        mgr.update_iter()
        torch.cuda.synchronize()
        dist.barrier()
        N, T, C, M, E = kw['N'], kw['T'], kw['C'], kw['M'], kw['E']
        Tp, Ng, Mg = T // C, kw['N'] * world, kw['M'] * world
        # ---- the concatenated rollout store (worlds of rank r at [r*N, (r+1)*N)) -------------------
        gstore = {}
        for name, x in mgr.rollout_mgr.store.items():
            xl = x.view(torch.uint8) if x.dtype == torch.bool else x
            parts = [torch.empty_like(xl) for _ in range(world)]
            dist.all_gather(parts, xl.contiguous())
            bdim = 2 if name.startswith('rnn_start') else 3
            g = torch.cat(parts, dim=bdim)
            gstore[name] = g.view(torch.bool) if x.dtype == torch.bool else g
        # ---- (1) bit-exact minibatches -------------------------------------------------------------
        ws = mgr.ppo_ws
        perm = ws.perm.clone()
        perms_all = [torch.empty_like(perm) for _ in range(world)]
        dist.all_gather(perms_all, perm)
        assert all(torch.equal(perms_all[0], q) for q in perms_all), 'ranks disagree on the global permutation'
        assert perm.shape == (E, C * Ng)
        names = [k for k in ('obs', 'actions', 'log_probs', 'advantages', 'returns') if k in ws.mb]
        checked = 0
        for (e, k) in ((0, 0), (E - 1, C * Ng // Mg - 1)):
            lo = k * Mg + rank * M
            idx = (ws.idx_local[e, k] if ws.idx_local is not None else perm[e, lo:lo + M]).contiguous()
            # the ranks' id lists partition the global minibatch perm[e, k*Mg:(k+1)*Mg] exactly
            ids_all = [torch.empty_like(idx) for _ in range(world)]
            dist.all_gather(ids_all, idx)
            got_ids = torch.cat(ids_all)
            ref_ids = perm[e, k * Mg:(k + 1) * Mg]
            assert torch.equal(got_ids.sort().values, ref_ids.sort().values), (case, e, k, 'id multiset differs')
            owners = (idx % (world * N)) // N
            remote_frac = float((owners != rank).float().mean())
            order = got_ids.sort().indices                         # position of every global id in the R-rank concat
            ref_order = ref_ids.sort().indices
            leaves = [(mgr.rollout_mgr.store[n], ws.mb[n], None) for n in names]
            K.mb_gather_multi_peer(leaves, ctx.peer_store_table(names), world, idx, C, Tp, N)
            torch.cuda.synchronize()
            for n in names:
                mine = ws.mb[n].clone()
                parts = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(parts, mine)
                got = torch.cat(parts, dim=1)                                  # [T', Mg, *] in R-rank order
                ref = K.mb_gather(gstore[n][:, :, 0].contiguous(), ref_ids.contiguous(), C, Tp, Ng)
                # same trajectory -> same bytes, whichever rank trained on it
                assert torch.equal(got[:, order].contiguous().view(torch.uint8), ref[:, ref_order].contiguous().view(torch.uint8)), (case, n, e, k)
                checked += 1
        # ---- (2) the update == the 1-GPU update on the concatenated rollout ---------------------------
        one, cfg1 = make(m, dev, dtype=dtype, seed=100, dist_ctx=None,
                         **{**kw, 'N': Ng, 'M': Mg})
        prog1 = one.state.policy_states.program
        assert torch.equal(prog1.params, p0), 'initial parameters differ between the R-rank and the 1-GPU run'
        assert torch.equal(one.state.train_states.update_prng_key, key0)
        if vn0 is not None:
            one.state.train_states.value_normalizer_state.copy_(vn0)
        for name, g in gstore.items():
            one.rollout_mgr.store[name].copy_(g)
        data = RolloutData(one.rollout_mgr.store, C, Tp, Ng)
        _ppo(cfg1, one.state.policy_states, one.state.train_states, data, lambda mt, *a: mt, one.metrics, ws=one.ppo_ws)
        torch.cuda.synchronize()
        assert torch.equal(one.ppo_ws.perm, perm), 'permutations differ from the 1-GPU run'
        a, b = prog.params.double(), prog1.params.double()
        rel = float((a - b).norm() / b.norm())
        rel_delta = float((a - b).norm() / (b - p0.double()).norm())
        tol = 1e-5 if dtype == torch.float32 else 2e-3
        # every rank holds the same parameters (the reduced gradient is identical everywhere)
        allp = [torch.empty_like(prog.params) for _ in range(world)]
        dist.all_gather(allp, prog.params)
        assert all(torch.equal(allp[0], q) for q in allp), 'ranks hold different parameters after the update'
        assert rel <= tol, (case, rel, rel_delta)
        if vn0 is not None:
            va = mgr.state.train_states.value_normalizer_state[:5]
            vb = one.state.train_states.value_normalizer_state[:5]
            assert torch.allclose(va, vb, rtol=1e-5, atol=1e-7), (va, vb)
        report[case] = dict(world=world, fused_allreduce=bool(ctx.fused), nvls=bool(getattr(ctx, 'nvls', False)),
                            assign=ws.assign, remote_row_fraction=remote_frac,
                            minibatch_leaves_bit_exact=checked, params_rel_l2=rel, delta_rel_l2=rel_delta)
        del mgr, one
        dist.barrier()
    if rank == 0:
        print('DP_UPDATE_OK ' + json.dumps(report), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
