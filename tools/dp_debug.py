"""Diagnostic: per-minibatch gradients of the R-rank update vs the 1-GPU update on the concatenated rollout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('MLB_CUDA_GRAPH', '0')
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
import dp_update_check as chk


def hook(prog, log):
    orig = prog.optimizer_step

    def wrapped(lr, mgn, gs=1.0, b1=0.9, b2=0.999, eps=1e-8, reduced=None):
        g = (prog.grads if reduced is None else reduced).clone()
        torch.cuda.synchronize()
        log.append((g, prog.grad_sumsq.clone(), prog.params.clone()))
        r = orig(lr, mgn, gs, b1, b2, eps, reduced=reduced)
        torch.cuda.synchronize()
        log[-1] = log[-1] + (prog.grad_sumsq.clone(), prog.params.clone())
        return r
    prog.optimizer_step = wrapped


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', device_id=dev)
    import madrona_learn_b200 as m
    from madrona_learn_b200.parallel import DistContext
    from madrona_learn_b200.ppo import _ppo
    from madrona_learn_b200.rollouts import RolloutData
    kw = dict(N=256, T=16, C=1, M=64, E=2, H=128)
    ctx = DistContext()
    mgr, cfg = chk.make(m, dev, dtype=torch.float32, seed=100 + rank, dist_ctx=ctx, **kw)
    prog = mgr.state.policy_states.program
    log_dp = []
    hook(prog, log_dp)
    mgr.update_iter()
    torch.cuda.synchronize()
    dist.barrier()
    N, T, C, M, E = kw['N'], kw['T'], kw['C'], kw['M'], kw['E']
    Tp, Ng, Mg = T // C, N * world, M * world
    gstore = {}
    for name, x in mgr.rollout_mgr.store.items():
        xl = x.view(torch.uint8) if x.dtype == torch.bool else x
        parts = [torch.empty_like(xl) for _ in range(world)]
        dist.all_gather(parts, xl.contiguous())
        g = torch.cat(parts, dim=3)
        gstore[name] = g.view(torch.bool) if x.dtype == torch.bool else g
    one, cfg1 = chk.make(m, dev, dtype=torch.float32, seed=100, dist_ctx=None, **{**kw, 'N': Ng, 'M': Mg})
    prog1 = one.state.policy_states.program
    for name, g in gstore.items():
        one.rollout_mgr.store[name].copy_(g)
    log_1 = []
    hook(prog1, log_1)
    data = RolloutData(one.rollout_mgr.store, C, Tp, Ng)
    _ppo(cfg1, one.state.policy_states, one.state.train_states, data, lambda mt, *a: mt, one.metrics, ws=one.ppo_ws)
    torch.cuda.synchronize()
    if rank == 0:
        print('mb_adv dp ', mgr.ppo_ws.mb_adv[:3].flatten().tolist())
        print('mb_adv one', one.ppo_ws.mb_adv[:3].flatten().tolist())
        print('obj_scale', list(mgr.ppo_ws.obj_scale)[:2], list(one.ppo_ws.obj_scale)[:2], 'ent', list(mgr.ppo_ws.ent_scale)[:1], list(one.ppo_ws.ent_scale)[:1])
        for i, (a, b) in enumerate(zip(log_dp, log_1)):
            ga, gb = a[0][:prog1.num_params].double(), b[0].double()
            print(f'mb {i}: |g_dp| {ga.norm():.6e} |g_1| {gb.norm():.6e} rel {float((ga - gb).norm() / gb.norm()):.3e} '
                  f'sumsq_dp {float(a[3]):.6e} sumsq_1 {float(b[3]):.6e}  p_before rel {float((a[2].double() - b[2].double()).norm() / b[2].double().norm()):.2e} '
                  f'p_after rel {float((a[4].double() - b[4].double()).norm() / b[4].double().norm()):.2e}')
    dist.destroy_process_group()


main()
