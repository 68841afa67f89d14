"""Per-kernel device-time breakdown of one data-parallel cfg2 update (torchrun, rank 0 prints)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['MLB_CUDA_GRAPH'] = '0'
os.environ['MLB_PDL'] = '0'
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile
import bench
import madrona_learn_b200 as m
from madrona_learn_b200.parallel import DistContext

rank = int(os.environ['RANK'])
torch.cuda.set_device(rank)
dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
ctx = DistContext()
N = bench.WORKLOAD['worlds']
env = m.SyntheticVectorEnv(N, bench.WORKLOAD['obs_dim'], len(bench.BUCKETS), seed=rank, device=dev)
mgr = m.init_training(dev, bench.make_cfg(m, N, dtype='bf16'), env.sim_fns(), bench.make_policy(m), None, dist_ctx=ctx,
                      verbose=False)
for _ in range(2):
    mgr.update_iter()
torch.cuda.synchronize(); dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    mgr.update_iter()
    torch.cuda.synchronize()
if rank == 0:
    rows = []
    for e in prof.key_averages():
        t = getattr(e, 'device_time_total', None) or 0
        if t:
            rows.append((t, e.count, e.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f'perm={ctx.perm_mode} kernel time {tot / 1e3:.2f} ms')
    for t, c, k in rows[:16]:
        print(f'{t / 1e3:8.3f} ms {100 * t / tot:5.1f}% x{c:4d} {t / c:8.1f} us  {k[:90]}')
dist.barrier()
dist.destroy_process_group()
