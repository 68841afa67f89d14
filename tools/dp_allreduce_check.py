"""Multi-GPU check of mlb_allreduce_sumsq_f32 (launched by torchrun, one rank per GPU):
the fused NVLink all-reduce must equal the rank-ordered fp32 sum BIT FOR BIT on every rank, equal
NCCL's all-reduce to fp32 rounding, report sum(g^2) to 1e-12 relative, and survive CUDA-graph
replay.  Prints 'DP_CHECK_OK' on rank 0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


class _Prog:                     # the slice of PolicyProgram the DistContext touches
    def __init__(self, n, dev):
        self.num_params, self.device = n, dev
        self.grads = torch.zeros(n, device=dev)
        self.grad_sumsq = torch.zeros(1, dtype=torch.float64, device=dev)

    def adopt_grad_arena(self, arena):
        arena.zero_()
        self.grads = arena


def main():
    rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', device_id=dev)
    import madrona_learn_b200  # noqa: F401  (loads the library)
    from madrona_learn_b200.parallel import DistContext
    keep = []                     # symmetric allocations must not be freed inside a graph capture
    for n in (150_331, 4096):
        ctx = DistContext()
        prog = _Prog(n, dev)
        keep.append((ctx, prog))
        assert ctx.enable_fused_allreduce(prog), getattr(ctx, 'fused_error', 'disabled')
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        for it in range(4):
            prog.grads.copy_(torch.randn(n, device=dev, generator=g) * (1 + it))
            red = ctx.allreduce_grads_fused(prog).clone()
            ssq = prog.grad_sumsq.clone()
            # reference: gather everyone's shard, sum in rank order in fp32
            shards = [torch.empty(n, device=dev) for _ in range(world)]
            dist.all_gather(shards, prog.grads.clone())
            ref = shards[0].clone()
            for r in range(1, world):
                ref += shards[r]
            if ctx.nvls:      # the switch's summation order is its own: identical on all ranks, fp32-close to ours
                allr = [torch.empty(n, device=dev) for _ in range(world)]
                dist.all_gather(allr, red)
                assert all(torch.equal(allr[0], a) for a in allr[1:]), 'ranks disagree'
                assert torch.allclose(red, ref, rtol=1e-5, atol=1e-5)
            else:
                assert torch.equal(red, ref), (rank, n, it, (red - ref).abs().max().item())
            nccl = prog.grads.clone()
            dist.all_reduce(nccl)
            assert torch.allclose(red, nccl, rtol=1e-5, atol=1e-5)
            want = (red.double() ** 2).sum()
            assert abs(ssq.item() - want.item()) <= 1e-12 * want.item(), (ssq.item(), want.item())
        # CUDA-graph replay: the epoch lives on the device
        prog.grads.copy_(torch.randn(n, device=dev, generator=g))
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            ctx.allreduce_grads_fused(prog)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                out = ctx.allreduce_grads_fused(prog)
            for _ in range(5):
                gr.replay()
            torch.cuda.synchronize()
        shards = [torch.empty(n, device=dev) for _ in range(world)]
        dist.all_gather(shards, prog.grads.clone())
        ref = shards[0].clone()
        for r in range(1, world):
            ref += shards[r]
        assert torch.equal(out, ref) if not ctx.nvls else torch.allclose(out, ref, rtol=1e-5, atol=1e-5)
        # timing vs NCCL (device time, back-to-back)
        if n > 100_000:
            def timeit(fn, reps=200):
                for _ in range(20):
                    fn()
                torch.cuda.synchronize(); dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record(); torch.cuda.synchronize()
                return e0.elapsed_time(e1) * 1e3 / reps
            buf = prog.grads.clone()
            t_f = timeit(lambda: ctx.allreduce_grads_fused(prog))
            t_n = timeit(lambda: dist.all_reduce(buf))
            if rank == 0:
                print(f'world={world} n={n} nvls={ctx.nvls}: fused all-reduce+sumsq {t_f:.1f} us/call, NCCL all-reduce {t_n:.1f} us/call',
                      flush=True)
    dist.barrier()
    if rank == 0:
        print('DP_CHECK_OK', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
