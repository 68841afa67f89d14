#!/usr/bin/env python
"""bench.py -- agent-steps/s of one full update_iter (rollout + GAE + PPO update), and the GAE
kernel's HBM roofline, on N B200s of one node.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE
JSON line; for N > 1 it is launched under torch.distributed.run (one rank per GPU, NCCL).
  value     whole-job agent-steps/s with the synthetic env resident on the device
            (BASELINE.json configs[1]: PPO MLP 3x256, 8192 worlds x 32 steps, 4 epochs x 4 mb)
  e2e       the same metric through the public API (init_training -> update_iter) with the
            simulator's outputs in PINNED HOST memory: every update copies T*N*(4D+5) bytes
            host->device and reads the metrics record back device->host
  roofline  the dominant kernel of the step (the Dense GEMM), CUDA-event timed in this run
  gae       the GAE kernel's HBM roofline on the BASELINE configs[4] sweep point T=256, N=1M
  cpu_baseline  oracle/ (NumPy restatement of the reference; jax is not installable) timed on
            this box's host cores on a bounded sample of the same workload
`--impl reference` times that CPU restatement as its own arm.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BUCKETS = [4, 8, 5, 5, 2, 2]
WORKLOAD = dict(name='cfg2: PPO MLP 3x256, 8192 worlds x 32 steps, 4 epochs x 4 minibatches',
                worlds=8192, steps=32, obs_dim=64, hidden=256, layers=3, epochs=4, minibatches=4)
METRIC = 'agent-steps/s (rollout+GAE+PPO update)'
UNIT = 'agent-steps/s'


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))), 'measured'
    except Exception:
        return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0), 'fallback'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag, self.proc = gpu, [], False, None
        self.t0, self.t1 = None, None          # the timed window (host clock)

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.gpu}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([time.time()] + [x.strip() for x in line.split(',')])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        rows = [r[1:] for r in self.rows if self.t0 is None or self.t0 <= r[0] <= (self.t1 or 1e30)]
        if not rows:                            # window shorter than one sample: take all
            rows = [r[1:] for r in self.rows]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                    'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx or None,
                    reasons=sorted(reasons), samples=len(sm))


def make_cfg(m, worlds, lr=3e-4, dtype=None):
    import torch
    return m.TrainConfig(
        num_worlds=worlds, num_agents_per_world=1, num_updates=1 << 30,
        actions={'act': m.DiscreteActionsConfig(BUCKETS)}, steps_per_update=WORKLOAD['steps'], lr=lr,
        algo=m.PPOConfig(num_epochs=WORKLOAD['epochs'],
                         minibatch_size=worlds // WORKLOAD['minibatches'], clip_coef=0.2,
                         value_loss_coef=0.5, entropy_coef={'act': 0.01}, max_grad_norm=0.5),
        num_bptt_chunks=1, gamma=0.99, seed=0, metrics_buffer_size=4, gae_lambda=0.95,
        dreamer_v3_critic=False, normalize_values=False,
        compute_dtype=torch.bfloat16 if dtype == 'bf16' else torch.float32)


def make_policy(m):
    return m.Policy(actor_critic=m.ActorCritic(
        backbone=m.BackboneShared(prefix=None, encoder=m.BackboneEncoder(
            net=m.models.MLP(WORKLOAD['hidden'], WORKLOAD['layers']))),
        actor=m.models.DenseLayerDiscreteActor(m.DiscreteActionsConfig(BUCKETS)),
        critic=m.models.DenseLayerCritic()))


def timed_updates(torch, mgr, steps, warmup, dist_ctx, dev, after_step=None, sampler=None):
    """W untimed + exactly K timed update_iters, barrier + synchronize on both sides, CUDA
    events on the launching stream, max over ranks.  Returns seconds."""
    if sampler:
        sampler.start()                        # nvidia-smi needs ~1 s to come up: start early,
    for _ in range(warmup):                    # keep only samples inside the timed window
        mgr.update_iter()
        if after_step:
            after_step()
    torch.cuda.synchronize()
    if dist_ctx:
        dist_ctx.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        mgr.update_iter()
        if after_step:
            after_step()
    e1.record()
    torch.cuda.synchronize()
    if dist_ctx:
        dist_ctx.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.t1 = time.time()
    sec = e0.elapsed_time(e1) * 1e-3
    if dist_ctx:
        sec = dist_ctx.max_over_ranks(sec, dev)
    return sec


def time_kernel(torch, fn, reps, flush=None):
    """Average device time of `fn` (one launch) over `reps`, GPU kept busy ahead of e0 so no
    host launch gap is inside the bracket; optional L2 flush between launches."""
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.add_(1)
        torch.cuda.synchronize()
        torch.cuda._sleep(300000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return statistics.mean(ts)


def time_kernel_rotating(torch, fns, reps):
    """Average device time per launch of `reps` back-to-back launches cycling through `fns`, which
    work on disjoint buffer sets whose total size exceeds L2 (every launch sees cold inputs) --
    one event bracket around all of them, so no per-launch host/event overhead is inside."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    torch.cuda._sleep(300000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def cpu_reference_arm(steps, warmup, worlds):
    """The reference restatement (oracle/) on the host cores; bounded sample of cfg2."""
    import numpy as np  # noqa: F401
    from oracle import ppo as oppo
    from oracle import train as otrain
    cfg = oppo.PPOCfg(BUCKETS, num_epochs=WORKLOAD['epochs'], minibatch_size=worlds // WORKLOAD['minibatches'])
    tr = otrain.OracleTrainer(worlds, WORKLOAD['steps'], WORKLOAD['obs_dim'], WORKLOAD['hidden'],
                              WORKLOAD['layers'], BUCKETS, cfg)
    for _ in range(warmup):
        tr.update_iter()
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.update_iter()
    dt = time.perf_counter() - t0
    return worlds * WORKLOAD['steps'] * steps / dt, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=60)
    ap.add_argument('--warmup', type=int, default=30)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'],
                    help='bf16: tcgen05 tensor-core MLP (fp32 accumulate/statistics); f32: SIMT fp32 MLP')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    cores = os.cpu_count() or 1

    if args.impl == 'reference':
        if rank != 0:
            return
        sample_worlds = 512
        steps = min(args.steps, 3)
        val, dt = cpu_reference_arm(steps, min(args.warmup, 1), sample_worlds)
        print(json.dumps({
            'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': steps, 'warmup': min(args.warmup, 1), 'ms_per_step': dt / steps * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': {'workload': WORKLOAD['name']},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': f'{sample_worlds} of {WORKLOAD["worlds"]} worlds x {WORKLOAD["steps"]} '
                                       f'steps, same model/epochs/minibatch count; NumPy restatement of the '
                                       f'reference (jax not installable in this image)'},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    import torch
    import madrona_learn_b200 as m
    from madrona_learn_b200 import _lib
    from madrona_learn_b200.engine import gemm
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist_ctx = None
    if world > 1:
        import torch.distributed as dist
        from madrona_learn_b200.parallel import DistContext
        dist.init_process_group('nccl', device_id=dev)
        dist_ctx = DistContext()
    pk, pk_src = peaks()
    N, T = WORKLOAD['worlds'], WORKLOAD['steps']          # per-GPU worlds (weak scaling)

    # ---- device-resident arm ---------------------------------------------------------
    env = m.SyntheticVectorEnv(N, WORKLOAD['obs_dim'], len(BUCKETS), seed=rank, device=dev)
    mgr = m.init_training(dev, make_cfg(m, N, dtype=args.dtype), env.sim_fns(), make_policy(m), None, dist_ctx=dist_ctx,
                          verbose=False)
    calls0 = _lib.CALLS
    mgr.update_iter()                                      # eager: counts the enqueue calls
    calls_per_update = _lib.CALLS - calls0
    sampler = ClockSampler(local_rank) if rank == 0 else None
    sec = timed_updates(torch, mgr, args.steps, args.warmup, dist_ctx, dev, sampler=sampler)
    clocks = sampler.finish() if sampler else None
    value = world * N * T * args.steps / sec
    graph_on = mgr._graph is not None
    del mgr, env
    torch.cuda.empty_cache()

    # ---- end-to-end arm: simulator outputs in pinned host memory ----------------------
    henv = m.HostTraceEnv(N, T, WORKLOAD['obs_dim'], seed=rank, device=dev)
    hmgr = m.init_training(dev, make_cfg(m, N, dtype=args.dtype), henv.sim_fns(), make_policy(m), None, dist_ctx=dist_ctx,
                           verbose=False)
    d2h = [0]

    def readback():
        d2h[0] = hmgr.metrics.ring.numel()
        hmgr.metrics.latest()                              # device->host read of the records

    sec_e2e = timed_updates(torch, hmgr, args.steps, args.warmup, dist_ctx, dev, after_step=readback)
    e2e_val = world * N * T * args.steps / sec_e2e
    h2d = henv.h2d_bytes_per_update
    del hmgr, henv
    torch.cuda.empty_cache()

    out = None
    if rank == 0:
        # ---- dominant kernel (Dense GEMM of one minibatch) roofline --------------------
        rows, H = (N // WORKLOAD['minibatches']) * T, WORKLOAD['hidden']
        flops = 2.0 * rows * H * H
        if args.dtype == 'bf16':
            from madrona_learn_b200._lib import c_int, call, ptr
            BF = torch.bfloat16
            Wt = (torch.randn(H, H, device=dev) * 0.06).to(BF)
            sc, bi = torch.ones(H, device=dev), torch.zeros(H, device=dev)
            sets = []
            for _ in range(4):          # 4 x 101 MB of operands/results > 126 MB L2: every launch is cold
                sets.append((torch.randn(rows, H, device=dev).to(BF), torch.empty(rows, H, device=dev, dtype=BF),
                             torch.empty(rows, H, device=dev, dtype=BF), torch.empty(rows, device=dev)))

            def mk(X, Y, XH, rs):
                return lambda: call('mlb_dense_ln_relu_fwd_tc', ptr(X), ptr(Wt), ptr(sc), ptr(bi), ptr(Y), ptr(XH),
                                    ptr(rs), c_int(rows), c_int(H), c_int(H), c_int(H), c_int(H))
            t_k = time_kernel_rotating(torch, [mk(*st) for st in sets], 40)
            # algorithmic HBM bytes of one launch: X in (bf16) + Y and xhat out (bf16) + rstd + W once
            hbm = rows * H * 2 * 3 + rows * 4 + H * H * 2
            roof = dict(bound='hbm',
                        kernel='fwd_persist_kernel (persistent tcgen05 Dense + LayerNorm + ReLU, training variant, '
                               f'{rows} x {H} x {H}, bf16 in/out, fp32 TMEM accumulate, W resident in smem)',
                        achieved=hbm / t_k / 1e9, peak=pk['hbm_gbs'], unit='GB/s',
                        frac=hbm / t_k / 1e9 / pk['hbm_gbs'], traffic=None, peak_source=pk_src,
                        us_per_launch=t_k * 1e6, algorithmic_bytes=hbm,
                        tensor_tflops=flops / t_k / 1e12,
                        tensor_frac=flops / t_k / 1e12 / pk['bf16_tflops_sustained'],
                        note='arithmetic intensity 85 flop/B < ridge (1395 TF / 6.5 TB/s = 213): the fused layer '
                             'is HBM-bound; 40 back-to-back launches over 4 rotating buffer sets (404 MB > L2)')
            del sets, Wt
        else:
            A = torch.randn(rows, H, device=dev)
            B = torch.randn(H, H, device=dev)
            C = torch.empty(rows, H, device=dev)
            t_k = time_kernel(torch, lambda: gemm(A, B, C, None, rows, H, H, H, H, H), 10)
            tf = flops / t_k / 1e12
            roof = dict(bound='tensor', kernel='sgemm_kernel<128,128,8,8> (fp32 SIMT Dense forward, 65536 x 256 x 256)',
                        achieved=tf, peak=pk['bf16_tflops_sustained'], unit='TFLOP/s',
                        frac=tf / pk['bf16_tflops_sustained'], traffic=None, peak_source=pk_src,
                        us_per_launch=t_k * 1e6, note='fp32 FFMA path (compute_dtype=float32)')
            del A, B, C
        # ---- GAE kernel HBM roofline (configs[4] sweep point, > L2) --------------------
        K = m.kernels
        Tg, Ng = 256, 1 << 20
        r = torch.randn(Tg, Ng, device=dev)
        v = torch.randn(Tg, Ng, device=dev)
        d = torch.rand(Tg, Ng, device=dev) < 0.02
        b = torch.randn(Ng, device=dev)
        adv, ret = torch.empty_like(r), torch.empty_like(r)
        t_gae = time_kernel(torch, lambda: K.gae(r, v, d, b, 0.99, 0.95, advantages=adv, returns=ret), 10)
        by = 17.0 * Tg * Ng + 4.0 * Ng
        gae = dict(bound='hbm', kernel='gae_kernel<4,4,false>', workload=f'T={Tg}, N={Ng} (inputs 4.6 GB > L2)',
                   achieved=by / t_gae / 1e9, peak=pk['hbm_gbs'], unit='GB/s',
                   frac=by / t_gae / 1e9 / pk['hbm_gbs'], algorithmic_bytes=by,
                   traffic=4.5228e9, traffic_source='profiles/r1_ncu_full_gae_raw.csv (dram read+write per launch)',
                   peak_source=pk_src, us_per_launch=t_gae * 1e6)
        del r, v, d, b, adv, ret
        torch.cuda.empty_cache()
        cpu = None
        if not args.no_cpu_baseline:
            sw = 512
            cval, cdt = cpu_reference_arm(2, 1, sw)
            cpu = {'value': cval, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                   'sample': f'{sw} of {N} worlds x {T} steps, 2 updates ({cdt:.1f} s); NumPy restatement '
                             f'of the reference (jax not installable in this image)'}
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': sec / args.steps * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': WORKLOAD['name'], 'worlds_per_gpu': N, 'steps_per_update': T,
                       'obs_dim': WORKLOAD['obs_dim'], 'actions': BUCKETS, 'parallelism': f'dp{world}',
                       'cuda_graph': graph_on,
                       'l2_policy': 'per-update working set (~1.2 GB of activations per minibatch) exceeds '
                                    'the 126 MB L2; kernel-only timings flush or exceed L2'},
            'e2e': {'value': e2e_val, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h[0],
                    'ms_per_step': sec_e2e / args.steps * 1e3},
            'gpu_launches': calls_per_update * args.steps,
            'gpu_launches_note': 'C-ABI enqueue calls (each >= 1 kernel) per update x steps; replayed from '
                                 'the captured CUDA graph after the first eager update',
            'clocks': clocks, 'roofline': roof, 'gae': gae, 'cpu_baseline': cpu,
        }
        print(json.dumps(out))
    if dist_ctx:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
