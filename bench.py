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
  roofline  the kernel with the LARGEST time share of the step (picked from `kernels`, the per-kernel
            table: standalone cold-cache launch time x launches per update), CUDA-event timed here
  kernels   per-kernel table: fused forward layer, backward-dx layer, dW GEMM, head GEMM, PPO loss,
            rollout step kernel, optimiser -- us/launch, algorithmic bytes, fraction of the HBM roofline
  f32       (N = 1) the same update with compute_dtype=float32 (the reference's default dtype)
  cfg4      (N = 1) BASELINE configs[3]: recurrent (LSTM) actor-critic, 16384 worlds x 128 steps, BPTT
            minibatches, value-normaliser EMA, bf16 tensor-core path
  cfg3      BASELINE configs[2] (MLP 3x512, 65536 worlds x 64 steps) STRONG scaling: 65536/N worlds
            per rank, index-exact global minibatch permutation, gradient all-reduce per minibatch
  dp_check  (N > 1) the fused all-reduce kernel self-check run before timing
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


CFG3 = dict(name='cfg3: PPO MLP 3x512, 65536 worlds x 64 steps, 4 epochs x 4 minibatches, worlds sharded',
            worlds=65536, steps=64, hidden=512, layers=3, epochs=4, minibatches=4)


CFG4 = dict(name='cfg4: recurrent actor-critic (MLP 2x256 + LSTM 256), 16384 worlds x 128 steps, 4 BPTT chunks, '
                 '2 epochs x 4 minibatches, value-normaliser EMA',
            worlds=16384, steps=128, hidden=256, layers=2, rnn=256, chunks=4, epochs=2, minibatches=4,
            normalize_values=True)


def make_cfg(m, worlds, lr=3e-4, dtype=None, wl=None):
    import torch
    wl = wl or WORKLOAD
    return m.TrainConfig(
        num_worlds=worlds, num_agents_per_world=1, num_updates=1 << 30,
        actions={'act': m.DiscreteActionsConfig(BUCKETS)}, steps_per_update=wl['steps'], lr=lr,
        algo=m.PPOConfig(num_epochs=wl['epochs'],
                         minibatch_size=worlds * wl.get('chunks', 1) // wl['minibatches'], clip_coef=0.2,
                         value_loss_coef=0.5, entropy_coef={'act': 0.01}, max_grad_norm=0.5),
        num_bptt_chunks=wl.get('chunks', 1), gamma=0.99, seed=0, metrics_buffer_size=4, gae_lambda=0.95,
        dreamer_v3_critic=False, normalize_values=bool(wl.get('normalize_values', False)),
        compute_dtype=torch.bfloat16 if dtype == 'bf16' else torch.float32)


def make_policy(m, wl=None):
    wl = wl or WORKLOAD
    net = m.models.MLP(wl['hidden'], wl['layers'])
    enc = m.RecurrentBackboneEncoder(net=net, rnn=m.rnn.LSTM(wl['rnn'], 1)) if wl.get('rnn') else m.BackboneEncoder(net=net)
    return m.Policy(actor_critic=m.ActorCritic(
        backbone=m.BackboneShared(prefix=None, encoder=enc),
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


def unpin_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 for its workers; the CPU arm must use every host core.
    Has to run before NumPy / the BLAS is first imported."""
    cores = os.cpu_count() or 1
    for k in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'NUMEXPR_NUM_THREADS'):
        os.environ[k] = str(cores)
    return cores


def cpu_reference_arm(steps, warmup, worlds):
    """The reference restatement (oracle/) on the host cores; cfg2 at `worlds` worlds."""
    import numpy as np  # noqa: F401
    try:
        import torch
        torch.set_num_threads(os.cpu_count() or 1)
    except Exception:
        pass
    threads = None
    try:
        from threadpoolctl import threadpool_info
        threads = max([p.get('num_threads', 1) for p in threadpool_info()] or [1])
    except Exception:
        pass
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
    return worlds * WORKLOAD['steps'] * steps / dt, dt, threads


def traffic_table():
    """Per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels, from the
    committed ncu --set full capture of this round (profiles/r2_traffic.json names its source CSV)."""
    try:
        return json.load(open(os.path.join(ROOT, 'profiles', 'r2_traffic.json')))
    except Exception:
        return {}


def kernel_table(torch, m, dev, pk, ms_per_update, dtype):
    """Standalone, cold-cache (L2 flushed between launches) device time of every hot kernel of the
    cfg2 learner at its real shapes, with its algorithmic bytes and launches per update.  `share` =
    us x launches / the measured update time (the in-situ shares are in profiles/r2_launches_*.csv)."""
    from madrona_learn_b200 import _lib
    from madrona_learn_b200._lib import c_float, c_int, c_ll, c_size_t, call, ptr
    from madrona_learn_b200.engine import PolicyProgram, gemm_tc, _splitk_tc
    import ctypes
    wl = WORKLOAD
    N, T, H, L, D = wl['worlds'], wl['steps'], wl['hidden'], wl['layers'], wl['obs_dim']
    nmb, E = wl['minibatches'], wl['epochs']
    M = N // nmb
    rows = M * T
    if dtype != 'bf16':
        return []
    prog = PolicyProgram(make_policy(m).actor_critic, D, {'act': m.DiscreteActionsConfig(BUCKETS)}, dev, torch.bfloat16)
    prog.init_params(0)
    w = prog.train_ws(rows)
    g = torch.Generator(device=dev).manual_seed(0)
    obs = torch.randn(rows, D, device=dev, generator=g)
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)       # 256 MiB > 126 MB L2
    A = prog.A
    acts = torch.stack([torch.randint(0, b, (rows,), device=dev, generator=g) for b in BUCKETS], 1).to(torch.int32)
    lp = -torch.rand(rows, A, device=dev, generator=g) - 0.5
    adv = torch.randn(rows, 1, device=dev, generator=g)
    ret = torch.randn(rows, 1, device=dev, generator=g)
    adv_mr = torch.tensor([0.0, 1.0, 1.0, float(rows)], device=dev)
    obj_scale = (ctypes.c_float * A)(*[1.0 / (rows * A)] * A)
    ent_scale = (ctypes.c_float * A)(*[0.01 / (rows * A)] * A)
    head = prog.forward_train(obs, rows)                      # fills y / xhat / rstd with real data
    torch.cuda.synchronize()
    _, s1, b1 = prog.layer_views(prog.params, 1)
    _, gs1, gb1 = prog.layer_views(prog.grads, 1)
    gk1, _, _ = prog.layer_views(prog.grads, 2)
    gW, _ = prog.head_views(prog.grads)
    _, hb = prog.head_views(prog.params)
    dz_a, dz_b = w['dzs'][0], w['dzs'][1]
    dz_a.copy_(torch.randn(rows, H, device=dev, generator=g).to(torch.bfloat16) * 1e-3)

    def k_fwd():
        call('mlb_dense_ln_relu_fwd_tc', ptr(w['y'][0]), ptr(prog.w_t[1]), ptr(s1), ptr(b1), ptr(w['y'][1]),
             ptr(w['xh'][1]), ptr(w['rstd'][1]), c_int(rows), c_int(H), c_int(H), c_int(H), c_int(H))

    def k_dx():
        call('mlb_dense_dx_lnbwd_tc', ptr(dz_a), ptr(prog.w_c[2]), ptr(s1), ptr(b1), ptr(w['xh'][1]),
             ptr(w['rstd'][1]), ptr(dz_b), ptr(gs1), ptr(gb1), c_int(rows), c_int(H), c_int(H), c_int(H), c_int(H))

    def k_dw():
        gemm_tc(w['y'][1], dz_a, gk1, None, H, H, rows, H, H, H, 1, 1, 2, _splitk_tc(H, H, rows))

    def k_head():
        gemm_tc(w['y'][L - 1], prog.wh_t, w['head'], hb, rows, prog.NH, H, H, H, prog.NH, 0, 0, 0)

    def k_dwh():
        gemm_tc(w['y'][L - 1], w['dhead'], gW, None, H, prog.NH, rows, H, prog.NH, prog.NH, 1, 1, 2,
                _splitk_tc(H, prog.NH, rows))

    def k_loss():
        call('mlb_ppo_loss_f32', ptr(head), c_int(prog.NH), ptr(acts), ptr(lp), ptr(adv), ptr(ret), ptr(None),
             ptr(None), ptr(adv_mr), ptr(None), prog._buckets_c, obj_scale, ent_scale, c_int(A), c_ll(rows),
             c_ll(M), c_float(0.2), c_float(0.5), c_int(prog.loss_flags), ptr(w['dhead']),
             ptr(prog.head_bias_grad()), ptr(w['stats_out']), ptr(w['loss_ws']), c_size_t(w['loss_ws'].numel()),
             None, c_int(0))

    def k_opt():
        prog.optimizer_step(3e-4, 0.5)

    robs = torch.randn(N, D, device=dev, generator=g)
    rstore = torch.empty(N, D, device=dev)
    ka, kb = torch.tensor([1, 2], dtype=torch.int32, device=dev), torch.zeros(2, dtype=torch.int32, device=dev)
    ra = torch.empty(N, A, dtype=torch.int32, device=dev)
    rl, rv = torch.empty(N, A, device=dev), torch.empty(N, 1, device=dev)

    def k_roll():
        prog.rollout_step_fused(robs, rstore, N, ka, kb, ra, rl, rv)

    nh = prog.NH
    per_update = E * nmb
    spec = [
        # name, fn, algorithmic bytes per launch, launches per update
        ('fwd_persist_kernel (Dense+LayerNorm+ReLU fwd, 65536x256x256)', k_fwd,
         rows * H * 2 * 3 + rows * 4 + H * H * 2, (L - 1) * per_update),
        ('dx_persist_kernel (dZ W^T + LayerNorm/ReLU bwd, 65536x256x256)', k_dx,
         rows * H * 2 * 3 + rows * 4 + H * H * 2, (L - 1) * per_update),
        ('gemm_tc split-K (dW = X^T dZ, 256x256x65536)', k_dw, rows * H * 2 * 2 + H * H * 4, (L - 1) * per_update),
        ('gemm_bias_persist_kernel (heads fwd, 65536x64x256)', k_head, rows * H * 2 + rows * nh * 4, per_update),
        ('gemm_tc split-K (dW_head, 256x64x65536)', k_dwh, rows * H * 2 + rows * nh * 2, per_update),
        ('ppo_loss_kernel (loss + d_head + metrics)', k_loss, rows * (nh * 4 + nh * 2 + 8 * A + 12), per_update),
        ('optimizer_fused_kernel (clip + Adam + re-projection + bf16 refresh)', k_opt, prog.num_params * 28,
         per_update),
        ('policy_rollout_kernel (key chain + obs store + MLP + heads + sampling, 8192 agents)', k_roll,
         N * (8 * D + 12 * A + 4), T + 1),
    ]
    traffic = traffic_table()
    table = []
    for name, fn, by, launches in spec:
        t = time_kernel(torch, fn, 8, flush=flush)
        short = name.split(' ')[0]
        table.append(dict(kernel=name, us_per_launch=t * 1e6, algorithmic_bytes=by, gbs=by / t / 1e9,
                          frac_hbm=by / t / 1e9 / pk['hbm_gbs'], launches_per_update=launches,
                          share=t * launches / (ms_per_update * 1e-3), traffic=traffic.get(short)))
    return table


def dominant_roofline(torch, m, dev, pk, pk_src, table):
    """`roofline` = the table's largest-share kernel, re-timed back to back over rotating buffer sets
    (> L2 in total, so every launch is cold) -- no per-launch event / host overhead in the bracket."""
    from madrona_learn_b200._lib import c_int, call, ptr
    wl = WORKLOAD
    rows, H = (wl['worlds'] // wl['minibatches']) * wl['steps'], wl['hidden']
    top = max(table, key=lambda r: r['share'])
    BF = torch.bfloat16
    Wm = (torch.randn(H, H, device=dev) * 0.06).to(BF)
    sc, bi = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    fns = []
    is_dx = top['kernel'].startswith('dx_persist')
    if not (is_dx or top['kernel'].startswith('fwd_persist')):
        return dict(bound='hbm', kernel=top['kernel'], achieved=top['gbs'], peak=pk['hbm_gbs'], unit='GB/s',
                    frac=top['frac_hbm'], traffic=top['traffic'], peak_source=pk_src,
                    us_per_launch=top['us_per_launch'], algorithmic_bytes=top['algorithmic_bytes'],
                    share_of_update=top['share'], note='L2 flushed between launches')
    for _ in range(4):               # 4 x 101 MB of operands/results > 126 MB L2
        X = (torch.randn(rows, H, device=dev) * (1e-3 if is_dx else 1.0)).to(BF)
        O1 = torch.empty(rows, H, device=dev, dtype=BF)
        XH = torch.randn(rows, H, device=dev).to(BF)
        rs = torch.rand(rows, device=dev) + 0.5
        ds, db = torch.zeros(H, device=dev), torch.zeros(H, device=dev)
        if is_dx:
            fns.append(lambda X=X, O1=O1, XH=XH, rs=rs, ds=ds, db=db: call(
                'mlb_dense_dx_lnbwd_tc', ptr(X), ptr(Wm), ptr(sc), ptr(bi), ptr(XH), ptr(rs), ptr(O1), ptr(ds),
                ptr(db), c_int(rows), c_int(H), c_int(H), c_int(H), c_int(H)))
        else:
            fns.append(lambda X=X, O1=O1, XH=XH, rs=rs: call(
                'mlb_dense_ln_relu_fwd_tc', ptr(X), ptr(Wm), ptr(sc), ptr(bi), ptr(O1), ptr(XH), ptr(rs),
                c_int(rows), c_int(H), c_int(H), c_int(H), c_int(H)))
    t_k = time_kernel_rotating(torch, fns, 40)
    hbm = top['algorithmic_bytes']
    flops = 2.0 * rows * H * H
    return dict(bound='hbm', kernel=top['kernel'], achieved=hbm / t_k / 1e9, peak=pk['hbm_gbs'], unit='GB/s',
                frac=hbm / t_k / 1e9 / pk['hbm_gbs'], traffic=top['traffic'],
                traffic_source='profiles/r2_traffic.json (ncu --set full, dram read+write per launch)',
                peak_source=pk_src, us_per_launch=t_k * 1e6, algorithmic_bytes=hbm, share_of_update=top['share'],
                tensor_tflops=flops / t_k / 1e12, tensor_frac=flops / t_k / 1e12 / pk['bf16_tflops_sustained'],
                note='largest time share of the update (see `kernels`); arithmetic intensity 85 flop/B < ridge '
                     '213: HBM-bound; 40 back-to-back launches over 4 rotating buffer sets (404 MB > L2)')


def run_arm(torch, m, dev, wl, worlds, dtype, dist_ctx, rank, steps, warmup, sampler=None):
    """Device-resident arm of workload `wl` at `worlds` worlds on this rank -> (seconds, mgr facts)."""
    from madrona_learn_b200 import _lib
    env = m.SyntheticVectorEnv(worlds, WORKLOAD['obs_dim'], len(BUCKETS), seed=rank, device=dev)
    mgr = m.init_training(dev, make_cfg(m, worlds, dtype=dtype, wl=wl), env.sim_fns(), make_policy(m, wl), None,
                          dist_ctx=dist_ctx, verbose=False)
    calls0 = _lib.CALLS
    mgr.update_iter()                                      # eager: counts the enqueue calls
    calls = _lib.CALLS - calls0
    sec = timed_updates(torch, mgr, steps, warmup, dist_ctx, dev, sampler=sampler)
    facts = dict(calls_per_update=calls, graph=mgr._graph is not None,
                 perm_mode=getattr(dist_ctx, 'perm_mode', None) if dist_ctx else None,
                 fused_allreduce=bool(getattr(dist_ctx, 'fused', False)) if dist_ctx else None)
    del mgr, env
    torch.cuda.empty_cache()
    return sec, facts


def dp_self_check(torch, dist_ctx, dev):
    """Fused all-reduce kernel vs the rank-ordered fp32 sum / NCCL, before anything is timed."""
    import torch.distributed as dist
    from madrona_learn_b200.parallel import DistContext

    class P:
        pass
    n = 150_331
    prog = P()
    prog.num_params, prog.device = n, dev
    prog.grads = torch.zeros(n, device=dev)
    prog.grad_sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
    prog.adopt_grad_arena = lambda arena: (arena.zero_(), setattr(prog, 'grads', arena))
    ctx = DistContext()
    if not ctx.enable_fused_allreduce(prog):
        return 'nccl-fallback: ' + str(getattr(ctx, 'fused_error', 'disabled')), ctx
    g = torch.Generator(device=dev).manual_seed(77 + ctx.rank)
    ok = True
    for it in range(3):
        prog.grads.copy_(torch.randn(n, device=dev, generator=g))
        red = ctx.allreduce_grads_fused(prog).clone()
        ref = prog.grads.clone()
        dist.all_reduce(ref)
        ok = ok and bool(torch.allclose(red, ref, rtol=1e-5, atol=1e-5))
        want = (red.double() ** 2).sum().item()
        ok = ok and abs(prog.grad_sumsq.item() - want) <= 1e-10 * want
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return ('ok' if flag.item() == 1.0 else 'MISMATCH'), ctx


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=60)
    ap.add_argument('--warmup', type=int, default=30)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the f32 / cfg3 / kernel-table arms')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'],
                    help='bf16: tcgen05 tensor-core MLP (fp32 accumulate/statistics); f32: fp32 activations, Dense '
                         'products on tcgen05 kind::tf32 (MLB_MATMUL_PRECISION=highest: SIMT FFMA)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        if rank != 0:
            return
        cores = unpin_host_threads()
        sample_worlds = WORKLOAD['worlds']                 # the FULL cfg2 workload, every host thread
        steps = max(1, min(args.steps, 2))
        val, dt, threads = cpu_reference_arm(steps, 1, sample_worlds)
        print(json.dumps({
            'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': steps, 'warmup': 1, 'ms_per_step': dt / steps * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': {'workload': WORKLOAD['name']},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'blas_threads': threads, 'kind': 'port',
                             'sample': f'all {sample_worlds} worlds x {WORKLOAD["steps"]} steps, {steps} timed '
                                       f'update(s) after 1 warm-up; NumPy restatement of the reference (jax not '
                                       f'installable in this image), BLAS threads unpinned'},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    cores = unpin_host_threads() if rank == 0 else (os.cpu_count() or 1)
    import torch
    import madrona_learn_b200 as m
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist_ctx = None
    dp_check = None
    if world > 1:
        import torch.distributed as dist
        from madrona_learn_b200.parallel import DistContext
        dist.init_process_group('nccl', device_id=dev)
        dp_check, _chk = dp_self_check(torch, None, dev)
        del _chk
        dist_ctx = DistContext()
    pk, pk_src = peaks()
    N, T = WORKLOAD['worlds'], WORKLOAD['steps']          # per-GPU worlds (weak scaling)

    # ---- device-resident arm (cfg2, weak scaling) ---------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    sec, facts = run_arm(torch, m, dev, WORKLOAD, N, args.dtype, dist_ctx, rank, args.steps, args.warmup, sampler)
    clocks = sampler.finish() if sampler else None
    value = world * N * T * args.steps / sec

    # ---- end-to-end arm: simulator outputs in pinned host memory ----------------------
    henv = m.HostTraceEnv(N, T, WORKLOAD['obs_dim'], seed=rank, device=dev, num_action_components=len(BUCKETS))
    hmgr = m.init_training(dev, make_cfg(m, N, dtype=args.dtype), henv.sim_fns(), make_policy(m), None,
                           dist_ctx=dist_ctx, verbose=False)
    d2h = [0]

    def readback():
        d2h[0] = hmgr.metrics.ring.numel() + henv.d2h_bytes_per_update
        hmgr.metrics.latest()                              # device->host read of the records

    sec_e2e = timed_updates(torch, hmgr, args.steps, args.warmup, dist_ctx, dev, after_step=readback)
    e2e_val = world * N * T * args.steps / sec_e2e
    h2d = henv.h2d_bytes_per_update
    del hmgr, henv
    torch.cuda.empty_cache()

    # ---- cfg3 strong scaling: 65536 worlds / N per rank ---------------------------------
    cfg3 = None
    if not args.no_extras and args.dtype == 'bf16':
        w3 = CFG3['worlds'] // world
        k3, wu3 = max(3, min(args.steps, 8)), 3
        sec3, f3 = run_arm(torch, m, dev, CFG3, w3, 'bf16', dist_ctx, rank, k3, wu3)
        cfg3 = dict(workload=CFG3['name'], scaling='strong', worlds_total=CFG3['worlds'], worlds_per_gpu=w3,
                    value=CFG3['worlds'] * CFG3['steps'] * k3 / sec3, unit=UNIT, ms_per_step=sec3 / k3 * 1e3,
                    steps=k3, warmup=wu3, dtype='bf16', permutation=f3['perm_mode'] or 'single-gpu',
                    fused_allreduce=f3['fused_allreduce'])

    # ---- the reference's default dtype (compute_dtype=float32), N = 1 only ----------------
    f32 = None
    if not args.no_extras and world == 1 and args.dtype == 'bf16':
        kf = max(3, min(args.steps, 10))
        secf, _ = run_arm(torch, m, dev, WORKLOAD, N, 'f32', None, rank, kf, 3)
        f32 = dict(value=N * T * kf / secf, unit=UNIT, ms_per_step=secf / kf * 1e3, steps=kf, warmup=3,
                   dtype='f32', matmul_precision=m.matmul_precision(),
                   note='compute_dtype=float32 (ml/cfg.py:96 default) on the same workload; Dense products on '
                        'tcgen05.mma.kind::tf32 (mlb_gemm_tf32_tc), XLA:GPU\'s default f32 dot precision; '
                        'MLB_MATMUL_PRECISION=highest selects exact fp32 FFMA')

    # ---- cfg4: recurrent (LSTM) actor-critic with BPTT minibatches, N = 1 only --------------
    cfg4 = None
    if not args.no_extras and world == 1 and args.dtype == 'bf16':
        k4 = max(3, min(args.steps, 6))
        sec4, _ = run_arm(torch, m, dev, CFG4, CFG4['worlds'], 'bf16', None, rank, k4, 3)
        cfg4 = dict(workload=CFG4['name'], value=CFG4['worlds'] * CFG4['steps'] * k4 / sec4, unit=UNIT,
                    ms_per_step=sec4 / k4 * 1e3, steps=k4, warmup=3, dtype='bf16',
                    note='LSTM products on tcgen05: fused [x|h] GEMM + cell epilogue per step (mlb_lstm_step_tc)')

    out = None
    if rank == 0:
        table, roof = [], None
        if not args.no_extras:
            table = kernel_table(torch, m, dev, pk, sec / args.steps * 1e3, args.dtype)
        if table:
            roof = dominant_roofline(torch, m, dev, pk, pk_src, table)
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
                   traffic=traffic_table().get('gae_kernel'), peak_source=pk_src, us_per_launch=t_gae * 1e6)
        del r, v, d, b, adv, ret
        torch.cuda.empty_cache()
        cpu = None
        if not args.no_cpu_baseline:
            sw = 1024
            cval, cdt, threads = cpu_reference_arm(2, 1, sw)
            cpu = {'value': cval, 'unit': UNIT, 'cores': cores, 'blas_threads': threads, 'kind': 'port',
                   'sample': f'{sw} of {N} worlds x {T} steps, 2 updates ({cdt:.1f} s); NumPy restatement '
                             f'of the reference (jax not installable in this image); the full 8192-world run is '
                             f'the --impl reference arm'}
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': sec / args.steps * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': WORKLOAD['name'], 'worlds_per_gpu': N, 'steps_per_update': T,
                       'obs_dim': WORKLOAD['obs_dim'], 'actions': BUCKETS, 'parallelism': f'dp{world}',
                       'cuda_graph': facts['graph'], 'permutation': facts['perm_mode'] or 'single-gpu',
                       'fused_allreduce': facts['fused_allreduce'],
                       'l2_policy': 'per-update working set (~1.2 GB of activations per minibatch) exceeds '
                                    'the 126 MB L2; kernel-only timings flush or exceed L2'},
            'e2e': {'value': e2e_val, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h[0],
                    'ms_per_step': sec_e2e / args.steps * 1e3},
            'gpu_launches': facts['calls_per_update'] * args.steps,
            'gpu_launches_note': 'C-ABI enqueue calls (each >= 1 kernel) per update x steps; replayed from '
                                 'the captured CUDA graph after the first eager update',
            'clocks': clocks, 'roofline': roof, 'kernels': table, 'gae': gae, 'cpu_baseline': cpu,
            'f32': f32, 'cfg3': cfg3, 'cfg4': cfg4, 'dp_check': dp_check,
        }
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
