"""PPO algorithm plugin (ml/ppo.py:24-488): PPOConfig / PPOHyperParams / PPO(AlgoBase).

`_ppo` is the reference's epoch x minibatch loop (default branch, :437-488) restructured for
the B200: all `num_epochs` permutations are generated up-front by the
device threefry + sort kernels (the key stream only depends on update_prng_key); the
per-minibatch z-score statistics of the advantages and the value-normaliser recurrence are
computed for every minibatch in three launches right after GAE; the minibatch loop itself
is gather -> fwd -> fused loss/grad -> bwd -> [grad all-reduce] -> fused optimiser.
"""
import ctypes
import os
from dataclasses import dataclass
from typing import Any, Callable, Dict, Union

import torch

from . import _lib
from . import kernels as K
from ._lib import c_float, c_int, c_ll, c_size_t, call, ptr
from .algo_common import AlgoBase, HyperParams
from .cfg import AlgoConfig, ParamExplore, TrainConfig
from .profile import profile

__all__ = ['PPOConfig']


@dataclass(frozen=True)
class PPOConfig(AlgoConfig):                    # ml/ppo.py:24-39
    num_epochs: int
    minibatch_size: int                         # in TRAJECTORIES (ml/ppo.py:437-439)
    clip_coef: float
    value_loss_coef: float
    entropy_coef: Union[float, Dict[str, float], ParamExplore]
    max_grad_norm: float
    clip_value_loss: bool = False
    huber_value_loss: bool = False

    def name(self):
        return 'ppo'

    def setup(self):
        return PPO()


@dataclass
class PPOHyperParams(HyperParams):              # ml/ppo.py:42-46
    clip_coef: float = 0.0
    value_loss_coef: float = 0.0
    entropy_coef: Any = 0.0
    max_grad_norm: float = 0.0


METRIC_NAMES = ['Loss', 'Action Obj', 'Value Loss', 'Value Errors', 'Entropy']


class PPO(AlgoBase):
    def init_hyperparams(self, cfg: TrainConfig):            # ml/ppo.py:50-81
        if cfg.dreamer_v3_critic or cfg.hlgauss_critic:
            assert not cfg.algo.clip_value_loss
            assert not cfg.algo.huber_value_loss
            assert not cfg.normalize_values
        lr = cfg.lr.base if isinstance(cfg.lr, ParamExplore) else cfg.lr
        ent = cfg.algo.entropy_coef.base if isinstance(cfg.algo.entropy_coef, ParamExplore) \
            else cfg.algo.entropy_coef
        return PPOHyperParams(
            lr=lr, gamma=cfg.gamma, gae_lambda=cfg.gae_lambda, normalize_values=cfg.normalize_values,
            value_normalizer_decay=cfg.value_normalizer_decay,
            max_advantage_est_decay=cfg.max_advantage_est_decay, clip_coef=cfg.algo.clip_coef,
            value_loss_coef=cfg.algo.value_loss_coef, entropy_coef=ent,
            max_grad_norm=cfg.algo.max_grad_norm)

    def make_optimizer(self, hyper_params):                  # ml/ppo.py:83-90
        """optax.chain(clip_by_global_norm, adam): lr is baked here, exactly like the
        reference (the PBT-mutable hyper_params.lr array is NOT what the optimiser reads)."""
        return dict(kind='clip_by_global_norm+adam', max_grad_norm=float(hyper_params.max_grad_norm),
                    lr=float(hyper_params.lr), b1=0.9, b2=0.999, eps=1e-8)

    def update(self, *args, **kwargs):
        return _ppo(*args, **kwargs)

    def add_metrics(self, cfg, metrics):                     # ml/ppo.py:95-106
        return metrics + METRIC_NAMES


def _entropy_coef(cfg, group):
    ec = cfg.algo.entropy_coef
    if isinstance(ec, dict):                                 # ml/ppo.py:233 indexes by group name
        return float(ec[group])
    if isinstance(ec, ParamExplore):
        return float(ec.base)
    return float(ec)


class _PPOWorkspace:
    """Per-run device buffers of the learner (allocated once)."""

    def __init__(self, cfg, prog, C, Tp, B, dist_ctx):
        dev = prog.device
        E, M = cfg.algo.num_epochs, cfg.algo.minibatch_size
        self.mode = ('filter' if cfg.filter_advantages else
                     'importance' if cfg.importance_sample_trajectories else 'default')
        self.store_dims = (C, Tp, B)
        if self.mode != 'default' and dist_ctx is not None:
            raise NotImplementedError('filter_advantages / importance_sample_trajectories are single-GPU')
        if self.mode == 'filter':
            # rollout_data.flatten_time() (ml/ppo.py:375, ml/rollouts.py:331-334): every step becomes
            # its own length-1 trajectory; minibatch_size then counts steps
            if prog.lstm is not None:
                raise NotImplementedError('filter_advantages flattens time: feed-forward policies only')
            C, Tp, B = 1, 1, C * Tp * B
        J = C * B
        if self.mode == 'default':
            assert J % M == 0, 'num trajectories must be divisible by minibatch_size (ml/ppo.py:439)'
        else:
            assert J >= M
        self.E, self.M, self.J, self.nmb, self.rows = E, M, J, J // M, Tp * M
        self.B = B                                       # trajectories per BPTT chunk on this rank
        e = lambda *s, dtype=torch.float32: torch.empty(*s, dtype=dtype, device=dev)
        # index-exact data-parallel mode (parallel.py): ONE permutation of the global trajectory ids,
        # identical on every rank; this rank trains on its M-wide slice of every (world*M)-wide minibatch
        self.index_exact = dist_ctx is not None and getattr(dist_ctx, 'perm_mode', 'fast') == 'index_exact'
        R = dist_ctx.world_size if self.index_exact else 1
        self.Jp = J * R                                  # length of one permutation
        self.Mp = M * R                                  # minibatch width in the permutation
        self.perm = e(E, self.Jp, dtype=torch.int32)
        # index-exact mode: this rank's ids of every global minibatch.  'owner' (default): every rank keeps the
        # trajectories it owns, only the binomial imbalance is fetched from peers (mlb_dp_assign_minibatches);
        # 'slice' (MLB_DP_ASSIGN=slice): the rank-th contiguous M-slice of the permutation.
        self.assign = os.environ.get('MLB_DP_ASSIGN', 'owner') if self.index_exact else None
        self.idx_local = e(E, J // M, M, dtype=torch.int32) if self.assign == 'owner' else None
        self.perm_ws = torch.empty(K.lib().mlb_ppo_permutations_workspace(E, self.Jp) + 16,
                                   dtype=torch.uint8, device=dev)
        self.tm_adv = e(J, 2, dtype=torch.float64)
        self.tm_ret = e(J, 2, dtype=torch.float64)
        if self.index_exact:
            self.tm_adv_g = e(self.Jp, 2, dtype=torch.float64)
            self.tm_ret_g = e(self.Jp, 2, dtype=torch.float64)
        self.mb_adv = e(E * self.nmb, 4)
        self.mb_ret = e(E * self.nmb, 4)
        self.vn_params = e(E * self.nmb, 4)
        self.raw = torch.zeros(2 * E * self.nmb, 2, dtype=torch.float64, device=dev)
        self.mb = {
            'obs': e(Tp, M, prog.obs_dim), 'actions': e(Tp, M, prog.A, dtype=torch.int32),
            'log_probs': e(Tp, M, prog.A), 'advantages': e(Tp, M, 1), 'returns': e(Tp, M, 1),
            'values': e(Tp, M, 1), 'dones': e(Tp, M, 1, dtype=torch.uint8),
        }
        if prog.lstm is not None:
            self.mb['rnn_start_c'] = e(M, prog.lstm.RH * prog.lstm.RL)      # layers side by side
            self.mb['rnn_start_h'] = e(M, prog.lstm.RH * prog.lstm.RL)
        # Double-buffered minibatches: the gather of minibatch k+1 (NVLink peer reads in the index-exact
        # data-parallel mode) runs on a side stream underneath the all-reduce + optimiser of minibatch k.
        # On a single GPU the gather hides underneath the loss tail / optimiser of minibatch k (cfg2: 5.83 -> 5.69 ms).
        # Default on; MLB_PREFETCH_GATHER=0 turns it off (one buffer set).
        pf = os.environ.get('MLB_PREFETCH_GATHER')
        self.prefetch = self.mode == 'default' and (True if pf is None else pf != '0')
        self.mb_sets = [self.mb]
        if self.prefetch:
            self.mb_sets.append({k: torch.empty_like(v) for k, v in self.mb.items()})
        self.Tp = Tp
        if self.mode != 'default':
            npad = K.sort_pad(J)
            self.sort_keys = e(npad, dtype=torch.int64)               # u64 keys of the argsort / top-k
            self.valid = e(J, dtype=torch.int32)                      # valid_inds (ml/ppo.py:403-405, :435)
            self.perm_raw = e(E, J, dtype=torch.int32)
            self.counts = torch.zeros(2, dtype=torch.int32, device=dev)
            self.max_abs = e(1)
            self.idx_tmp = e(M, dtype=torch.int32)
            self.mb_stats = e(4)
            self.mb_ret_stats = e(1, 4)
            self.vn_one = e(1, 4)
            self.mom_ws = torch.empty(K.lib().mlb_moments_workspace(Tp * M) + 16, dtype=torch.uint8, device=dev)
            if self.mode == 'importance':
                self.nsel = cfg.importance_sample_num_minibatches * M
                assert 0 < self.nsel < J, 'ml/ppo.py:416-417'
                self.scores, self.probs, self.traj_w = e(J), e(J), e(J)
                self.mb_w = e(M)
                self.sample_key = torch.zeros(2, dtype=torch.int32, device=dev)
        self.side = torch.cuda.Stream(device=dev)      # minibatch-gather prefetch stream
        A = prog.A
        g_name, _, g_size = prog.groups[0]
        coef = _entropy_coef(cfg, g_name)
        rows_global = self.rows * (dist_ctx.world_size if dist_ctx else 1)
        self.rows_global = rows_global
        self.obj_scale = (ctypes.c_float * A)(*[1.0 / (rows_global * g_size)] * A)
        ent = [coef / (rows_global * g_size)] * A
        if prog.continuous is not None:          # MLB_PPO_CONTINUOUS_ACTIONS: + stddev_min, stddev_max
            ent = ent + [prog.continuous[0], prog.continuous[1]]
        self.ent_scale = (ctypes.c_float * len(ent))(*ent)
        prog.train_ws(self.rows)


def _layer_states(lstm, mb):
    """Per-layer [M, RH] views of the gathered chunk-start states (ml/rollouts.py:471-478)."""
    RH = lstm.RH
    return dict(c0=[mb['rnn_start_c'][:, l * RH:(l + 1) * RH] for l in range(lstm.RL)],
                h0=[mb['rnn_start_h'][:, l * RH:(l + 1) * RH] for l in range(lstm.RL)])


def _compute_minibatch_indices(train_state, ws, dist_ctx, partitionable=False):
    """ml/ppo.py:445-458: E x (split the update key, permutation(arange(J))); data-parallel index-exact mode adds
    the owner-affine split of every global minibatch.  Depends on the update key only."""
    K.ppo_permutations(train_state.update_prng_key, ws.E, ws.Jp, partitionable, ws.perm, ws.perm_ws)
    if ws.idx_local is not None:
        call('mlb_dp_assign_minibatches', ptr(ws.perm), c_ll(ws.Jp), c_int(ws.E), c_int(ws.nmb),
             c_int(dist_ctx.world_size), c_int(dist_ctx.rank), c_ll(ws.B), c_ll(ws.M), ptr(ws.idx_local))


def hoist_permutations(train_state, ws, dist_ctx=None, partitionable=False):
    """The minibatch permutations of an update depend only on the update PRNG key, not on the rollout: enqueue
    them on the workspace's side stream BEFORE rollout collection, where the sort kernels (4 x 27 us at cfg2) run
    underneath the rollout phase -- the one-CTA-per-128-agents rollout kernel leaves more than half of the SMs
    idle.  _ppo joins the side stream where the reference computes the indices.  Default branch only (the
    filter_advantages / importance-sampling permutations shuffle data-dependent index sets).
    MLB_HOIST_PERM=0 disables."""
    if ws is None or ws.mode != 'default' or os.environ.get('MLB_HOIST_PERM', '1') == '0':
        return
    main = torch.cuda.current_stream()
    ws.side.wait_stream(main)
    with torch.cuda.stream(ws.side):
        _compute_minibatch_indices(train_state, ws, dist_ctx, partitionable)
    ws.perm_pending = True


def _ppo(cfg, policy_state, train_state, rollout_data, user_metrics_cb, metrics, dist_ctx=None,
         ws=None, partitionable=False):
    """ml/ppo.py:366-488, default branch (valid_inds = arange(J), weights = 1)."""
    prog = policy_state.program
    if bool(cfg.dreamer_v3_critic) != bool(prog.twohot) or bool(cfg.hlgauss_critic) != bool(prog.hlgauss):
        raise ValueError('cfg.dreamer_v3_critic / cfg.hlgauss_critic must match the critic module '
                         '(DreamerV3Critic <-> dreamer_v3_critic, HLGaussCritic <-> hlgauss_critic, '
                         'DenseLayerCritic <-> neither)')
    st = rollout_data.store
    C, Tp, B = rollout_data.C, rollout_data.Tp, rollout_data.B
    if ws is None:
        ws = _PPOWorkspace(cfg, prog, C, Tp, B, dist_ctx)
    if ws.mode != 'default':
        return _ppo_selected(cfg, policy_state, train_state, rollout_data, user_metrics_cb, metrics, ws,
                             partitionable)
    E, M, J, nmb, rows = ws.E, ws.M, ws.J, ws.nmb, ws.rows
    T, N = C * Tp, B
    tx = train_state.tx
    hp = train_state.hyper_params

    with profile('Compute Minibatch Indices'):
        if getattr(ws, 'perm_pending', False):       # hoisted underneath the rollout phase (hoist_permutations)
            torch.cuda.current_stream().wait_stream(ws.side)
            ws.perm_pending = False
        else:
            _compute_minibatch_indices(train_state, ws, dist_ctx, partitionable)

    # per-minibatch statistics for ALL minibatches of the update (App. C.1)
    score_key = 'advantages' if cfg.compute_advantages else 'returns'
    normalize_scores = cfg.normalize_advantages if cfg.compute_advantages else cfg.normalize_returns
    vn = train_state.value_normalizer
    if dist_ctx is None:
        if normalize_scores:
            K.traj_moments(st[score_key].view(T, N), C, ws.tm_adv)
            K.mb_moments(ws.tm_adv, ws.perm, M, Tp, 1e-5, ws.mb_adv)
        if vn is not None:
            K.traj_moments(st['returns'].view(T, N), C, ws.tm_ret)
            K.mb_moments(ws.tm_ret, ws.perm, M, Tp, 0.0, ws.mb_ret)
    elif ws.index_exact:
        # per-trajectory moments are all-gathered into global trajectory order; every rank then
        # evaluates the statistics of every GLOBAL minibatch itself (identical bits on all ranks).
        # The collective doubles as the "all ranks finished GAE" point for the peer gathers below.
        synced = False
        if normalize_scores:
            K.traj_moments(st[score_key].view(T, N), C, ws.tm_adv)
            dist_ctx.allgather_traj_moments(ws.tm_adv, ws.tm_adv_g, C)
            K.mb_moments(ws.tm_adv_g, ws.perm, ws.Mp, Tp, 1e-5, ws.mb_adv)
            synced = True
        if vn is not None:
            K.traj_moments(st['returns'].view(T, N), C, ws.tm_ret)
            dist_ctx.allgather_traj_moments(ws.tm_ret, ws.tm_ret_g, C)
            K.mb_moments(ws.tm_ret_g, ws.perm, ws.Mp, Tp, 0.0, ws.mb_ret)
            synced = True
        if not synced:
            dist_ctx.stream_barrier(prog.device)
    else:
        # raw (sum, sumsq) of every minibatch -> ONE all-reduce for the whole update
        Kmb = E * nmb
        if normalize_scores:
            K.traj_moments(st[score_key].view(T, N), C, ws.tm_adv)
            K.mb_moments(ws.tm_adv, ws.perm, M, Tp, 1e-5, None, ws.raw[:Kmb])
        if vn is not None:
            K.traj_moments(st['returns'].view(T, N), C, ws.tm_ret)
            K.mb_moments(ws.tm_ret, ws.perm, M, Tp, 0.0, None, ws.raw[Kmb:])
        dist_ctx.allreduce_raw_moments(ws.raw)
        cnt = float(ws.rows_global)
        if normalize_scores:
            K.moments_finalize(ws.raw[:Kmb], cnt, 1e-5, ws.mb_adv)
        if vn is not None:
            K.moments_finalize(ws.raw[Kmb:], cnt, 0.0, ws.mb_ret)
    if vn is not None:
        K.ema_scan(train_state.value_normalizer_state, ws.mb_ret, vn.decay, vn.eps, ws.vn_params)

    flags = (1 if cfg.algo.clip_value_loss else 0) | (2 if cfg.algo.huber_value_loss else 0)
    # The loss kernel averages the value loss over ITS rows; the action / entropy terms carry explicit
    # 1 / (global rows) scales.  Data-parallel: the gradient all-reduce SUMS the ranks' arenas, so the value
    # term must be this rank's share of the GLOBAL mean too (rows_local / rows_global of its local mean).
    vscale = 1.0 if dist_ctx is None else float(rows) / float(ws.rows_global)
    keys = ['obs', 'actions', 'log_probs', score_key, 'returns']
    if cfg.algo.clip_value_loss:
        keys.append('values')
    tw = prog.train_ws(rows)
    if prog.lstm is not None:
        keys.append('dones')
    nset = len(ws.mb_sets)
    if prog.tc and nset > 1 and getattr(ws, 'x_sets', None) is None:
        ws.x_sets = [tw['x'], torch.empty_like(tw['x'])]      # the bf16 observation copy the gather writes
    leaf_names = list(dict.fromkeys(keys))
    if ws.index_exact and getattr(ws, 'peer_tab', None) is None:
        ws.peer_tab = dist_ctx.peer_store_table(leaf_names)
        ws.peer_tab_rnn = dist_ctx.peer_store_table(['rnn_start_c', 'rnn_start_h']) if prog.lstm is not None else None

    def gather(e, k, mb, xbf):
        """Fill the minibatch buffer set `mb` (and the bf16 observation copy `xbf`) with minibatch (e, k)."""
        if ws.index_exact:
            # this rank's slice of the global minibatch; rows are fetched from their owners' stores
            lo = k * ws.Mp + dist_ctx.rank * M
            idx = ws.idx_local[e, k] if ws.idx_local is not None else ws.perm[e, lo:lo + M]
            leaves = [(st[name], mb[name], xbf if (name == 'obs' and prog.tc) else None)
                      for name in leaf_names]
            K.mb_gather_multi_peer(leaves, ws.peer_tab, dist_ctx.world_size, idx, C, Tp, B)
            if prog.lstm is not None:
                K.mb_gather_multi_peer([(st['rnn_start_c'], mb['rnn_start_c'], None),
                                        (st['rnn_start_h'], mb['rnn_start_h'], None)],
                                       ws.peer_tab_rnn, dist_ctx.world_size, idx, C, 1, B)
            return
        idx = ws.perm[e, k * M:(k + 1) * M]
        leaves = []
        for name in leaf_names:
            src = st[name][:, :, 0]
            leaves.append((src.view(torch.uint8) if src.dtype == torch.bool else src, mb[name],
                           xbf if (name == 'obs' and prog.tc) else None))
        K.mb_gather_multi(leaves, idx, C, Tp, B)
        if prog.lstm is not None:
            K.mb_gather_rnn(st['rnn_start_c'][:, 0], idx, C, B, mb['rnn_start_c'])
            K.mb_gather_rnn(st['rnn_start_h'][:, 0], idx, C, B, mb['rnn_start_h'])

    # Minibatch k+1 is gathered on a side stream into the OTHER buffer set while minibatch k runs its gradient
    # all-reduce and optimiser on the main stream; both streams are part of the captured update graph.
    prefetch = ws.prefetch and nset > 1
    main = torch.cuda.current_stream()
    order = [(e, k) for e in range(E) for k in range(nmb)]
    xs = getattr(ws, 'x_sets', None) or [tw.get('x') if prog.tc else None] * nset
    for it, (e, k) in enumerate(order):
            mbi = e * nmb + k
            mb, xbf = ws.mb_sets[it % nset], xs[it % nset]
            if it == 0 or not prefetch:
                with profile('Gather Minibatch'):
                    gather(e, k, mb, xbf)
            else:
                main.wait_stream(ws.side)                     # the prefetch of this minibatch has landed
            if prog.tc:
                tw['x'] = xbf                                 # layer 0's A operand / dW operand of this minibatch
            seq = None
            if prog.lstm is not None:
                seq = dict(Tp=Tp, M=M, ends=mb['dones'].view(Tp, M), **_layer_states(prog.lstm, mb))
            with profile('AC Forward'):
                head = prog.forward_train(mb['obs'].view(rows, prog.obs_dim), rows, seq, x_ready=prog.tc)
            with profile('Optimize'):
                prog.zero_grads()
                call('mlb_ppo_loss_f32', ptr(head), c_int(prog.NH), ptr(mb['actions']),
                     ptr(mb['log_probs']), ptr(mb[score_key]), ptr(mb['returns']),
                     ptr(mb['values']) if cfg.algo.clip_value_loss else ptr(None), ptr(None),
                     ptr(ws.mb_adv[mbi]) if normalize_scores else ptr(None),
                     ptr(ws.vn_params[mbi]) if vn is not None else ptr(None),
                     prog._buckets_c, ws.obj_scale, ws.ent_scale, c_int(prog.A), c_ll(rows), c_ll(M),
                     c_float(hp.clip_coef), c_float(hp.value_loss_coef * vscale), c_int(flags | prog.loss_flags),
                     ptr(tw['dhead']), ptr(prog.head_bias_grad()), ptr(tw['stats_out']), ptr(tw['loss_ws']),
                     c_size_t(tw['loss_ws'].numel()), prog._bins_c, c_int(prog.V))
                prog.backward(mb['obs'].view(rows, prog.obs_dim), rows, seq)
                grad_scale = 1.0
                if prefetch and it + 1 < len(order):
                    # the persistent layer kernels own every SM's registers, so a gather issued earlier would only
                    # displace them; the all-reduce and optimiser kernels that follow are latency-bound and leave
                    # the SMs (and NVLink) idle: the next minibatch's peer gather runs underneath THEM
                    ws.side.wait_stream(main)
                    with torch.cuda.stream(ws.side):
                        gather(*order[it + 1], ws.mb_sets[(it + 1) % nset], xs[(it + 1) % nset])
                reduced = None
                if dist_ctx is not None:                      # sum over ranks; scales already global
                    if dist_ctx.fused:
                        reduced = dist_ctx.allreduce_grads_fused(prog)     # NVLink peer reads + norm, one kernel
                    else:
                        dist_ctx.allreduce_grads(prog.grads)
                prog.optimizer_step(tx['lr'], tx['max_grad_norm'], grad_scale, tx['b1'], tx['b2'], tx['eps'],
                                    reduced=reduced)
            with profile('Metrics Callback'):
                metrics = user_metrics_cb(metrics, e, mb, policy_state, train_state)
    if prefetch:
        main.wait_stream(ws.side)
    with profile('Record Metrics'):
        # per-minibatch records overwrite the same slot: the last minibatch wins (App. C.4)
        dst = metrics.slot('Loss', 5)
        call('mlb_copy_bytes', ptr(tw['stats_out'][16:]), ptr(dst), c_size_t(dst.numel()))
    return policy_state, train_state, metrics


def _ppo_selected(cfg, policy_state, train_state, rollout_data, user_metrics_cb, metrics, ws, partitionable):
    """The two alternate minibatch selections of ml/ppo.py:374-435 in front of the same epoch x
    minibatch loop (:445-486).

    filter_advantages (:374-405): time is flattened; valid_inds = the indices of the largest
    abs(advantage) elements, as many whole minibatches as there are elements >= 1 % of the running
    (EMAEstimate) maximum; the rest is -1 and is partitioned to the tail of every epoch's
    permutation.  The minibatch count is data dependent: it is read back once per update (one
    4-byte device->host copy) and the update runs eagerly (no CUDA graph).
    importance_sample_trajectories (:407-435): trajectories are drawn without replacement with
    probability softmax(mean abs(adv) + mean abs(value error)) (Gumbel top-k, jax.random.choice) and
    weighted by (1/J)/p in the loss."""
    prog = policy_state.program
    st = rollout_data.store
    C0, Tp0, B0 = ws.store_dims
    E, M, J = ws.E, ws.M, ws.J
    Tp = ws.Tp                                    # 1 in filter mode
    rows = Tp * M
    hp, tx = train_state.hyper_params, train_state.tx
    score_key = 'advantages' if cfg.compute_advantages else 'returns'
    normalize_scores = cfg.normalize_advantages if cfg.compute_advantages else cfg.normalize_returns
    vn = train_state.value_normalizer
    flat = lambda name: st[name].view(-1) if st[name].dtype != torch.bool else st[name].view(torch.uint8).view(-1)
    mb_w = None
    if ws.mode == 'filter':
        with profile('Filter Advantages'):
            call('mlb_filter_adv_keys', ptr(flat('advantages')), c_int(C0), c_int(Tp0), c_ll(B0),
                 c_ll(ws.sort_keys.numel()), ptr(ws.sort_keys), ptr(ws.max_abs))
            K.sort_u64(ws.sort_keys)
            est = train_state.max_advantage_est_state
            train_state.max_advantage_est.update_estimates(est, ws.max_abs)           # :381-388
            call('mlb_filter_adv_select', ptr(ws.sort_keys), c_ll(J), c_ll(M), ptr(est), ptr(ws.valid),
                 ptr(ws.counts))
            nmb = int(ws.counts[0].item())            # data-dependent loop bound (:399-402)
            nsel = J
    else:
        with profile('Importance Sample Trajectories'):
            call('mlb_traj_scores_f32', ptr(flat('advantages')), ptr(flat('values')), ptr(flat('returns')),
                 c_int(C0), c_int(Tp0), c_ll(B0), ptr(ws.scores))
            call('mlb_softmax_weights_f32', ptr(ws.scores), c_ll(J), ptr(ws.probs), ptr(ws.traj_w))
            ks = K.threefry_split(train_state.update_prng_key, 2, partitionable)      # gen_update_rnd (:429)
            ws.sample_key.copy_(ks[0])
            train_state.update_prng_key.copy_(ks[1])
            call('mlb_gumbel_topk_keys', ptr(ws.sample_key), ptr(ws.probs), c_ll(J), c_ll(ws.sort_keys.numel()),
                 c_int(int(partitionable)), ptr(ws.sort_keys))
            K.sort_u64(ws.sort_keys)
            nsel = ws.nsel
            call('mlb_take_sorted_indices', ptr(ws.sort_keys), c_ll(nsel), ptr(ws.valid))
            nmb = nsel // M
            mb_w = ws.mb_w
    with profile('Compute Minibatch Indices'):
        # per epoch: permutation(rnd, valid_inds), then the stable partition of the -1s (:445-458)
        perm_raw = ws.perm_raw[:, :nsel] if nsel == J else ws.perm_raw.view(-1)[:E * nsel].view(E, nsel)
        perm = ws.perm[:, :nsel] if nsel == J else ws.perm.view(-1)[:E * nsel].view(E, nsel)
        call('mlb_ppo_permutations_of', ptr(train_state.update_prng_key), ptr(ws.valid), ptr(perm_raw), c_int(E),
             c_ll(nsel), c_int(int(partitionable)), ptr(ws.perm_ws), c_size_t(ws.perm_ws.numel()))
        call('mlb_partition_valid', ptr(perm_raw), ptr(perm), c_int(E), c_ll(nsel))
    flags = (1 if cfg.algo.clip_value_loss else 0) | (2 if cfg.algo.huber_value_loss else 0)
    keys = ['obs', 'actions', 'log_probs', score_key, 'returns']
    if cfg.algo.clip_value_loss:
        keys.append('values')
    keys = list(dict.fromkeys(keys))
    tw = prog.train_ws(rows)
    mb = ws.mb
    seq = None
    if prog.lstm is not None:
        keys.append('dones')
        seq = dict(Tp=Tp, M=M, ends=mb['dones'].view(Tp, M), **_layer_states(prog.lstm, mb))
    for e in range(E):
        for k in range(nmb):
            idx = perm[e, k * M:(k + 1) * M]
            with profile('Gather Minibatch'):
                if ws.mode == 'filter':
                    call('mlb_flat_time_index', ptr(idx), c_ll(M), c_int(Tp0), c_ll(B0), ptr(ws.idx_tmp))
                    leaves = []
                    for name in keys:
                        src = st[name]
                        src = src.view(torch.uint8) if src.dtype == torch.bool else src
                        leaves.append((src.view(1, 1, C0 * Tp0 * B0, *src.shape[4:]), mb[name],
                                       tw['x'] if (name == 'obs' and prog.tc) else None))
                    K.mb_gather_multi(leaves, ws.idx_tmp, 1, 1, C0 * Tp0 * B0)
                else:
                    leaves = []
                    for name in keys:
                        src = st[name][:, :, 0]
                        leaves.append((src.view(torch.uint8) if src.dtype == torch.bool else src, mb[name],
                                       tw['x'] if (name == 'obs' and prog.tc) else None))
                    K.mb_gather_multi(leaves, idx, C0, Tp0, B0)
                    if seq is not None:
                        K.mb_gather_rnn(st['rnn_start_c'][:, 0], idx, C0, B0, mb['rnn_start_c'])
                        K.mb_gather_rnn(st['rnn_start_h'][:, 0], idx, C0, B0, mb['rnn_start_h'])
                    call('mlb_gather_f32', ptr(ws.traj_w), ptr(idx), c_ll(M), ptr(mb_w))      # :468
            if normalize_scores:                      # zscore_data over THIS minibatch (:134-143)
                K.moments(mb[score_key].view(-1), 1e-5, ws.mb_stats, ws.mom_ws)
            if vn is not None:                        # normalize_and_update_estimates (:205-211)
                K.moments(mb['returns'].view(-1), 0.0, ws.mb_ret_stats.view(-1), ws.mom_ws)
                K.ema_scan(train_state.value_normalizer_state, ws.mb_ret_stats, vn.decay, vn.eps, ws.vn_one)
            with profile('AC Forward'):
                head = prog.forward_train(mb['obs'].view(rows, prog.obs_dim), rows, seq, x_ready=prog.tc)
            with profile('Optimize'):
                prog.zero_grads()
                call('mlb_ppo_loss_f32', ptr(head), c_int(prog.NH), ptr(mb['actions']),
                     ptr(mb['log_probs']), ptr(mb[score_key]), ptr(mb['returns']),
                     ptr(mb['values']) if cfg.algo.clip_value_loss else ptr(None), ptr(mb_w),
                     ptr(ws.mb_stats) if normalize_scores else ptr(None),
                     ptr(ws.vn_one) if vn is not None else ptr(None),
                     prog._buckets_c, ws.obj_scale, ws.ent_scale, c_int(prog.A), c_ll(rows), c_ll(M),
                     c_float(hp.clip_coef), c_float(hp.value_loss_coef), c_int(flags | prog.loss_flags),
                     ptr(tw['dhead']), ptr(prog.head_bias_grad()), ptr(tw['stats_out']), ptr(tw['loss_ws']),
                     c_size_t(tw['loss_ws'].numel()), prog._bins_c, c_int(prog.V))
                prog.backward(mb['obs'].view(rows, prog.obs_dim), rows, seq)
                prog.optimizer_step(tx['lr'], tx['max_grad_norm'], 1.0, tx['b1'], tx['b2'], tx['eps'])
            with profile('Metrics Callback'):
                metrics = user_metrics_cb(metrics, e, mb, policy_state, train_state)
    ws.last_num_minibatches = nmb
    if nmb > 0:
        with profile('Record Metrics'):
            dst = metrics.slot('Loss', 5)
            call('mlb_copy_bytes', ptr(tw['stats_out'][16:]), ptr(dst), c_size_t(dst.numel()))
    return policy_state, train_state, metrics
