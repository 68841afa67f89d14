"""Rollout buffers + inference loop (ml/rollouts.py:28-978), non-PBT branch (P = 1).

Differences from the reference that are deliberate B200 design, not omissions:
  * buffers are allocated once and reused (XLA donation made explicit);
  * the per-step store (`store.at[(c, s)].set`, :356-367) is the kernels writing straight
    into the [c, s] slab of the [C, T', P, B, ...] store;
  * `_finalize_rollouts` never materialises the [P, C*B, T', ...] relayout (:788-804):
    GAE runs on the [T, N] view and minibatches are gathered from the store by the K5 kernel
    (`RolloutData.minibatch`), bit-identical to take+swapaxes on the relayout.
"""
import math
import os
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional

import torch

from . import kernels as K
from ._lib import c_float, c_int, c_ll, c_size_t, call, ptr
from .algo_common import compute_advantages, compute_returns
from .cfg import DiscreteActionsConfig
from .profile import profile


@dataclass(frozen=True)
class PBTMatchmakeConfig:               # the arithmetic RolloutConfig.setup needs (ml/pbt.py:21-118)
    num_current_policies: int
    num_past_policies: int
    total_num_policies: int
    num_teams: int
    team_size: int
    self_play_batch_size: int
    cross_play_batch_size: int
    past_play_batch_size: int
    static_play_batch_size: int
    complex_matchmaking: bool
    custom_policy_ids: List[int]


@dataclass(frozen=True)
class RolloutConfig:                    # ml/rollouts.py:28-134
    sim_batch_size: int
    num_worlds: int
    actions_cfg: Dict[str, Any]
    policy_chunk_size: int
    num_policy_chunks: int
    total_policy_batch_size: int
    reward_gamma: float
    policy_dtype: Any
    reward_dtype: Any
    prob_dtype: Any
    pbt: PBTMatchmakeConfig

    @staticmethod
    def setup(num_current_policies, num_past_policies, num_teams, team_size, sim_batch_size,
              actions_cfg, self_play_portion, cross_play_portion, past_play_portion,
              static_play_portion, reward_gamma, custom_policy_ids, policy_dtype,
              reward_dtype=torch.float32, prob_dtype=torch.float32,
              policy_chunk_size_override=0):
        assert (self_play_portion + cross_play_portion + past_play_portion +
                static_play_portion == 1.0)
        if self_play_portion != 1.0 or num_past_policies != 0 or num_current_policies != 1:
            raise NotImplementedError('PBT matchmaking / multi-policy batching is out of scope '
                                      '(SURVEY 8f rank 1); only pbt=None configs are lowered')
        pbt = PBTMatchmakeConfig(
            num_current_policies=1, num_past_policies=0, total_num_policies=1,
            num_teams=num_teams, team_size=team_size, self_play_batch_size=sim_batch_size,
            cross_play_batch_size=0, past_play_batch_size=0, static_play_batch_size=0,
            complex_matchmaking=False, custom_policy_ids=list(custom_policy_ids))
        chunk = sim_batch_size // num_current_policies
        if policy_chunk_size_override != 0:
            chunk = policy_chunk_size_override
        nchunks = -(sim_batch_size // -chunk)
        return RolloutConfig(
            sim_batch_size=sim_batch_size, num_worlds=sim_batch_size // (team_size * num_teams),
            actions_cfg=actions_cfg, policy_chunk_size=chunk, num_policy_chunks=nchunks,
            total_policy_batch_size=nchunks * chunk, reward_gamma=reward_gamma,
            policy_dtype=policy_dtype, reward_dtype=reward_dtype, prob_dtype=prob_dtype, pbt=pbt)


@dataclass
class RolloutState:                     # ml/rollouts.py:171-183
    cfg: RolloutConfig
    step_fn: Callable
    load_ckpts_fn: Optional[Callable]
    get_ckpts_fn: Optional[Callable]
    sim_state: Any
    cur_obs: Dict[str, torch.Tensor]
    prng_key: torch.Tensor              # int32[2] on device
    rnn_states: Any
    reorder_state: Any
    policy_assignments: torch.Tensor
    sim_ctrl: Any
    env_returns: torch.Tensor

    @staticmethod
    def create(rollout_cfg, sim_fns, prng_key, rnn_states, init_sim_ctrl,
               static_play_assignments, device, partitionable=False):
        # prng_key, assign_rnd = split(prng_key)   (ml/rollouts.py:200)
        ks = K.threefry_split(prng_key, 2, partitionable)
        prng_key = ks[0].clone()
        init_out = sim_fns['init']()
        return RolloutState(
            cfg=rollout_cfg, step_fn=sim_fns['step'], load_ckpts_fn=sim_fns.get('load_ckpts'),
            get_ckpts_fn=sim_fns.get('get_ckpts'), sim_state=init_out['state'],
            cur_obs=dict(init_out['obs']), prng_key=prng_key, rnn_states=rnn_states,
            reorder_state=None,
            policy_assignments=torch.zeros(rollout_cfg.sim_batch_size, dtype=torch.int32, device=device),
            sim_ctrl=init_sim_ctrl,
            env_returns=torch.zeros(rollout_cfg.sim_batch_size, 1, dtype=torch.float32, device=device))

    def update(self, **kw):
        for k, v in kw.items():
            if v is not None:
                setattr(self, k, v)
        return self


class RolloutData:                      # ml/rollouts.py:311-334
    """`store` holds every per-step leaf as [C, T', P=1, B, *leaf] (+ rnn_start_states
    [C, P, B, *]).  `all()` materialises the reference's training layout [J, T', ...] on
    demand (hooks / debugging); the learner never does."""

    def __init__(self, store, C, Tp, B):
        self.store = store
        self.C, self.Tp, self.B = C, Tp, B
        self.num_train_seqs_per_policy = C * B
        self.num_train_policies = 1
        self._out = {}

    def all(self):
        def relayout(x):
            t = x.permute(2, 0, 3, 1, *range(4, x.dim()))
            return t.reshape(t.shape[0], -1, *t.shape[3:])[0]
        return {k: relayout(v) for k, v in self.store.items() if not k.startswith('rnn_start')}

    def minibatch(self, indices, keys=None, out=None):
        """-> dict name -> [T', M, *leaf]  (take(axis 0) + swapaxes(0, 1), :319-329)."""
        res = {}
        for k, v in self.store.items():
            if k.startswith('rnn_start') or (keys is not None and k not in keys):
                continue
            res[k] = K.mb_gather(v[:, :, 0], indices, self.C, self.Tp, self.B,
                                 None if out is None else out[k])
        return res


def _copy_state(dst, src):
    """dst[...] = src for one [N, RH] recurrent-state block (a feature slice of the store when RL > 1)."""
    if dst.is_contiguous() and src.is_contiguous():
        call('mlb_copy_bytes', ptr(src), ptr(dst), c_size_t(src.numel() * src.element_size()))
    else:
        dst.copy_(src)


class RolloutManager:                   # ml/rollouts.py:373-826
    def __init__(self, train_cfg, init_rollout_state, example_policy_states, dist_ctx=None):
        self._cfg = init_rollout_state.cfg
        # distributional critics store their mean as the value estimate (ml/rollouts.py:384, 601-605);
        # the sampling kernel emits the two-hot / HL-Gauss mean directly
        self._critic_outputs_distribution = train_cfg.dreamer_v3_critic or train_cfg.hlgauss_critic
        self._num_bptt_chunks = train_cfg.num_bptt_chunks
        assert train_cfg.steps_per_update % train_cfg.num_bptt_chunks == 0
        self._num_bptt_steps = train_cfg.steps_per_update // train_cfg.num_bptt_chunks
        self._num_train_policies = 1
        self._num_train_agents_per_policy = self._cfg.sim_batch_size
        self._num_train_seqs_per_policy = self._num_train_agents_per_policy * self._num_bptt_chunks
        self._use_advantages = train_cfg.compute_advantages
        self._train_cfg = train_cfg
        prog = example_policy_states.program
        self._prog = prog
        dev = prog.device
        C, Tp, B = self._num_bptt_chunks, self._num_bptt_steps, self._cfg.sim_batch_size
        (ob_name, ob), = init_rollout_state.cur_obs.items()
        self._ob_name = ob_name
        self._act_name = prog.groups[0][0]
        e = lambda *s, dtype=torch.float32: torch.empty(*s, dtype=dtype, device=dev)
        # store schema: ml/rollouts.py:409-480
        f32 = torch.float32
        specs = {
            'obs': ((C, Tp, 1, B, prog.obs_dim), f32),
            'actions': ((C, Tp, 1, B, prog.A), torch.int32),
            'log_probs': ((C, Tp, 1, B, prog.A), f32),
            'rewards': ((C, Tp, 1, B, 1), f32),
            'dones': ((C, Tp, 1, B, 1), torch.bool),
            'values': ((C, Tp, 1, B, 1), f32),
            'advantages': ((C, Tp, 1, B, 1), f32),
            'returns': ((C, Tp, 1, B, 1), f32),
        }
        self._lstm = prog.lstm
        if self._lstm is not None:
            # rnn_start_states [C, P, B, *] (ml/rollouts.py:471-478): the carry entering each BPTT chunk
            # (layers concatenated along the feature axis: [.., l*RH:(l+1)*RH] is layer l's state)
            RH = self._lstm.RH * self._lstm.RL
            specs['rnn_start_c'] = ((C, 1, B, RH), f32)
            specs['rnn_start_h'] = ((C, 1, B, RH), f32)
            self._boot_states = self._lstm.init_states(B, dev)
        # data-parallel index-exact mode: the stores live in NVLink symmetric memory so that peers can
        # gather the trajectories of a GLOBAL minibatch permutation straight out of them (parallel.py)
        self.store = dist_ctx.alloc_symmetric_stores(specs, dev) if dist_ctx is not None and \
            hasattr(dist_ctx, 'alloc_symmetric_stores') else None
        if self.store is None:
            self.store = {k: e(*shape, dtype=dt) for k, (shape, dt) in specs.items()}
        self.bootstrap = e(1, B, 1)
        self._scratch_actions = e(B, prog.A, dtype=torch.int32)
        self.env_returns_trace = e(C * Tp, B)
        self.policy_key = torch.zeros(2, dtype=torch.int32, device=dev)
        self._key_alt = torch.zeros(2, dtype=torch.int32, device=dev)
        self.resets = torch.zeros(self._cfg.num_worlds, 1, dtype=torch.int32, device=dev)
        T = C * Tp
        self._gae_ws = torch.empty(K.lib().mlb_gae_workspace(T, B) + 16, dtype=torch.uint8, device=dev)
        self._met_ws = torch.empty(K.lib().mlb_moments_workspace(T * B) + 16, dtype=torch.uint8, device=dev)
        self.partitionable = False
        # fused one-launch rollout step: the post-step store of step t rides in the policy launch of step t + 1
        # (mlb_policy_rollout_ps_tc); MLB_FUSE_POST_STEP=0 keeps the separate mlb_post_step_store_f32 launch
        self._fuse_post_step = (prog.fused_rollout and self._lstm is None and
                                os.environ.get('MLB_FUSE_POST_STEP', '1') != '0')
        self._pending_post_step = None
        # per-update observation statistics (ml/rollouts.py:670-678); None entries for preprocessors
        # without state (Noop / Caster)
        self._obs_stats = example_policy_states.obs_preprocess.init_obs_stats(
            example_policy_states.obs_preprocess_state, True, num_steps=C * Tp,
            example_obs=init_rollout_state.cur_obs)

    def add_metrics(self, train_cfg, metrics):          # :482-499
        names = ['Rewards', 'Est Returns', 'Env Returns', 'Values']
        if train_cfg.compute_advantages:
            names.append('Advantages')
        names.append('Bootstrap Values')
        return metrics + names

    # ---------------------------------------------------------------------------------
    def collect(self, train_state_mgr, rollout_state, metrics, user_start_rollouts_hook,
                user_finish_rollouts_hook, user_metrics_hook):
        """ml/rollouts.py:501-577."""
        rollout_state, user_state = user_start_rollouts_hook(rollout_state, train_state_mgr.user_state)
        train_state_mgr.user_state = user_state
        policy_states = train_state_mgr.policy_states
        for c in range(self._num_bptt_chunks):
            self.begin_chunk(rollout_state, c)
            rollout_state = self.rollout_loop(rollout_state, policy_states, c)
        return self.finish(train_state_mgr, rollout_state, metrics, user_finish_rollouts_hook, user_metrics_hook)

    # The pieces of collect() -- also driven in lockstep over several policies that share one simulator
    # (multi_policy.py: every policy runs policy_step for step s, THEN the simulator steps once).
    def begin_chunk(self, rs, c):
        if self._lstm is not None:
            with profile('Cache RNN state'):              # :533-537
                RH = self._lstm.RH
                for l in range(self._lstm.RL):
                    _copy_state(self.store['rnn_start_c'][c, 0][:, l * RH:(l + 1) * RH], rs.rnn_states[0][l])
                    _copy_state(self.store['rnn_start_h'][c, 0][:, l * RH:(l + 1) * RH], rs.rnn_states[1][l])
        self._key_home = rs.prng_key

    def finish(self, train_state_mgr, rollout_state, metrics, user_finish_rollouts_hook, user_metrics_hook):
        with profile('Bootstrap Values'):
            self._bootstrap_values(train_state_mgr.policy_states, rollout_state)
        with profile('Finalize Rollouts'):
            rollout_data, metrics, user_state = self._finalize_rollouts(
                train_state_mgr.train_states, metrics, train_state_mgr.user_state, user_finish_rollouts_hook,
                user_metrics_hook)
        train_state_mgr.user_state = user_state
        return train_state_mgr, rollout_state, rollout_data, self._obs_stats, metrics

    def policy_step(self, rs, policy_states, c, s):
        """Policy inference of step s of BPTT chunk c on rs.cur_obs; returns the sampled actions (the store slab:
        int32 [B, A], fp32 bit patterns for a continuous group)."""
        prog, N, st = self._prog, self._cfg.sim_batch_size, self.store
        Tp = self._num_bptt_steps
        with profile('Policy Inference'):
            pre = policy_states.obs_preprocess.preprocess(
                policy_states.obs_preprocess_state, rs.cur_obs, True)
            ob = pre[self._ob_name]
            policy_states.obs_preprocess.update_obs_stats(
                policy_states.obs_preprocess_state, self._obs_stats, c * Tp + s, rs.cur_obs, True)
            slab = st['obs'][c, s, 0]
            actions = st['actions'][c, s, 0]
            if prog.fused_rollout:
                # key chain + obs store + MLP + heads + sampling in one launch; the advanced
                # PRNG key lands in the alternate buffer, so the two buffers trade places
                prog.rollout_step_fused(ob, slab, N, rs.prng_key, self._key_alt, actions,
                                        st['log_probs'][c, s, 0], st['values'][c, s, 0],
                                        self.partitionable, post_step=self._take_pending_post_step())
                rs.prng_key, self._key_alt = self._key_alt, rs.prng_key
            else:
                call('mlb_rollout_keys', ptr(rs.prng_key), ptr(self.policy_key),
                     c_int(int(self.partitionable)))
                call('mlb_copy_bytes', ptr(ob), ptr(slab), c_size_t(slab.numel() * 4))
                head = prog.forward_infer(slab, N, rs.rnn_states)
                prog.sample(head, N, self.policy_key, actions, st['log_probs'][c, s, 0],
                            st['values'][c, s, 0], self.partitionable)
        return actions

    def post_step(self, rs, c, s, out):
        """Consumes one simulator step's outputs (this policy's rows): next observations, reward / done store,
        env-return trace, RNN reset."""
        N, st = self._cfg.sim_batch_size, self.store
        Tp = self._num_bptt_steps
        rs.sim_state = out['state']
        rs.cur_obs = dict(out['obs'])
        dones, rewards = out['dones'], out['rewards']
        if dones.dtype not in (torch.bool, torch.uint8):
            dones = dones != 0
        if rewards.dtype != torch.float32:
            rewards = rewards.float()
        with profile('Post Step Rollout Store'):
            d_slab, r_slab = st['dones'][c, s, 0], st['rewards'][c, s, 0]
            if self._fuse_post_step:
                # rides in the NEXT policy launch of this rollout (the next step's, or the bootstrap's): one
                # launch fewer per step.  The simulator's reward / done buffers are still intact then -- the
                # policy launch precedes the next simulator step.
                assert self._pending_post_step is None
                self._pending_post_step = (rewards.contiguous(), dones.contiguous(), r_slab, d_slab, rs.env_returns,
                                           self.env_returns_trace[c * Tp + s], self._cfg.reward_gamma)
                return
            call('mlb_post_step_store_f32', ptr(rewards), ptr(dones), ptr(r_slab), ptr(d_slab),
                 ptr(rs.env_returns), ptr(self.env_returns_trace[c * Tp + s]), c_ll(N),
                 c_float(self._cfg.reward_gamma))
            if self._lstm is not None:                    # rnn_reset_fn(rnn_states, dones)  (:942)
                self._lstm.reset(rs.rnn_states, d_slab.view(torch.uint8), N)

    def _take_pending_post_step(self):
        p, self._pending_post_step = self._pending_post_step, None
        return p

    def end_chunk(self, rs):
        key_home = self._key_home
        if rs.prng_key is not key_home:                   # odd number of fused steps: move the key home
            call('mlb_copy_bytes', ptr(rs.prng_key), ptr(key_home), c_size_t(8))
            rs.prng_key, self._key_alt = key_home, rs.prng_key

    def rollout_loop(self, rs, policy_states, c):
        """rollout_iter x T' (ml/rollouts.py:829-978) for BPTT chunk c."""
        prog = self._prog
        for s in range(self._num_bptt_steps):
            actions = self.policy_step(rs, policy_states, c, s)
            with profile('Rollout Step'):
                step_input = {
                    'state': rs.sim_state,
                    # continuous groups: the int32 buffer holds fp32 bit patterns (ml/rollouts.py:985-998 hands f32)
                    'actions': {self._act_name: actions.view(torch.float32) if prog.continuous is not None else actions},
                    'resets': self.resets, 'sim_ctrl': rs.sim_ctrl,
                    'pbt': {'policy_assignments': rs.policy_assignments},
                }
                out = rs.step_fn(step_input)
            self.post_step(rs, c, s, out)
        self.end_chunk(rs)
        return rs

    def _bootstrap_values(self, policy_states, rs):     # :607-635
        prog, N = self._prog, self._cfg.sim_batch_size
        pre = policy_states.obs_preprocess.preprocess(policy_states.obs_preprocess_state, rs.cur_obs, False)
        ob = pre[self._ob_name].reshape(N, prog.obs_dim)
        states = None
        if self._lstm is not None:                        # critic_only must not advance the rollout's carry
            states = self._boot_states
            for l in range(self._lstm.RL):
                _copy_state(states[0][l], rs.rnn_states[0][l])
                _copy_state(states[1][l], rs.rnn_states[1][l])
        # critic column of the head -> bootstrap [1, B, 1] (the greedy actions are discarded)
        if prog.fused_rollout:
            prog.rollout_step_fused(ob, None, N, None, None, self._scratch_actions, None, self.bootstrap,
                                    deterministic=True, post_step=self._take_pending_post_step())
            return
        head = prog.forward_infer(ob, N, states)
        prog.sample(head, N, None, self._scratch_actions, None, self.bootstrap, deterministic=True)

    def _finalize_rollouts(self, train_states, metrics, user_state, finish_hook, metrics_hook):
        """ml/rollouts.py:716-826."""
        st, cfg = self.store, self._train_cfg
        vn_state = train_states.value_normalizer_state if train_states.value_normalizer is not None else None
        # ml/rollouts.py:726-745: with a value normaliser the hook (and the 'Bootstrap Values'
        # metric, :810) see value_normalizer.invert(...) of the stored values / bootstrap
        unnorm_values, unnorm_boot = st['values'], self.bootstrap
        if vn_state is not None:
            if getattr(self, '_unnorm_boot', None) is None:
                self._unnorm_boot = torch.empty_like(self.bootstrap)
                self._unnorm_values = torch.empty_like(st['values'])
            unnorm_boot = K.ema_invert(vn_state, 1, self.bootstrap, self._unnorm_boot)
            hook_fn = getattr(finish_hook, '__func__', finish_hook)
            if getattr(hook_fn, '__qualname__', '') != 'TrainHooks.finish_rollouts':
                unnorm_values = K.ema_invert(vn_state, 1, st['values'], self._unnorm_values)
        hooked, user_state = finish_hook(st, self.bootstrap, unnorm_values, unnorm_boot, user_state)
        if hooked is not st:
            raise NotImplementedError('finish_rollouts must modify the rollout store in place')
        if self._use_advantages:
            compute_advantages(cfg, st['rewards'], st['values'], st['dones'], self.bootstrap,
                               advantages=st['advantages'], returns=st['returns'],
                               vn_mu_sigma=vn_state, metrics=metrics.slot('Rewards', 4)
                               if self._metrics_contiguous(metrics) else None, ws=self._gae_ws)
        else:
            if vn_state is not None:
                raise NotImplementedError('normalize_values with compute_advantages=False')
            compute_returns(cfg, st['rewards'], st['dones'], self.bootstrap, returns=st['returns'])
        K.metric(unnorm_boot, metrics.slot('Bootstrap Values'), self._met_ws)
        K.metric(self.env_returns_trace, metrics.slot('Env Returns'), self._met_ws)
        data = RolloutData(st, self._num_bptt_chunks, self._num_bptt_steps, self._cfg.sim_batch_size)
        metrics = metrics_hook(metrics, data, user_state)
        return data, metrics, user_state

    @staticmethod
    def _metrics_contiguous(metrics):
        """The GAE kernel writes 4 consecutive records: Rewards, Values, Est Returns,
        Advantages -- the training metrics table is laid out that way by init_training."""
        i = metrics.index
        return (i.get('Values') == i['Rewards'] + 1 and i.get('Est Returns') == i['Rewards'] + 2 and
                i.get('Advantages') == i['Rewards'] + 3)
