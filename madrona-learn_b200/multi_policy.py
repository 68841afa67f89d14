"""Multi-policy learner: P policies trained side by side on one simulator.

The reference vmaps rollout inference and `_update_impl` over the policy axis (ml/train.py:165-174) after
bringing the simulator's batch into training order with `_sim_to_train` (ml/rollouts.py:579-588): without complex
matchmaking that is `x.reshape(num_train_policies, -1, ...)` -- policy p owns rows [p * B, (p + 1) * B) of the
simulator batch, B = sim_batch_size / P -- and every policy's rollout store, value normaliser, optimiser state
and metrics are independent leaves of the vmapped pytrees.

Here a vmap over P is P programs on P parameter arenas.  Each policy is a complete single-policy learner
(`TrainingManager`: PolicyProgram, RolloutManager with its own [C, T', 1, B, *] store, PPO workspace) built on a
VIEW of its block of the simulator's buffers, and this manager drives them in lockstep:

    for every step:   policy_step(p) for all p  ->  actions scattered into the simulator's [S, A] action buffer
                      (mlb_copy_bytes for the block layout; mlb_gather_rows_clip through PolicyBatchReorderState
                      .to_sim for an arbitrary assignment vector)  ->  ONE simulator step  ->  post_step(p) on the
                      block views of the new observations / rewards / dones
    then per policy:  bootstrap values, GAE, the PPO epochs  (RolloutManager.finish, train._learn_impl)

and the whole multi-policy update is captured into one CUDA graph, as the single-policy update is.

Scope: the reference's simple-matchmaking mode -- PBTConfig(num_past_policies=0, self_play_portion=1.0): every
match is self-play of one training policy.  `policy_assignments` (what the simulator is told, ml/rollouts.py:
905-910) is block p -> policy p.  With `assignments=` (a permutation of that: agents of the P policies
interleaved in simulator order) the observation gather / action scatter go through pbt_reorder's
`_compute_reorder_chunks` index construction.  Cross-play / past-play matchmaking, Elo and policy culling
(ml/pbt.py, ml/train.py:397-574) stay out of scope and are refused.
"""
import dataclasses
import os

import torch

from . import kernels as K
from ._lib import c_size_t, call, ptr
from .pbt_reorder import reorder_state_for


def sim_to_train_indices(num_train_policies, sim_batch_size, assignments=None):
    """Rows of the simulator batch that policy p trains on, [P, B] (host ints): the reference's
    `_compute_sim_to_train_indices` for self-play-only matchmaking (ml/rollouts.py:1071-1104 -- there
    `arange(S).reshape(P, -1)`, pinned by tests/golden/multi_policy.npz), or, for an explicit assignment vector,
    the rows assigned to p in simulator order (what pbt_reorder's chunk p holds)."""
    import numpy as np
    P, S = int(num_train_policies), int(sim_batch_size)
    if assignments is None:
        return np.arange(S, dtype=np.int64).reshape(P, S // P)
    a = np.asarray(assignments).reshape(-1)
    return np.stack([np.flatnonzero(a == p) for p in range(P)]).astype(np.int64)


def _check_pbt(cfg):
    pbt = cfg.pbt
    if pbt.num_past_policies != 0 or pbt.self_play_portion != 1.0 or pbt.cross_play_portion != 0.0 or \
            pbt.past_play_portion != 0.0:
        raise NotImplementedError('multi-policy learner: only self-play of the training policies is lowered '
                                  '(num_past_policies=0, self_play_portion=1.0); cross-play / past-play '
                                  'matchmaking is out of scope (SURVEY 8f)')
    if pbt.reward_hyper_params_explore:
        raise NotImplementedError('reward hyper-parameter exploration is part of PBT culling (out of scope)')
    P = int(pbt.num_train_policies)
    if P < 1 or cfg.num_worlds % P != 0:
        raise ValueError('num_worlds must be a multiple of pbt.num_train_policies')
    if pbt.num_teams * pbt.team_size != cfg.num_agents_per_world:
        raise ValueError('pbt.num_teams * pbt.team_size must equal num_agents_per_world')
    return P


class MultiPolicyTrainingManager:
    """TrainingManager of P policies (ml/train.py:35-64 with num_train_policies = P)."""

    def __init__(self, subs, cfg, sim_fns, sim_init, assignments, dev):
        self.subs, self.cfg, self.P = subs, cfg, len(subs)
        self.update_idx = subs[0].update_idx
        self._step_fn = sim_fns['step']
        self._sim_state = sim_init['state']
        self.device = dev
        S = cfg.num_worlds * cfg.num_agents_per_world
        self.S, self.B = S, S // self.P
        prog = subs[0].state.policy_states.program
        self._act_name, self._continuous = prog.groups[0][0], prog.continuous is not None
        self.actions = torch.zeros(S, prog.A, dtype=torch.int32, device=dev)          # what the simulator reads
        self._train_actions = torch.zeros(S, prog.A, dtype=torch.int32, device=dev)    # training (block) order
        self.resets = torch.zeros(cfg.num_worlds, 1, dtype=torch.int32, device=dev)
        # policy_assignments [S] (ml/rollouts.py:200-204): block p <-> policy p, or the caller's vector
        if assignments is None:
            self.policy_assignments = torch.arange(self.P, dtype=torch.int32, device=dev).repeat_interleave(self.B)
            self.reorder = None
        else:
            a = assignments.to(device=dev, dtype=torch.int32).contiguous()
            if a.numel() != S or torch.bincount(a.long(), minlength=self.P).tolist() != [self.B] * self.P:
                raise ValueError('assignments must give every training policy exactly sim_batch_size / P agents')
            self.policy_assignments = a
            # chunk size = B: one full chunk per policy, chunk p = the rows of policy p in simulator order
            self.reorder = reorder_state_for(a, self.P, self.B)
            self._train_obs = None
        self._graph = None
        self._eager_iters = 0
        self.use_cuda_graph = os.environ.get('MLB_CUDA_GRAPH', '1') != '0' and \
            all(m.use_cuda_graph for m in subs)
        self.metrics = [m.metrics for m in subs]
        self.state = [m.state for m in subs]

    # -- simulator batch <-> training order -------------------------------------------------------------
    def _to_train(self, x):
        """[S, ...] simulator order -> [P * B, ...] with policy p's rows in block p (_sim_to_train)."""
        if self.reorder is None:
            return x
        g = self.reorder.to_policy(x)                      # [P + extra chunks, B, ...]; chunks 0..P-1 are full
        return g[:self.P].reshape(self.S, *x.shape[1:])

    def _blocks(self, out):
        obs = {k: self._to_train(v) for k, v in out['obs'].items()}
        rew, don = self._to_train(out['rewards']), self._to_train(out['dones'])
        B = self.B
        return [{'state': None, 'obs': {k: v[p * B:(p + 1) * B] for k, v in obs.items()},
                 'rewards': rew[p * B:(p + 1) * B], 'dones': don[p * B:(p + 1) * B]} for p in range(self.P)]

    def _update(self):
        subs, B = self.subs, self.B
        from .ppo import hoist_permutations
        for m in subs:
            hoist_permutations(m.state.train_states, m.ppo_ws)       # every policy's own key / side stream
            m.rollout, m.state.user_state = m.hooks.start_rollouts(m.rollout, m.state.user_state)
        C = subs[0].rollout_mgr._num_bptt_chunks
        Tp = subs[0].rollout_mgr._num_bptt_steps
        for c in range(C):
            for m in subs:
                m.rollout_mgr.begin_chunk(m.rollout, c)
            for s in range(Tp):
                for p, m in enumerate(subs):
                    a = m.rollout_mgr.policy_step(m.rollout, m.state.policy_states, c, s)
                    dst = self._train_actions if self.reorder is not None else self.actions
                    call('mlb_copy_bytes', ptr(a), ptr(dst[p * B:(p + 1) * B]), c_size_t(a.numel() * 4))
                if self.reorder is not None:               # training order -> simulator order
                    acts = self.reorder.to_sim(self._train_actions.view(self.P, B, -1))
                    call('mlb_copy_bytes', ptr(acts), ptr(self.actions), c_size_t(acts.numel() * 4))
                out = self._step_fn({
                    'state': self._sim_state,
                    'actions': {self._act_name: self.actions.view(torch.float32) if self._continuous else self.actions},
                    'resets': self.resets, 'sim_ctrl': subs[0].rollout.sim_ctrl,
                    'pbt': {'policy_assignments': self.policy_assignments}})
                self._sim_state = out['state']
                for m, blk in zip(subs, self._blocks(out)):
                    m.rollout_mgr.post_step(m.rollout, c, s, blk)
            for m in subs:
                m.rollout_mgr.end_chunk(m.rollout)
        from .train import _learn_impl
        for m in subs:                                     # the vmap over policies of ml/train.py:165-174
            collected = m.rollout_mgr.finish(m.state, m.rollout, m.metrics, m.hooks.finish_rollouts,
                                             m.hooks.rollout_metrics)
            m.state, m.rollout, m.metrics = _learn_impl(m.algo, m.cfg, m.hooks, collected, None, m.ppo_ws)

    def update_iter(self):
        if self.use_cuda_graph and self._graph is None and self._eager_iters >= 1:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._update()
            self._graph = g
        if self._graph is not None:
            self._graph.replay()
        else:
            self._update()
            self._eager_iters += 1
        for m in self.subs:
            m.metrics.advance()
            m.update_idx += 1
        self.update_idx += 1
        return self

    def save_ckpt(self, path):
        for p, m in enumerate(self.subs):
            m.state.save(int(self.update_idx), os.path.join(path, str(int(self.update_idx)), f'policy_{p}'))

    def log_metrics_tensorboard(self, tb_writer):
        for m in self.subs:
            m.metrics.tensorboard_log(self.update_idx - 1, tb_writer)


def init_multi_policy_training(dev, cfg, sim_fns, policy, init_sim_ctrl, user_hooks, restore_ckpt, profile_port,
                               dist_ctx, assignments=None):
    from .train import _init_training, _key_from_seed
    P = _check_pbt(cfg)
    if dist_ctx is not None:
        raise NotImplementedError('multi-policy learner + data parallelism')
    if restore_ckpt is not None:
        raise NotImplementedError('multi-policy checkpoint restore')
    sim_init = sim_fns['init']()
    S = cfg.num_worlds * cfg.num_agents_per_world
    B = S // P
    # per-policy keys: split(key(seed), P) (the reference splits its init / rollout keys over the policy axis)
    seed_key = _key_from_seed(cfg.seed, dev) if isinstance(cfg.seed, int) else cfg.seed.to(dev)
    keys = K.threefry_split(seed_key, P) if P > 1 else seed_key.view(1, 2)
    tmp = MultiPolicyTrainingManager.__new__(MultiPolicyTrainingManager)
    tmp.P, tmp.S, tmp.B, tmp.reorder = P, S, B, None
    if assignments is not None:
        a = assignments.to(device=dev, dtype=torch.int32).contiguous()
        tmp.reorder = reorder_state_for(a, P, B)
    obs0 = {k: tmp._to_train(v) for k, v in sim_init['obs'].items()}
    subs = []
    for p in range(P):
        blk = {k: v[p * B:(p + 1) * B] for k, v in obs0.items()}

        def no_step(_):
            raise RuntimeError('the multi-policy manager steps the simulator')
        sub_cfg = dataclasses.replace(cfg, pbt=None, num_worlds=cfg.num_worlds // P, seed=keys[p].clone())
        m = _init_training(dev, sub_cfg, {'init': (lambda blk=blk: {'state': None, 'obs': blk}), 'step': no_step},
                           policy, init_sim_ctrl, user_hooks, None, profile_port, None)
        m.hooks, m.algo = user_hooks, sub_cfg.algo.setup()
        subs.append(m)
    return MultiPolicyTrainingManager(subs, cfg, sim_fns, sim_init, assignments, dev)
