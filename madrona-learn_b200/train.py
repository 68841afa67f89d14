"""Train loop / orchestration (ml/train.py:35-391).

Public entry points keep the reference's names and arguments:
    init_training(dev, cfg, sim_fns, policy, init_sim_ctrl, user_hooks, restore_ckpt,
                  profile_port) -> TrainingManager
    TrainingManager.update_iter() -> TrainingManager
`dev` is a torch CUDA device (one process per GPU; torch.distributed initialised by the
caller makes the run data-parallel over worlds, see parallel.py).

Where the reference relies on `aot_compile` (jit + donate, ml/utils.py:42-57) to turn
update_iter into one XLA executable, the B200 path captures the whole update -- T rollout
steps, bootstrap, GAE, every minibatch step and the NCCL all-reduces -- into ONE CUDA graph
on the second call and replays it afterwards (all buffers are static, all randomness lives in
device-side threefry keys).  Set MLB_CUDA_GRAPH=0 to run eagerly.
"""
import os
from dataclasses import dataclass
from typing import Any, Callable, Dict, Optional

import torch

from . import kernels as K
from .cfg import TrainConfig
from .metrics import TrainingMetrics
from .policy import Policy
from .ppo import _PPOWorkspace, hoist_permutations
from .profile import profile
from .rollouts import RolloutConfig, RolloutManager, RolloutState
from .train_state import TrainStateManager


@dataclass(frozen=True)
class TrainHooks:                                      # ml/train.py:75-128 -- must be stateless
    def init_user_state(self):
        return None

    def start_rollouts(self, rollout_state, user_state):
        return rollout_state, user_state

    def finish_rollouts(self, rollouts, bootstrap_values, unnormalized_values,
                        unnormalized_bootstrap_values, user_state):
        return rollouts, user_state

    def add_metrics(self, metrics):
        return metrics

    def rollout_metrics(self, metrics, rollouts, user_state):
        return metrics

    def optimize_metrics(self, metrics, epoch_idx, minibatch, policy_state, train_state):
        return metrics


class TrainingManager:                                 # ml/train.py:35-64
    def __init__(self, state, rollout, metrics, update_idx, cfg, update_fn, profile_port):
        self.state, self.rollout, self.metrics = state, rollout, metrics
        self.update_idx, self.cfg, self.update_fn, self.profile_port = update_idx, cfg, update_fn, profile_port
        self._graph = None
        self._eager_iters = 0
        self.use_cuda_graph = os.environ.get('MLB_CUDA_GRAPH', '1') != '0'
        self.kernel_launches_per_update = None

    def save_ckpt(self, path):
        self.state.save(int(self.update_idx), os.path.join(path, str(int(self.update_idx))))

    def load_ckpt(self, path):
        self.state, self.update_idx = self.state.load(path)
        self._graph = None
        return self

    def update_iter(self):
        """One rollout + GAE + PPO update (functional in the reference; in place here -- the
        returned manager is `self`, mirroring aot_compile's donate-all)."""
        if self.use_cuda_graph and self._graph is None and self._eager_iters >= 1:
            self._capture()
        if self._graph is not None:
            self._graph.replay()
        else:
            self.update_fn(self.state, self.rollout, self.metrics, self.update_idx)
            self._eager_iters += 1
        self.metrics.advance()
        self.update_idx += 1
        return self

    def _capture(self):
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.update_fn(self.state, self.rollout, self.metrics, self.update_idx)
        self._graph = g

    def log_metrics_tensorboard(self, tb_writer):        # ml/train.py:62-64
        self.metrics.tensorboard_log(self.update_idx - 1, tb_writer)


def init_training(dev, cfg: TrainConfig, sim_fns: Dict[str, Callable], policy: Policy,
                  init_sim_ctrl, user_hooks: TrainHooks = TrainHooks(), restore_ckpt: str = None,
                  profile_port: int = None, dist_ctx=None, verbose=True) -> TrainingManager:
    """ml/train.py:131-146."""
    if verbose:
        print(cfg)
        print()
    dev = torch.device(dev)
    if dev.type != 'cuda':
        raise RuntimeError('madrona_learn_b200 runs on CUDA devices only (no CPU fallback)')
    if cfg.pbt is not None:
        from .multi_policy import init_multi_policy_training
        with torch.cuda.device(dev):
            return init_multi_policy_training(dev, cfg, sim_fns, policy, init_sim_ctrl, user_hooks, restore_ckpt,
                                              profile_port, dist_ctx)
    with torch.cuda.device(dev):
        return _init_training(dev, cfg, sim_fns, policy, init_sim_ctrl, user_hooks, restore_ckpt,
                              profile_port, dist_ctx)


def stop_training(training_mgr):                       # ml/train.py:148-153
    torch.cuda.synchronize()


def _setup_rollout_cfg(cfg):                           # ml/train.py:227-266
    sim_batch_size = cfg.num_agents_per_world * cfg.num_worlds
    if cfg.pbt is not None:
        raise NotImplementedError('pbt != None: population training is out of scope (SURVEY 8f)')
    return RolloutConfig.setup(
        num_current_policies=1, num_past_policies=0, num_teams=1,
        team_size=cfg.num_agents_per_world, sim_batch_size=sim_batch_size, actions_cfg=cfg.actions,
        self_play_portion=1.0, cross_play_portion=0.0, past_play_portion=0.0,
        static_play_portion=0.0, reward_gamma=cfg.gamma, custom_policy_ids=cfg.custom_policy_ids,
        policy_dtype=cfg.compute_dtype)


def _key_from_seed(seed, dev):
    """random.key(seed) -> [hi32, lo32] (ml/train.py:284-287)."""
    seed = int(seed)
    words = [(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF]
    t = torch.tensor(words, dtype=torch.int64)
    return torch.where(t >= 2 ** 31, t - 2 ** 32, t).to(torch.int32).to(dev)


def _init_training(dev, cfg, sim_fns, policy, sim_ctrl, user_hooks, restore_ckpt, profile_port,
                   dist_ctx):
    algo = cfg.algo.setup()                                            # :282
    seed = cfg.seed if dist_ctx is None else dist_ctx.rank_seed(cfg.seed)
    seed_key = _key_from_seed(seed, dev) if isinstance(seed, int) else seed.to(dev)
    ks = K.threefry_split(seed_key, 2)                                 # rollout_rng, init_rng
    rollout_rng, init_rng = ks[0].clone(), ks[1].clone()
    if dist_ctx is not None:
        # parameters and the update key stream must be identical on every rank
        init_rng = K.threefry_split(_key_from_seed(cfg.seed, dev), 2)[1].clone()
    rollout_cfg = _setup_rollout_cfg(cfg)
    rnn_states = policy.actor_critic.init_recurrent_state(rollout_cfg.sim_batch_size, dev)
    rollout_state = RolloutState.create(rollout_cfg, sim_fns, rollout_rng, rnn_states, sim_ctrl,
                                        None, dev)
    train_state_mgr = TrainStateManager.create(
        policy=policy, cfg=cfg, algo=algo, init_user_state_cb=user_hooks.init_user_state,
        base_rng=init_rng, example_obs=rollout_state.cur_obs, example_rnn_states=rnn_states,
        use_competitive_mmr=False, device=dev)
    start_update_idx = 0
    if restore_ckpt is not None:
        train_state_mgr, start_update_idx = train_state_mgr.load(restore_ckpt)
    rollout_mgr = RolloutManager(cfg, rollout_state, train_state_mgr.policy_states, dist_ctx)
    # metric table order: rollout metrics first (Rewards, Values, Est Returns, Advantages are
    # consecutive so the GAE kernel can emit them in one go), then the algorithm's, then user's
    names = ['Rewards', 'Values', 'Est Returns']
    if cfg.compute_advantages:
        names.append('Advantages')
    names += ['Env Returns', 'Bootstrap Values']
    names = algo.add_metrics(cfg, names)
    names = user_hooks.add_metrics(names)
    metrics = TrainingMetrics.create(cfg, names, start_update_idx, dev)
    prog = train_state_mgr.policy_states.program
    if dist_ctx is not None and hasattr(dist_ctx, 'enable_fused_allreduce'):
        dist_ctx.enable_fused_allreduce(prog)          # NVLink peer-memory all-reduce when available
    ppo_ws = _PPOWorkspace(cfg, prog, cfg.num_bptt_chunks,
                           cfg.steps_per_update // cfg.num_bptt_chunks,
                           rollout_cfg.sim_batch_size, dist_ctx)

    def update_wrapper(train_state_mgr, rollout_state, metrics, update_idx):
        return _update_impl(algo, cfg, user_hooks, rollout_state, rollout_mgr, train_state_mgr,
                            metrics, update_idx, dist_ctx, ppo_ws)

    mgr = TrainingManager(state=train_state_mgr, rollout=rollout_state, metrics=metrics,
                          update_idx=start_update_idx, cfg=cfg, update_fn=update_wrapper,
                          profile_port=profile_port)
    mgr.rollout_mgr = rollout_mgr
    mgr.ppo_ws = ppo_ws
    if cfg.filter_advantages:
        # the number of minibatches is data dependent and read back every update (ml/ppo.py:399-402):
        # the update is not one static launch sequence, so it is not captured into a CUDA graph
        mgr.use_cuda_graph = False
    return mgr


def _update_impl(algo, cfg, user_hooks, rollout_state, rollout_mgr, train_state_mgr, metrics,
                 update_idx, dist_ctx, ppo_ws):
    """ml/train.py:155-225 (P = 1: the vmap over policies is the identity)."""
    with profile('Update Iter'):
        # the minibatch permutations depend on the update key only: side stream, underneath the rollout phase
        hoist_permutations(train_state_mgr.train_states, ppo_ws, dist_ctx)
        with profile('Collect Rollouts'):
            collected = rollout_mgr.collect(
                train_state_mgr, rollout_state, metrics, user_hooks.start_rollouts,
                user_hooks.finish_rollouts, user_hooks.rollout_metrics)
        return _learn_impl(algo, cfg, user_hooks, collected, dist_ctx, ppo_ws)


def _learn_impl(algo, cfg, user_hooks, collected, dist_ctx, ppo_ws):
    """The part of _update_impl after rollout collection (ml/train.py:193-225): one policy's observation
    statistics + algo.update.  The multi-policy learner (multi_policy.py) calls it once per policy."""
    train_state_mgr, rollout_state, rollout_data, obs_stats, metrics = collected
    with profile('Update Observations Stats'):
        # ml/train.py:193-204: safe right away, the learner only sees preprocessed observations
        ps0 = train_state_mgr.policy_states
        ps0.obs_preprocess_state = ps0.obs_preprocess.update_state(
            ps0.obs_preprocess_state, obs_stats, True, dist_ctx=dist_ctx)
    with profile('Learn'):
        ps, ts, metrics = algo.update(cfg, train_state_mgr.policy_states,
                                      train_state_mgr.train_states, rollout_data,
                                      user_hooks.optimize_metrics, metrics,
                                      dist_ctx=dist_ctx, ws=ppo_ws)
    train_state_mgr.policy_states, train_state_mgr.train_states = ps, ts
    return train_state_mgr, rollout_state, metrics
