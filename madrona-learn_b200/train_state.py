"""Train / policy state containers (ml/train_state.py:34-488).  Field names follow the
reference; leaves are device tensors.  P (number of train policies) is 1: PBT is out of scope.
"""
import os
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, Optional

import torch

from . import kernels as K
from .algo_common import HyperParams
from .engine import PolicyProgram
from .moving_avg import EMAEstimate, EMANormalizer
from .observations import ObservationsPreprocessNoop


@dataclass
class PolicyState:                                     # ml/train_state.py:34-82
    apply_fn: Callable
    rnn_reset_fn: Callable
    params: Dict[str, Any]
    batch_stats: Dict[str, Any]
    obs_preprocess: Any
    obs_preprocess_state: Any
    reward_hyper_params: Any
    get_episode_scores_fn: Callable
    episode_score: Any
    mmr: Any
    program: PolicyProgram = None

    def update(self, **kw):
        for k, v in kw.items():
            if v is not None:
                setattr(self, k, v)
        return self


@dataclass
class PolicyTrainState:                                # ml/train_state.py:85-136
    value_normalizer: Optional[EMANormalizer]
    max_advantage_est: EMAEstimate
    initial_weight_norms: Dict[str, Any]
    tx: Any
    value_normalizer_state: Any
    max_advantage_est_state: Any
    hyper_params: HyperParams
    opt_state: Dict[str, Any]
    scheduler: Any
    scaler: Any
    update_prng_key: torch.Tensor          # int32[2] device tensor holding the uint32 key words

    def update(self, **kw):
        for k, v in kw.items():
            if v is not None:
                setattr(self, k, v)
        return self

    def gen_update_rnd(self, partitionable=False):     # :134-136
        ks = K.threefry_split(self.update_prng_key, 2, partitionable)
        self.update_prng_key.copy_(ks[1])
        return ks[0], self


def _tree_to_host(t):
    if isinstance(t, dict):
        return {k: _tree_to_host(v) for k, v in t.items()}
    return t.detach().cpu().contiguous().clone()


@dataclass
class TrainStateManager:                               # ml/train_state.py:139-304
    policy_states: PolicyState
    train_states: PolicyTrainState
    pbt_rng: torch.Tensor
    user_state: Any

    def replace(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)
        return self

    # checkpoint: the reference's key tree (ml/train_state.py:145-164 saves the PolicyState /
    # PolicyTrainState pytrees: :34-47, :85-99) with every leaf a host (CPU) tensor; torch.save stands
    # in for orbax.  Only tensors / scalars / None are written, so load() runs with weights_only=True;
    # `user_state` (arbitrary user pytree) goes to a separate opt-in pickle next to it.
    def save(self, next_update, path):
        ps, ts = self.policy_states, self.train_states
        prog = ps.program
        np_ = lambda t: None if t is None else t.detach().cpu().contiguous().clone()
        obs_state = {k: np_(v) for k, v in (ps.obs_preprocess_state or {}).items()}
        est = ts.max_advantage_est_state
        ckpt = {
            'next_update': int(next_update),
            'policy_states': {
                'params': _tree_to_host(prog.param_tree()),          # flax parameter tree (:34-40)
                'params_flat': np_(prog.params),                     # our arena (what load() restores)
                'batch_stats': {}, 'obs_preprocess_state': obs_state, 'reward_hyper_params': None,
                'episode_score': None, 'mmr': None,
            },
            'train_states': {
                'opt_state': {'m': np_(prog.adam_m), 'v': np_(prog.adam_v), 'count': np_(prog.adam_step)},
                'value_normalizer_state': np_(ts.value_normalizer_state),
                'max_advantage_est_state': np_(est),
                'hyper_params': {k: float(v) for k, v in vars(ts.hyper_params).items()
                                 if isinstance(v, (int, float, bool))},
                'scaler': None,
                'update_prng_key': np_(ts.update_prng_key),
                # per kernel leaf of the parameter tree, None elsewhere (:413-423)
                'initial_weight_norms': prog.initial_weight_norms_tree(),
                'initial_weight_norms_flat': {k: float(v) for k, v in ts.initial_weight_norms.items()},
            },
            'pbt_rng': np_(self.pbt_rng),
            'user_state': None,            # the user pytree itself is in <path>.user_state (opt-in pickle)
        }
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        torch.save(ckpt, path)
        if self.user_state is not None:
            torch.save({'user_state': self.user_state}, path + '.user_state')

    def load(self, path, load_user_state=False):
        ckpt = torch.load(path, map_location='cpu', weights_only=True)
        ps, ts = self.policy_states, self.train_states
        prog = ps.program
        t = lambda a: a if torch.is_tensor(a) else torch.from_numpy(a)
        prog.params.copy_(t(ckpt['policy_states']['params_flat']))
        prog.adam_m.copy_(t(ckpt['train_states']['opt_state']['m']))
        prog.adam_v.copy_(t(ckpt['train_states']['opt_state']['v']))
        prog.adam_step.copy_(t(ckpt['train_states']['opt_state']['count']))
        if ts.value_normalizer_state is not None:
            ts.value_normalizer_state.copy_(t(ckpt['train_states']['value_normalizer_state']))
        if ts.max_advantage_est_state is not None and ckpt['train_states'].get('max_advantage_est_state') is not None:
            ts.max_advantage_est_state.copy_(t(ckpt['train_states']['max_advantage_est_state']))
        for k, v in (ckpt['policy_states'].get('obs_preprocess_state') or {}).items():
            if v is not None and ps.obs_preprocess_state.get(k) is not None:
                ps.obs_preprocess_state[k].copy_(t(v))
        ts.update_prng_key.copy_(t(ckpt['train_states']['update_prng_key']))
        ts.initial_weight_norms = dict(ckpt['train_states']['initial_weight_norms_flat'])
        # re-derive the device segment table from the restored initial norms
        prog.initial_weight_norms = ts.initial_weight_norms
        prog.rebuild_segments()
        self.pbt_rng.copy_(t(ckpt['pbt_rng']))
        if load_user_state and os.path.exists(path + '.user_state'):       # opt-in: unpickles arbitrary objects
            self.user_state = torch.load(path + '.user_state', map_location='cpu', weights_only=False)['user_state']
        return self, ckpt['next_update']

    @staticmethod
    def create(policy, cfg, algo, init_user_state_cb, base_rng, example_obs, example_rnn_states,
               use_competitive_mmr, device, partitionable=False):
        """ml/train_state.py:278-304 + _make_policies :439-488 (P = 1)."""
        ks = K.threefry_split(base_rng, 2, partitionable)          # base_init_rng, pbt_rng
        base_init, pbt_rng = ks[0].clone(), ks[1].clone()
        ks = K.threefry_split(base_init, 2, partitionable)         # policy_init_base, train_init_base
        policy_init_base, train_init_base = ks[0].clone(), ks[1].clone()
        policy_init = K.threefry_split(policy_init_base, 1, partitionable)[0]
        train_init = K.threefry_split(train_init_base, 1, partitionable)[0].clone()

        obs_preprocess = policy.obs_preprocess or ObservationsPreprocessNoop.create()
        obs_state = obs_preprocess.init_state(example_obs, False)
        pre = obs_preprocess.preprocess(obs_state, example_obs, False)
        if len(pre) != 1:
            raise NotImplementedError('exactly one observation tensor is supported '
                                      '(multi-tensor prefix concat: next)')
        (ob,) = pre.values()
        obs_dim = ob[0].numel()
        prog = PolicyProgram(policy.actor_critic, obs_dim, cfg.actions, device, cfg.compute_dtype)
        seed_words = policy_init.cpu().numpy().view('uint32')
        prog.init_params(int(seed_words[0]) ^ (int(seed_words[1]) << 1))

        hyper = algo.init_hyperparams(cfg)
        tx = algo.make_optimizer(hyper)
        if cfg.normalize_values:
            vn = EMANormalizer(decay=hyper.value_normalizer_decay, norm_dtype=torch.float32,
                               inv_dtype=torch.float32)
            vn_state = K.ema_state_init(1, device)
        else:
            vn, vn_state = None, None

        def apply_fn(variables, *args, train=False, method='rollout', **kw):
            return getattr(prog, 'apply_' + method)(*args, train=train, **kw)

        ps = PolicyState(
            apply_fn=apply_fn, rnn_reset_fn=lambda states, dones: states,
            params=prog.param_tree(), batch_stats={}, obs_preprocess=obs_preprocess,
            obs_preprocess_state=obs_state, reward_hyper_params=None,
            get_episode_scores_fn=policy.get_episode_scores or (lambda x: 0.0),
            episode_score=None, mmr=None, program=prog)
        max_adv_est = EMAEstimate(hyper.max_advantage_est_decay)        # ml/train_state.py:407-411
        ts = PolicyTrainState(
            value_normalizer=vn, max_advantage_est=max_adv_est,
            initial_weight_norms=dict(prog.initial_weight_norms), tx=tx,
            value_normalizer_state=vn_state,
            max_advantage_est_state=max_adv_est.init_estimates(prog.params), hyper_params=hyper,
            opt_state={'m': prog.adam_m, 'v': prog.adam_v, 'count': prog.adam_step},
            scheduler=None, scaler=None, update_prng_key=train_init)
        return TrainStateManager(policy_states=ps, train_states=ts, pbt_rng=pbt_rng,
                                 user_state=init_user_state_cb())
