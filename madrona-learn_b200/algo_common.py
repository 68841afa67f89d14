"""Algorithm plugin interface + GAE / returns / z-score entry points
(ml/algo_common.py:15-140), executed by the K1/K2 kernels."""
from dataclasses import dataclass

import torch

from . import kernels as K


@dataclass
class HyperParams:                                     # ml/algo_common.py:15-21
    lr: float
    gamma: float
    gae_lambda: float
    normalize_values: bool
    value_normalizer_decay: float
    max_advantage_est_decay: float


class AlgoBase:                                        # ml/algo_common.py:24-42
    def init_hyperparams(self, cfg):
        raise NotImplementedError

    def make_optimizer(self, hyper_params):
        raise NotImplementedError

    def update(self, *args, **kwargs):
        raise NotImplementedError

    def add_metrics(self, cfg, metrics):
        raise NotImplementedError


def _tn(x):
    """[C, T', P, B, 1] -> T, N (plain reshape, ml/algo_common.py:91-98)."""
    if x.dim() >= 4:
        T = x.shape[0] * x.shape[1]
    else:
        T = x.shape[0]
    return T, x.numel() // T


def compute_advantages(cfg, rewards, values, dones, bootstrap_values, advantages=None,
                       returns=None, vn_mu_sigma=None, metrics=None, ws=None):
    """ml/algo_common.py:84-130 (+ fused returns = advantages + values, ml/rollouts.py:769).
    Returns (advantages, returns) shaped like `rewards`."""
    T, N = _tn(rewards)
    adv, ret = K.gae(rewards.view(T, N), values.view(T, N), dones.view(T, N),
                     bootstrap_values.reshape(N), cfg.gamma, cfg.gae_lambda,
                     advantages=None if advantages is None else advantages.view(T, N),
                     returns=None if returns is None else returns.view(T, N),
                     vn_mu_sigma=vn_mu_sigma, metrics=metrics, ws=ws)
    return adv.view(rewards.shape), ret.view(rewards.shape)


def compute_returns(cfg, rewards, dones, bootstrap_values, returns=None):
    """ml/algo_common.py:45-81."""
    T, N = _tn(rewards)
    ret = K.discounted_returns(rewards.view(T, N), dones.view(T, N), bootstrap_values.reshape(N),
                               cfg.gamma, None if returns is None else returns.view(T, N))
    return ret.view(rewards.shape)


def zscore_data(data, out=None):
    """ml/algo_common.py:133-140."""
    return K.zscore(data, out)
