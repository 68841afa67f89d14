"""EMANormalizer / EMAEstimate facade over the K3 kernels (ml/moving_avg.py:7-198).

State is ONE device tensor: 5*dim float32 (mu | inv_sigma | sigma | mu_biased |
sigma_sq_biased) followed by the int32 counter N -- `state_dict()` exposes it under the
reference's key names.
"""
from dataclasses import dataclass
from typing import Any

import torch

from . import kernels as K


@dataclass(frozen=True)
class EMANormalizer:
    decay: float
    norm_dtype: Any = torch.float32
    inv_dtype: Any = torch.float32
    eps: float = 1e-5
    disable: bool = False

    def init_estimates(self, x):                      # :56-75
        if self.disable:
            return None
        return K.ema_state_init(x.shape[-1], x.device)

    def normalize(self, est, x, out=None):            # :77-85
        if self.disable:
            return x
        return K.ema_normalize(est, x.shape[-1], x, out)

    def invert(self, est, x, out=None):               # :87-95
        if self.disable:
            return x
        return K.ema_invert(est, x.shape[-1], x, out)

    def update_estimates(self, est, input_stats):     # :131-181 (in place: buffers are donated)
        if self.disable:
            return None
        mean, var = input_stats
        return K.ema_update(est, mean.numel(), mean, var, self.decay, self.eps)

    @staticmethod
    def state_dict(est):
        dim = (est.numel() - 1) // 5
        names = ['mu', 'inv_sigma', 'sigma', 'mu_biased', 'sigma_sq_biased']
        d = {n: est[i * dim:(i + 1) * dim] for i, n in enumerate(names)}
        d['N'] = est[5 * dim:].view(torch.int32)
        return d


@dataclass(frozen=True)
class EMAEstimate:                                     # :7-44 (max-abs-advantage tracker of filter_advantages)
    decay: float
    eps: float = 1e-5

    def init_estimates(self, x):                      # :12-20 -- state: [mu, mu_biased | int32 N]
        return K.ema_estimate_init(x.device)

    def update_estimates(self, est, x):               # :22-44 (in place); x: device tensor, its mean is tracked
        if x.numel() != 1:
            x = K.moments(x.reshape(-1).float(), 0.0)[0:1]
        return K.ema_estimate_update(est, x.reshape(1), self.decay)

    @staticmethod
    def state_dict(est):
        return {'mu': est[0:1], 'mu_biased': est[1:2], 'N': est[2:3].view(torch.int32)}
