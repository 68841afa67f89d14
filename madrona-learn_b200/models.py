"""Model-zoo descriptors for the hot path (ml/models.py:46-174).

The reference's modules are flax `nn.Module`s traced by XLA.  Here they are declarative
descriptors: `engine.PolicyProgram` lowers an `ActorCritic` built from them onto the
hand-written kernels (Dense -> LayerNorm -> ReLU stacks, fused actor+critic head GEMM,
sampling / loss epilogues).  Constructor arguments keep the reference's names.
"""
from dataclasses import dataclass
from typing import Any, Optional

import numpy as np
import torch

from .cfg import DiscreteActionsConfig


@dataclass(frozen=True)
class LayerNorm:                         # ml/models.py:46-56 (flax nn.LayerNorm: eps 1e-6)
    dtype: Any = torch.float32
    use_ref: bool = True


@dataclass(frozen=True)
class MLP:                               # ml/models.py:99-119
    """num_layers x [Dense(num_channels, no bias, orthogonal(sqrt 2)) -> LayerNorm -> ReLU].
    Parameter paths: Dense_i/kernel [in, H]; LayerNorm_i/impl/{scale,bias} [H]."""
    num_channels: int
    num_layers: int
    dtype: Any = torch.float32
    weight_init_scale: float = float(np.sqrt(2))


@dataclass(frozen=True)
class DenseLayerDiscreteActor:           # ml/models.py:122-139
    """Dense(sum(buckets), bias, orthogonal(0.01)) -> DiscreteActionDistributions.
    Parameter paths: impl/kernel [F, sumA], impl/bias [sumA]."""
    cfg: DiscreteActionsConfig
    dtype: Any = torch.float32
    weight_init_scale: float = 0.01


@dataclass(frozen=True)
class DenseLayerCritic:                  # ml/models.py:142-154
    """Dense(1, bias, orthogonal(1.0)) cast to f32.  Paths: Dense_0/kernel [F,1], Dense_0/bias."""
    dtype: Any = torch.float32
    weight_init_scale: float = 1.0


@dataclass(frozen=True)
class DreamerV3Critic:                   # ml/models.py:157-174 (SURVEY 8f rank 3 -- "next")
    dtype: Any = torch.float32
    num_bins: int = 63


@dataclass(frozen=True)
class HLGaussCritic:                     # ml/models.py:253-306
    """Dense(num_bins, bias, zero init) -> HLGaussDist: logits over linearly spaced bin centres; value =
    softmax-weighted centres (symmetric summation), loss = cross-entropy against the histogram of a
    Gaussian around the return (sigma = smoothness * bin width).  Build with `create`, like the reference."""
    dtype: Any = torch.float32
    centers: Any = None
    bounds: Any = None
    smoothness: float = 0.75

    @staticmethod
    def create(dtype=torch.float32, num_bins: int = 127, min_bound=-100, max_bound=100, smoothness: float = 0.75):
        import numpy as np
        half = np.linspace(min_bound, 0, num_bins // 2 + 1)           # gen_bins (ml/models.py:271-283)
        bins = np.concatenate([half, -half[:-1][::-1]], axis=0)
        width = bins[1] - bins[0]
        bounds = bins - 0.5 * width
        bounds = np.concatenate([bounds, np.asarray([bounds[-1] + width])], axis=0)
        return HLGaussCritic(dtype=dtype, centers=bins.astype(np.float32), bounds=bounds.astype(np.float32),
                             smoothness=float(smoothness))

    @property
    def num_bins(self):
        return int(self.centers.shape[0])


@dataclass(frozen=True)
class DenseLayerContinuousActor:         # (ours) the continuous counterpart of DenseLayerDiscreteActor
    """Dense(2 * num_dims, bias, orthogonal(0.01)) -> ContinuousActionDistributions (ml/dists.py:211-284):
    columns [0, n) are the raw means (tanh'ed), [n, 2n) the raw stds ((max - min) * sigmoid(raw + 2) + min).
    The reference ships the distribution class but no actor module for it (users write their own head); this is
    the Dense head with the discrete actor's initialisation.  Paths: impl/kernel [F, 2n], impl/bias [2n]."""
    cfg: Any
    dtype: Any = torch.float32
    weight_init_scale: float = 0.01
