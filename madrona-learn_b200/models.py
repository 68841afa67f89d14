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
