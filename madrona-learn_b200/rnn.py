"""Recurrent descriptors (ml/rnn.py:47-111)."""
from dataclasses import dataclass
from typing import Any

import torch

__all__ = ['LSTM']


@dataclass(frozen=True)
class LSTM:
    """Multi-layer LSTM over flax OptimizedLSTMCell semantics (gates i,f,g,o; input kernels
    without bias, hidden kernels with bias; output = concat of every layer's h)."""
    num_hidden_channels: int
    num_layers: int
    dtype: Any = torch.float32

    def init_recurrent_state(self, N, device='cuda'):          # ml/rnn.py:52-63
        z = lambda: torch.zeros((N, self.num_hidden_channels), dtype=torch.float32, device=device)
        return [z() for _ in range(self.num_layers)], [z() for _ in range(self.num_layers)]
