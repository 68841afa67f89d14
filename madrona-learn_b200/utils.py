"""Small helpers mirroring ml/utils.py:14-57."""
from dataclasses import dataclass
from typing import Tuple

import torch


@dataclass
class TypedShape:                 # ml/utils.py:14-17
    shape: Tuple[int, ...]
    dtype: torch.dtype


def cfg_jax_mem(mem_fraction):
    """ml/utils.py:21-24 sets XLA's allocator fraction; here: torch's caching allocator cap."""
    if torch.cuda.is_available():
        torch.cuda.set_per_process_memory_fraction(float(mem_fraction))


def symlog(x):                    # ml/utils.py:36-37
    return torch.sign(x) * torch.log1p(torch.abs(x))


def symexp(x):                    # ml/utils.py:39-40
    return torch.sign(x) * torch.expm1(torch.abs(x))


def aot_compile(func, *args):
    """ml/utils.py:42-57 jit+donate.  The B200 path captures update_iter as a CUDA graph
    inside TrainingManager itself, so this just returns the callable unchanged."""
    return func
