"""`profile(name)` scopes (ml/profile.py:6-32) as NVTX ranges, same scope names as the
reference ("Update Iter", "Collect Rollouts", "Policy Inference", ...)."""
import contextlib
import os

import torch

_ENABLED = os.environ.get('MLB_NVTX', '0') == '1'


@contextlib.contextmanager
def profile(name, **kwargs):
    if _ENABLED and torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield
