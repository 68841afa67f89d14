"""Policy bundle (ml/policy.py:13-17)."""
from dataclasses import dataclass
from typing import Any, Callable, Optional


@dataclass(frozen=True)
class Policy:
    actor_critic: Any
    obs_preprocess: Optional[Any] = None
    get_episode_scores: Optional[Callable] = None
