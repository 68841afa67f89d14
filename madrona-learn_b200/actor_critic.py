"""Actor-critic wrapper descriptors (ml/actor_critic.py:38-303).

`ActorCritic(backbone, actor, critic)` keeps the reference's constructor; its four methods
(`rollout`, `update`, `critic_only`, `actor_only`) are executed by `engine.PolicyProgram`,
reachable through `PolicyState.apply_fn(..., method=...)` exactly like flax's `apply`.
"""
from dataclasses import dataclass
from typing import Any, Callable, Optional, Union


class Backbone:
    pass


@dataclass(frozen=True)
class BackboneEncoder:                    # ml/actor_critic.py:131-153
    net: Any

    def init_recurrent_state(self, N, device='cuda'):
        return ()


@dataclass(frozen=True)
class RecurrentBackboneEncoder:           # ml/actor_critic.py:156-199
    net: Any
    rnn: Any

    def init_recurrent_state(self, N, device='cuda'):
        return self.rnn.init_recurrent_state(N, device)


@dataclass(frozen=True)
class BackboneShared(Backbone):           # ml/actor_critic.py:202-244
    prefix: Optional[Union[Callable, Any]]
    encoder: Any

    def init_recurrent_state(self, N, device='cuda'):
        return self.encoder.init_recurrent_state(N, device)


@dataclass(frozen=True)
class BackboneSeparate(Backbone):         # ml/actor_critic.py:247-303
    prefix: Optional[Union[Callable, Any]]
    actor_encoder: Any
    critic_encoder: Any

    def init_recurrent_state(self, N, device='cuda'):
        return (self.actor_encoder.init_recurrent_state(N, device),
                self.critic_encoder.init_recurrent_state(N, device))


@dataclass(frozen=True)
class ActorCritic:                        # ml/actor_critic.py:38-128
    backbone: Backbone
    actor: Any
    critic: Any

    def init_recurrent_state(self, N, device='cuda'):
        return self.backbone.init_recurrent_state(N, device)
