"""madrona-learn learner hot path, B200-native (sm_100a kernels behind the reference's API).

Public surface mirrors /root/reference/src/madrona_learn/__init__.py:1-45 for the hot path.
"""
from . import _lib  # noqa: F401
from . import kernels  # noqa: F401
from . import models, rnn  # noqa: F401
from .actor_critic import (ActorCritic, Backbone, BackboneEncoder, BackboneSeparate,  # noqa: F401
                           BackboneShared, RecurrentBackboneEncoder)
from .cfg import (ContinuousActionsConfig, DiscreteActionsConfig, EvalConfig, ParamExplore,  # noqa: F401
                  PBTConfig, TrainConfig)
from .engine import matmul_precision, set_matmul_precision  # noqa: F401
from .envs import HostTraceEnv, SyntheticVectorEnv  # noqa: F401
from .moving_avg import EMANormalizer  # noqa: F401
from . import pbt_reorder  # noqa: F401
from .observations import ObservationsCaster, ObservationsEMANormalizer  # noqa: F401
from .policy import Policy  # noqa: F401
from .ppo import PPOConfig  # noqa: F401
from .profile import profile  # noqa: F401
from .train import TrainHooks, TrainingManager, init_training, stop_training  # noqa: F401
from .train_state import TrainStateManager  # noqa: F401
from .utils import aot_compile, cfg_jax_mem  # noqa: F401
