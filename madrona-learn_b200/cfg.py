"""Configuration dataclasses -- field names and defaults follow the reference verbatim
(ml/cfg.py:9-142) because they ARE the user-facing API of the path.  `compute_dtype` takes a
torch dtype (torch.float32 -> fp32 activations with the Dense products on tcgen05 kind::tf32 (exact FFMA under
set_matmul_precision("highest")), torch.bfloat16 -> fused tcgen05 tensor-core path).
"""
import dataclasses
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Union

import torch


@dataclass(frozen=True)
class DiscreteActionsConfig:            # ml/cfg.py:9-11
    actions_num_buckets: List[int]


@dataclass(frozen=True)
class ContinuousActionsConfig:          # ml/cfg.py:13-17
    stddev_min: float
    stddev_max: float
    num_dims: int


class AlgoConfig:                       # ml/cfg.py:19-24 -- the algorithm plugin seam
    def name(self):
        raise NotImplementedError

    def setup(self):
        raise NotImplementedError


@dataclass(frozen=True)
class ParamExplore:                     # ml/cfg.py:27-46 (PBT hyper-parameter exploration)
    base: float
    min_scale: float
    max_scale: float
    log10_scale: bool = False
    ln_scale: bool = False
    clip_perturb: bool = False
    perturb_rnd_min: float = 0.8
    perturb_rnd_max: float = 1.2


@dataclass(frozen=True)
class PBTConfig:                        # ml/cfg.py:49-65 -- accepted for API parity; the PBT
    num_teams: int                      # branches are out of scope
    team_size: int
    num_train_policies: int
    num_past_policies: int
    self_play_portion: float
    cross_play_portion: float
    past_play_portion: float
    policy_overwrite_threshold: float = 0.7
    reward_hyper_params_explore: Dict[str, ParamExplore] = field(default_factory=dict)
    rollout_policy_chunk_size_override: int = 0


@dataclass(frozen=True)
class TrainConfig:                      # ml/cfg.py:68-96
    num_worlds: int
    num_agents_per_world: int
    num_updates: int
    actions: Dict[str, Union[DiscreteActionsConfig, ContinuousActionsConfig]]
    steps_per_update: int
    lr: Union[float, ParamExplore]
    algo: AlgoConfig
    num_bptt_chunks: int
    gamma: float
    seed: int
    metrics_buffer_size: int
    baseline_policy_id: int = 0
    custom_policy_ids: List[int] = field(default_factory=list)
    gae_lambda: float = 1.0
    pbt: Optional[PBTConfig] = None
    dreamer_v3_critic: bool = True
    hlgauss_critic: bool = False
    compute_advantages: bool = True
    normalize_advantages: bool = True   # only used if compute_advantages
    normalize_returns: bool = True      # only used if not compute_advantages
    normalize_values: bool = False
    filter_advantages: bool = False
    importance_sample_trajectories: bool = False
    importance_sample_num_minibatches: int = 0
    value_normalizer_decay: float = 0.99999
    max_advantage_est_decay: float = 0.99999
    compute_dtype: torch.dtype = torch.float32

    def __repr__(self):
        lines = ['TrainConfig:']
        for f in dataclasses.fields(self):
            v = getattr(self, f.name)
            if f.name == 'algo':
                lines.append(f'  {v.name()}:')
                lines += [f'    {k}: {x}' for k, x in vars(v).items()]
            elif f.name == 'pbt':
                lines.append('  pbt: Disabled' if v is None else f'  pbt: {v}')
            elif f.name == 'compute_dtype':
                lines.append('  compute_dtype: ' + {torch.float32: 'fp32', torch.float16: 'fp16',
                                                    torch.bfloat16: 'bf16'}.get(v, str(v)))
            else:
                lines.append(f'  {f.name}: {v}')
        return '\n'.join(lines)


@dataclass(frozen=True)
class EvalConfig:                       # ml/cfg.py:130-142 (eval is out of scope; kept for API)
    num_worlds: int
    num_teams: int
    team_size: int
    num_eval_steps: int
    actions: Dict[str, Union[DiscreteActionsConfig, ContinuousActionsConfig]]
    reward_gamma: float
    policy_dtype: torch.dtype
    eval_competitive: bool
    use_deterministic_policy: bool = True
    clear_fitness: bool = True
    custom_policy_ids: List[int] = field(default_factory=list)
