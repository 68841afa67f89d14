"""Observation preprocessing (ml/observations.py:13-158).  Noop and Caster are lowered; the
EMA normaliser variant reuses the K3 kernels ("next": not wired into the rollout yet)."""
from dataclasses import dataclass
from typing import Any

import torch


@dataclass(frozen=True)
class ObservationsPreprocess:
    def preprocess(self, states, obs, vmap):
        return {k: self._preprocess(k, None if states is None else states.get(k), v)
                for k, v in obs.items()}

    def init_state(self, obs, vmap):
        return {k: None for k in obs}

    def update_state(self, states, o_stats, vmap):
        return states

    def init_obs_stats(self, states, vmap):
        return {k: None for k in (states or {})}

    def update_obs_stats(self, states, cur_obs_stats, num_prev_updates, obs, vmap):
        return cur_obs_stats

    def _preprocess(self, ob_name, state, ob):
        return ob


@dataclass(frozen=True)
class ObservationsPreprocessNoop(ObservationsPreprocess):     # :151-158
    @staticmethod
    def create():
        return ObservationsPreprocessNoop()


@dataclass(frozen=True)
class ObservationsCaster(ObservationsPreprocess):             # :135-148
    dtype: Any = torch.float32

    @staticmethod
    def create(dtype):
        return ObservationsCaster(dtype=dtype)

    def _preprocess(self, ob_name, state, ob):
        return ob if ob.dtype == self.dtype else ob.to(self.dtype)


@dataclass(frozen=True)
class ObservationsEMANormalizer(ObservationsPreprocess):      # :70-132
    decay: float = 0.99999
    dtype: Any = torch.float32
    eps: float = 1e-5

    @staticmethod
    def create(decay, dtype, eps=1e-5, prep_fns=None, skip_normalization=None):
        raise NotImplementedError('ObservationsEMANormalizer is a "next" row (SURVEY 8f); '
                                  'use ObservationsPreprocessNoop / ObservationsCaster')
