"""Observation preprocessing (ml/observations.py:13-158): Noop, Caster and the EMA normaliser
(normalise with the K3 kernels; per-step batch moments + the equal-weight Chan merge over the steps
of an update in select.cu; the EMA update itself is mlb_ema_update_f32)."""
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

    def update_state(self, states, o_stats, vmap, **kw):
        return states

    def init_obs_stats(self, states, vmap, **kw):
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


class ObsStats:
    """Per-update raw observation moments of one observation tensor: f64 [T, D, 2] = per step,
    per feature {sum, sum of squares} over the step's rows (what EMANormalizer.update_input_stats
    reduces, ml/moving_avg.py:103-129, kept un-merged so a data-parallel run can SUM-all-reduce
    it once per update); `count` = rows per step."""

    def __init__(self, raw, count):
        self.raw, self.count = raw, count


@dataclass(frozen=True)
class ObservationsEMANormalizer(ObservationsPreprocess):      # :70-132
    normalizer: Any = None
    prep_fns: Any = None
    skip_normalization: Any = None

    @staticmethod
    def create(decay, dtype, eps=1e-5, prep_fns=None, skip_normalization=None):
        from .moving_avg import EMANormalizer
        if dtype not in (torch.float32,):
            raise NotImplementedError('ObservationsEMANormalizer: float32 statistics only')
        return ObservationsEMANormalizer(
            normalizer=EMANormalizer(decay=decay, norm_dtype=dtype, inv_dtype=dtype, eps=eps),
            prep_fns=dict(prep_fns or {}), skip_normalization=frozenset(skip_normalization or ()))

    def _prep_ob(self, ob_name, ob):
        fn = self.prep_fns.get(ob_name)
        return ob if fn is None else fn(ob)

    def _skip(self, ob_name):
        return ob_name in self.skip_normalization

    # -- state: one EMA-normaliser state tensor per observation (None when skipped) --------------
    def init_state(self, obs, vmap):                          # :104-109
        out = {}
        for k, ob in obs.items():
            out[k] = None if self._skip(k) else self.normalizer.init_estimates(self._prep_ob(k, ob))
        return out

    def preprocess(self, states, obs, vmap):                  # :97-103
        res = {}
        for k, ob in obs.items():
            ob = self._prep_ob(k, ob)
            if self._skip(k) or states is None or states.get(k) is None:
                res[k] = ob
                continue
            key = (k, tuple(ob.shape), ob.device)
            buf = _SCRATCH.get(key)
            if buf is None:
                buf = _SCRATCH[key] = torch.empty_like(ob, dtype=torch.float32)
            res[k] = self.normalizer.normalize(states[k], ob, out=buf)
        return res

    # -- per-update statistics ---------------------------------------------------------------------
    def init_obs_stats(self, states, vmap, num_steps=None, example_obs=None):    # :116-120
        """num_steps / example_obs: the steps of one update and the raw observations (allocation
        sizes; the reference gets them from tracing)."""
        out = {}
        for k, st in (states or {}).items():
            if st is None or num_steps is None:
                out[k] = None
                continue
            ob = self._prep_ob(k, example_obs[k])
            D = ob.shape[-1]
            out[k] = ObsStats(torch.zeros(num_steps, D, 2, dtype=torch.float64, device=ob.device),
                              ob.numel() // D)
        return out

    def update_obs_stats(self, states, cur_obs_stats, num_prev_updates, obs, vmap):   # :122-132
        from . import kernels as K
        for k, stats in (cur_obs_stats or {}).items():
            if stats is None:
                continue
            ob = self._prep_ob(k, obs[k])
            K.obs_moments(ob, stats.raw[num_prev_updates])
        return cur_obs_stats

    def update_state(self, states, o_stats, vmap, dist_ctx=None):              # :110-115
        from . import kernels as K
        for k, stats in (o_stats or {}).items():
            if stats is None or states.get(k) is None:
                continue
            count = stats.count
            if dist_ctx is not None:                 # global batch statistics: ONE all-reduce per update
                dist_ctx.allreduce_sum(stats.raw)
                count *= dist_ctx.world_size
            mean, var = K.obs_stats_merge(stats.raw, count)
            self.normalizer.update_estimates(states[k], (mean, var))
        return states


_SCRATCH = {}
