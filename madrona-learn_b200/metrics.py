"""Device-side Welford metric records in a ring buffer (ml/metrics.py:12-188).

Each metric is a row of `mlb_metric` (mean, m2, min, max f32 + count i32 = 20 bytes) in a
device uint8 table [metrics_buffer_size, num_metrics, 20].  Kernels write the CURRENT
update's records into a fixed staging row (so the update stays CUDA-graph-capturable);
`advance()` copies staging into the ring slot `cur_buffer_offset` and bumps the offset, which
reproduces "every record() overwrites slot cur_buffer_offset; advance() once per update"
(ml/metrics.py:161-188, ml/train.py:223).
"""
import ctypes
from typing import Dict, List

import numpy as np
import torch

from . import _lib
from ._lib import c_size_t, call, ptr

REC = ctypes.sizeof(_lib.Metric)


class Metric:
    """Host view of one record (ml/metrics.py:12-18)."""

    def __init__(self, per_policy=True, mean=0.0, m2=0.0, min=np.finfo(np.float32).max,
                 max=np.finfo(np.float32).min, count=0):
        self.per_policy, self.mean, self.m2, self.min, self.max, self.count = \
            per_policy, mean, m2, min, max, count

    @staticmethod
    def init(per_policy):
        return Metric(per_policy)

    def __repr__(self):
        return (f'Metric(mean={self.mean:.6g}, m2={self.m2:.6g}, min={self.min:.6g}, '
                f'max={self.max:.6g}, count={self.count})')


class TrainingMetrics:
    def __init__(self, names: List[str], buffer_size: int, start_update_idx: int, device):
        self.names = list(names)
        self.index = {n: i for i, n in enumerate(self.names)}
        self.update_buffer_size = int(buffer_size)
        self.update_idx = int(start_update_idx)
        self.cur_buffer_offset = 0
        self.device = device
        n = len(self.names)
        self.staging = torch.zeros(n * REC, dtype=torch.uint8, device=device)
        self.ring = torch.zeros(self.update_buffer_size * n * REC, dtype=torch.uint8, device=device)
        init = _lib.Metric(0.0, 0.0, np.finfo(np.float32).max, np.finfo(np.float32).min, 0)
        row = np.frombuffer(bytes(init) * n, dtype=np.uint8).copy()
        self.staging.copy_(torch.from_numpy(row))
        self.ring.copy_(torch.from_numpy(np.tile(row, self.update_buffer_size)))

    @staticmethod
    def create(cfg, metric_names, start_update_idx, device):
        return TrainingMetrics(metric_names, cfg.metrics_buffer_size, start_update_idx, device)

    def slot(self, name, count=1):
        """uint8 view of the staging record(s) starting at `name` (kernels write here)."""
        i = self.index[name]
        return self.staging[i * REC:(i + count) * REC]

    def advance(self):
        n = len(self.names) * REC
        dst = self.ring[self.cur_buffer_offset * n:(self.cur_buffer_offset + 1) * n]
        call('mlb_copy_bytes', ptr(self.staging), ptr(dst), c_size_t(n))
        self.update_idx += 1
        self.cur_buffer_offset = (self.cur_buffer_offset + 1) % self.update_buffer_size
        return self

    # host side (off the hot path) ------------------------------------------------------
    def to_host(self) -> Dict[str, List[Metric]]:
        raw = self.ring.cpu().numpy().tobytes()
        n = len(self.names)
        out = {k: [] for k in self.names}
        for s in range(self.update_buffer_size):
            for i, k in enumerate(self.names):
                m = _lib.Metric.from_buffer_copy(raw[(s * n + i) * REC:(s * n + i + 1) * REC])
                out[k].append(Metric(True, m.mean, m.m2, m.min, m.max, m.count))
        return out

    def latest(self) -> Dict[str, Metric]:
        h = self.to_host()
        s = (self.cur_buffer_offset - 1) % self.update_buffer_size
        return {k: v[s] for k, v in h.items()}

    def tensorboard_log(self, base_update_idx, writer):
        """ml/metrics.py:218-244: every ring slot `buf_idx` is written at step base_update_idx +
        buf_idx with the reference's tag set; all metrics on this path are per-policy records
        (Metric.init(True), ml/ppo.py:98-104, ml/rollouts.py:482-499) and P = 1, so the tags are
        `p0/<name> Mean`, `p0/<name> σ`, `p0/<name> Min`, `p0/<name> Max`."""
        h = self.to_host()
        for buf_idx in range(self.update_buffer_size):
            out_idx = base_update_idx + buf_idx
            for name, recs in h.items():
                m = recs[buf_idx]
                stddev = float(np.sqrt(m.m2 / m.count)) if m.count else float('nan')
                pre = 'p0/' if m.per_policy else ''
                writer.scalar(f'{pre}{name} Mean', m.mean, out_idx)
                writer.scalar(f'{pre}{name} σ', stddev, out_idx)
                writer.scalar(f'{pre}{name} Min', m.min, out_idx)
                writer.scalar(f'{pre}{name} Max', m.max, out_idx)

    def pretty_print(self, tab=2):
        for k, m in self.latest().items():
            std = (m.m2 / max(m.count, 1)) ** 0.5
            print(' ' * tab + f'{k}: avg {m.mean: .3e}, min {m.min: .3e}, max {m.max: .3e}, '
                              f'std {std: .3e}')
