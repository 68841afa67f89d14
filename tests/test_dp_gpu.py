"""Multi-GPU parity of the fused NVLink gradient all-reduce (mlb_allreduce_sumsq_f32): skipped on
single-GPU boxes; with >= 2 GPUs it runs tools/dp_allreduce_check.py under torchrun (bit-exact
against the rank-ordered fp32 sum, fp32-rounding-close to NCCL, CUDA-graph replay)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_allreduce_matches_rank_ordered_sum():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    world = 2
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
                        '--master-addr', '127.0.0.1', '--master-port', '29561',
                        os.path.join(ROOT, 'tools', 'dp_allreduce_check.py')],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0 and 'DP_CHECK_OK' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
