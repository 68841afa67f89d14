"""Multi-GPU parity, run under torchrun for every world size in {2, 4, 8} the box offers (skipped on
single-GPU boxes; world > 2 selects the NVLS `multimem` all-reduce variant):
  * tools/dp_allreduce_check.py -- the fused NVLink gradient all-reduce + global-norm kernel: bit-exact
    against the rank-ordered fp32 sum (peer-load variant) / identical on all ranks and fp32-close
    (NVLS), equal to NCCL to fp32 rounding, CUDA-graph replay;
  * tools/dp_update_check.py -- index-exact global minibatch permutation (the ranks' owner-affine id lists partition every global minibatch; R-rank rows ==
    1-GPU minibatches bit for bit) and the R-rank update_iter == the 1-GPU update on the
    concatenated rollout (parameters rel-L2 <= 1e-5, fp32)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(world, script, port, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
                           '--master-addr', '127.0.0.1', '--master-port', str(port),
                           os.path.join(ROOT, 'tools', script)],
                          capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=e)


def _need(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs >= {world} GPUs')


@pytest.mark.parametrize('world', [2, 4, 8])
def test_fused_allreduce_matches_rank_ordered_sum(world):
    _need(world)
    r = _torchrun(world, 'dp_allreduce_check.py', 29561 + world)
    assert r.returncode == 0 and 'DP_CHECK_OK' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize('world', [2, 4, 8])
def test_data_parallel_update_equals_single_gpu(world):
    _need(world)
    r = _torchrun(world, 'dp_update_check.py', 29581 + world)
    assert r.returncode == 0 and 'DP_UPDATE_OK' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_data_parallel_update_equals_single_gpu_bf16():
    _need(2)
    r = _torchrun(2, 'dp_update_check.py', 29591, env={'MLB_DP_DTYPE': 'bf16'})
    assert r.returncode == 0 and 'DP_UPDATE_OK' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
