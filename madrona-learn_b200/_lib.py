"""ctypes binding of libmlb200.so (the C-ABI in include/mlb200.h).

This is the host-side stub a maintainer of the reference would add where the reference
today relies on XLA to compile its jnp expressions (INTEGRATION.md shows the jax.ffi
equivalent).  PyTorch is used only for device memory and streams: every call passes raw
device pointers (``Tensor.data_ptr()``) and ``torch.cuda.current_stream().cuda_stream``.

There is NO CPU fallback: if the shared library is missing or a call fails, an exception is
raised (``MLBError``).
"""
import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(_HERE, 'csrc')
LIB_PATH = os.path.join(_HERE, 'lib', 'libmlb200.so')
ABI_VERSION = 11

c_void_p, c_int, c_ll, c_float, c_size_t = (ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong,
                                            ctypes.c_float, ctypes.c_size_t)


class MLBError(RuntimeError):
    pass


class Metric(ctypes.Structure):
    """mlb_metric (ml/metrics.py:12-18)."""
    _fields_ = [('mean', c_float), ('m2', c_float), ('min', c_float), ('max', c_float),
                ('count', ctypes.c_int32)]


P = c_void_p
# name -> (restype, argtypes); mirrors include/mlb200.h one to one
SIGNATURES = {
    'mlb_abi_version': (c_int, []),
    'mlb_gae_workspace': (c_size_t, [c_int, c_ll]),
    'mlb_gae_f32': (c_int, [P, P, P, P, P, P, P, c_int, c_ll, c_float, c_float, P, P, P, c_size_t]),
    'mlb_returns_f32': (c_int, [P, P, P, P, P, c_int, c_ll, c_float]),
    'mlb_moments_workspace': (c_size_t, [c_ll]),
    'mlb_moments_f32': (c_int, [P, P, c_ll, c_float, P, P, c_size_t]),
    'mlb_zscore_apply_f32': (c_int, [P, P, P, c_ll, P]),
    'mlb_zscore_f32': (c_int, [P, P, P, c_ll, P, P, c_size_t]),
    'mlb_metric_f32': (c_int, [P, P, c_ll, P, P, c_size_t]),
    'mlb_traj_moments_f32': (c_int, [P, P, c_int, c_ll, c_int, P]),
    'mlb_mb_moments_f32': (c_int, [P, P, P, c_int, c_ll, c_ll, c_int, c_float, P, P]),
    'mlb_moments_finalize_f32': (c_int, [P, P, c_int, ctypes.c_double, c_float, P]),
    'mlb_ema_update_f32': (c_int, [P, P, c_int, P, P, c_float, c_float]),
    'mlb_ema_scan_f32': (c_int, [P, P, P, c_int, c_float, c_float, P]),
    'mlb_ema_normalize_f32': (c_int, [P, P, c_int, P, P, c_ll]),
    'mlb_ema_invert_f32': (c_int, [P, P, c_int, P, P, c_ll]),
    'mlb_env_returns_f32': (c_int, [P, P, P, P, P, c_ll, c_float]),
    'mlb_post_step_store_f32': (c_int, [P, P, P, P, P, P, P, c_ll, c_float]),
    'mlb_threefry_split': (c_int, [P, P, P, c_int, c_int]),
    'mlb_threefry_bits': (c_int, [P, P, P, c_ll, c_int]),
    'mlb_ppo_permutations_workspace': (c_size_t, [c_int, c_ll]),
    'mlb_ppo_permutations': (c_int, [P, P, P, c_int, c_ll, c_int, P, c_size_t]),
    'mlb_ppo_permutations_of': (c_int, [P, P, P, P, c_int, c_ll, c_int, P, c_size_t]),
    'mlb_sort_pad': (c_ll, [c_ll]),
    'mlb_sort_u64': (c_int, [P, P, c_ll]),
    'mlb_obs_moments_f32': (c_int, [P, P, c_ll, c_int, P]),
    'mlb_obs_stats_merge_f32': (c_int, [P, P, c_int, ctypes.c_double, c_int, P, P]),
    'mlb_ema_estimate_update_f32': (c_int, [P, P, P, c_float]),
    'mlb_filter_adv_keys': (c_int, [P, P, c_int, c_int, c_ll, c_ll, P, P]),
    'mlb_filter_adv_select': (c_int, [P, P, c_ll, c_ll, P, P, P]),
    'mlb_partition_valid': (c_int, [P, P, P, c_int, c_ll]),
    'mlb_flat_time_index': (c_int, [P, P, c_ll, c_int, c_ll, P]),
    'mlb_traj_scores_f32': (c_int, [P, P, P, P, c_int, c_int, c_ll, P]),
    'mlb_softmax_weights_f32': (c_int, [P, P, c_ll, P, P]),
    'mlb_gumbel_topk_keys': (c_int, [P, P, P, c_ll, c_ll, c_int, P]),
    'mlb_take_sorted_indices': (c_int, [P, P, c_ll, P]),
    'mlb_gather_f32': (c_int, [P, P, P, c_ll, P]),
    'mlb_mb_gather': (c_int, [P, P, P, P, c_int, c_int, c_ll, c_ll, c_ll]),
    'mlb_mb_gather_rnn': (c_int, [P, P, P, P, c_int, c_ll, c_ll, c_ll]),
    'mlb_gemm_f32': (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                             c_int, c_int]),
    'mlb_gemm_tf32_tc': (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                         c_int, c_int]),
    'mlb_gemm_tf32_ok': (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    'mlb_dense_ln_relu_fwd_tf32': (c_int, [P, P, P, P, P, P, P, P, c_ll, c_int, c_int, c_int]),
    'mlb_ln_relu_fwd_f32': (c_int, [P, P, P, P, P, P, c_ll, c_int]),
    'mlb_ln_relu_bwd_f32': (c_int, [P, P, P, P, P, P, P, P, P, c_ll, c_int]),
    'mlb_ln_relu_fwd_bf16': (c_int, [P, P, P, P, P, P, c_ll, c_int]),
    'mlb_ln_relu_bwd_bf16': (c_int, [P, P, P, P, P, P, P, P, P, c_ll, c_int]),
    'mlb_gemm_bf16_tc': (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_int]),
    'mlb_dense_ln_relu_fwd_tc': (c_int, [P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int]),
    'mlb_dense_dx_lnbwd_tc': (c_int, [P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int]),
    'mlb_cast_f32_bf16': (c_int, [P, P, P, c_ll]),
    'mlb_cast_weight_bf16': (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int]),
    'mlb_lstm_cell_fwd_f32': (c_int, [P, P, P, P, P, P, P, P, P, c_ll, c_int]),
    'mlb_lstm_cell_bwd_f32': (c_int, [P, P, c_int, P, P, P, P, P, P, P, c_ll, c_int]),
    'mlb_lstm_cell_fwd_tc': (c_int, [P, P, P, P, P, P, P, P, P, P, c_ll, c_int]),
    'mlb_lstm_cell_bwd_tc': (c_int, [P, P, c_int, P, P, P, P, P, P, P, c_ll, c_int]),
    'mlb_lstm_pack_weights_bf16': (c_int, [P, P, P, P, P, P, c_int, c_int]),
    'mlb_lstm_step_tc': (c_int, [P, P, c_int, P, P, P, P, P, P, P, P, P, P, c_ll, c_int, c_int]),
    'mlb_rnn_reset_f32': (c_int, [P, P, P, c_ll, c_int]),
    'mlb_rollout_keys': (c_int, [P, P, P, c_int]),
    'mlb_sample_discrete_f32': (c_int, [P, P, c_int, P, P, c_int, c_ll, c_int, c_int, P, P, P, P, c_int]),
    'mlb_sample_continuous_f32': (c_int, [P, P, c_int, P, c_int, c_float, c_float, c_ll, c_int, c_int, P, P, P, P, c_int]),
    'mlb_mb_gather_multi': (c_int, [P, P, c_int, P, c_int, c_int, c_ll, c_ll]),
    'mlb_mb_gather_multi_peer': (c_int, [P, P, c_int, P, c_int, P, c_int, c_int, c_ll, c_ll]),
    'mlb_dp_assign_minibatches': (c_int, [P, P, c_ll, c_int, c_int, c_int, c_int, c_ll, c_ll, P]),
    'mlb_reorder_chunks_workspace': (c_size_t, [c_ll, c_int]),
    'mlb_reorder_chunks': (c_int, [P, P, c_ll, c_int, c_int, c_ll, P, P, P, c_size_t]),
    'mlb_gather_rows_clip': (c_int, [P, P, P, P, c_ll, c_ll, c_ll]),
    'mlb_allreduce_workspace': (c_size_t, []),
    'mlb_allreduce_nvls_f32': (c_int, [P, P, P, P, P, c_ll, P, P, P, c_size_t]),
    'mlb_allreduce_sumsq_f32': (c_int, [P, P, P, c_ll, P, P, P, c_size_t]),
    'mlb_policy_rollout_tc': (c_int, [P, P, P, P, c_ll, P, P, P, c_int, c_int, c_int, P, P, P, P, c_int, P]),
    'mlb_policy_rollout_ps_tc': (c_int, [P, P, P, P, c_ll, P, P, P, c_int, c_int, c_int, P, P, P, P, c_int, P, P]),
    'mlb_ppo_loss_workspace': (c_size_t, [c_ll]),
    'mlb_ppo_loss_f32': (c_int, [P, P, c_int, P, P, P, P, P, P, P, P, P, P, P, c_int, c_ll, c_ll,
                                 c_float, c_float, c_int, P, P, P, P, c_size_t, P, c_int]),
    'mlb_fill_zero': (c_int, [P, P, c_size_t]),
    'mlb_fill_zero_2d': (c_int, [P, P, c_int, c_int, c_int]),
    'mlb_copy_bytes': (c_int, [P, P, P, c_size_t]),
    'mlb_sumsq_workspace': (c_size_t, [c_ll]),
    'mlb_sumsq_f32': (c_int, [P, P, c_ll, P, P, c_size_t]),
    'mlb_adam_step_f32': (c_int, [P, P, P, P, P, c_ll, P, P, c_float, c_float, c_float, c_float,
                                  c_float, c_float]),
    'mlb_renorm_segments': (c_int, [P, P, P, c_int, P, P]),
    'mlb_optimizer_fused_workspace': (c_size_t, []),
    'mlb_optimizer_step_fused': (c_int, [P, P, P, P, P, c_ll, P, c_int, P, P, P, c_int, c_float, c_float, c_float,
                                         c_float, c_float, c_float, P, P, c_size_t, P]),
    'mlb_colsum_f32': (c_int, [P, P, c_ll, c_int, c_int, P]),
    'mlb_colsum_bf16': (c_int, [P, P, c_ll, c_int, c_int, P]),
    'mlb_synth_env_init': (c_int, [P, P, c_ll, c_int, ctypes.c_uint32, P]),
    'mlb_synth_env_step': (c_int, [P, P, P, P, c_int, P, P, P, c_ll, c_int, ctypes.c_uint32,
                                   c_float]),
}


class Segment(ctypes.Structure):
    """mlb_segment."""
    _fields_ = [('offset', c_ll), ('length', c_ll), ('kind', ctypes.c_int32), ('target', c_float)]


class Bf16Copy(ctypes.Structure):
    """mlb_bf16_copy."""
    _fields_ = [('dst_t', c_void_p), ('dst', c_void_p), ('rows', ctypes.c_int32), ('cols', ctypes.c_int32),
                ('ld_t', ctypes.c_int32), ('ld_d', ctypes.c_int32)]


class GatherLeaf(ctypes.Structure):
    """mlb_gather_leaf."""
    _fields_ = [('store', c_void_p), ('out', c_void_p), ('out_bf16', c_void_p), ('row_bytes', ctypes.c_longlong)]


class PeerTable(ctypes.Structure):
    """mlb_peer_table."""
    _fields_ = [('rank', ctypes.c_int32), ('world', ctypes.c_int32), ('grads', c_void_p * 16),
                ('signals', c_void_p * 16)]


class MlpTcDesc(ctypes.Structure):
    """mlb_mlp_tc_desc."""
    _fields_ = [('num_layers', ctypes.c_int32), ('obs_dim', ctypes.c_int32), ('hidden', ctypes.c_int32),
                ('head_width', ctypes.c_int32), ('w_t', c_void_p * 4), ('scale', c_void_p * 4),
                ('bias', c_void_p * 4), ('wh_t', c_void_p), ('head_bias', c_void_p)]


class PostStep(ctypes.Structure):
    """mlb_post_step."""
    _fields_ = [('rewards', c_void_p), ('dones', c_void_p), ('reward_slab', c_void_p), ('done_slab', c_void_p),
                ('env_returns', c_void_p), ('trace', c_void_p), ('gamma', c_float)]


class PPOStats(ctypes.Structure):
    """mlb_ppo_stats."""
    _fields_ = [('loss', c_float), ('action_obj', c_float), ('value_loss', c_float),
                ('entropy', c_float), ('metrics', Metric * 5)]

_lib = None


def build(verbose=False):
    """Compile libmlb200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ['make', '-C', CSRC_DIR, '-j', str(os.cpu_count() or 4)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise MLBError('libmlb200 build failed')
    return LIB_PATH


def lib():
    """Load (once) and return the ctypes handle; raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MLBError(f'{LIB_PATH} not found: run `python -c "import __graft_entry__ as g; '
                       f'g.build()"` (there is no CPU fallback)')
    h = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(h, name)
        except AttributeError as e:
            raise MLBError(f'libmlb200.so does not export {name}; rebuild') from e
        fn.restype = res
        fn.argtypes = args
    v = h.mlb_abi_version()
    if v != ABI_VERSION:
        raise MLBError(f'libmlb200 ABI {v} != binding {ABI_VERSION}; rebuild')
    _lib = h
    return h


def check(rc, name):
    if rc != 0:
        kind = 'cudaError' if rc > 0 else 'MLB_E'
        raise MLBError(f'{name} failed: {kind} {rc}')


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise MLBError('libmlb200 takes device pointers only (tensor is on CPU)')
    if not t.is_contiguous():
        raise MLBError('libmlb200 requires contiguous buffers')
    return c_void_p(t.data_ptr())


CALLS = 0     # number of C-ABI enqueue calls made by this process (bench.py's gpu_launches)


def call(name, *args):
    """Invoke an int-returning entry point on the current torch stream and check its code."""
    global CALLS
    CALLS += 1
    fn = getattr(lib(), name)
    rc = fn(stream_ptr(), *args)
    check(rc, name)
