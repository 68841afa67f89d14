"""PolicyProgram: lowers an ActorCritic descriptor onto libmlb200 kernels.

This is the B200-side replacement for "flax module + XLA": parameters live in ONE flat fp32
arena (with gradient / Adam-moment arenas of the same layout) so the optimiser, the gradient
all-reduce and the norm re-projection are single launches; activations live in preallocated
workspaces so the whole update can be captured in a CUDA graph.

Supported family (SURVEY 8a rows a15-a19 and the 8f rows lowered in round 2): BackboneShared or BackboneSeparate
(prefix=None|identity; encoders BackboneEncoder(net=MLP) or, shared only, RecurrentBackboneEncoder(net=MLP,
rnn=LSTM)) + DenseLayerDiscreteActor | DenseLayerContinuousActor + DenseLayerCritic | DreamerV3Critic |
HLGaussCritic; compute_dtype float32 (Dense products on tcgen05 kind::tf32 by default, exact FFMA with
set_matmul_precision('highest')) or bfloat16 (fused tcgen05 layer kernels).  Everything else raises
NotImplementedError loudly (no fallback).
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import c_float, c_int, c_ll, c_size_t, call, ptr
from .actor_critic import (ActorCritic, BackboneEncoder, BackboneSeparate, BackboneShared,
                           RecurrentBackboneEncoder)
from .cfg import ContinuousActionsConfig, DiscreteActionsConfig
from .models import (MLP, DenseLayerContinuousActor, DenseLayerCritic, DenseLayerDiscreteActor, DreamerV3Critic,
                     HLGaussCritic)

F32 = torch.float32


def _round_up(x, m):
    return (x + m - 1) // m * m


# compute_dtype=float32 products: 'tf32' (the default) = tcgen05.mma.kind::tf32, what XLA:GPU runs for an f32 dot
# at its default precision, i.e. the reference's default configuration (ml/cfg.py:96); 'highest' = exact fp32 FFMA
# (SIMT).  MLB_MATMUL_PRECISION / set_matmul_precision() mirror jax_default_matmul_precision; the exact-parity
# test suites pin 'highest' (tests/conftest.py).
_PRECISION = {'tf32': os.environ.get('MLB_MATMUL_PRECISION', 'tf32').lower() in ('tf32', 'default')}


def set_matmul_precision(precision):
    """'tf32' | 'default' (tensor cores) or 'highest' | 'float32' (exact fp32) for compute_dtype=float32."""
    p = str(precision).lower()
    if p not in ('tf32', 'default', 'highest', 'float32'):
        raise ValueError("matmul precision must be 'tf32'/'default' or 'highest'/'float32'")
    _PRECISION['tf32'] = p in ('tf32', 'default')


def matmul_precision():
    return 'tf32' if _PRECISION['tf32'] else 'highest'


def _tma_ok(t, ld):
    return ld % 4 == 0 and t.data_ptr() % 16 == 0


def gemm(A, B, C, bias, M, N, K, lda, ldb, ldc, ta=0, tb=0, accumulate=0, splitk=1):
    if (_PRECISION['tf32'] and N % 4 == 0 and _tma_ok(A, lda) and _tma_ok(B, ldb) and _tma_ok(C, ldc) and
            (bias is None or bias.data_ptr() % 16 == 0)):
        call('mlb_gemm_tf32_tc', ptr(A), ptr(B), ptr(C), ptr(bias), c_int(M), c_int(N), c_int(K),
             c_int(lda), c_int(ldb), c_int(ldc), c_int(ta), c_int(tb), c_int(accumulate), c_int(splitk))
        return
    call('mlb_gemm_f32', ptr(A), ptr(B), ptr(C), ptr(bias), c_int(M), c_int(N), c_int(K),
         c_int(lda), c_int(ldb), c_int(ldc), c_int(ta), c_int(tb), c_int(accumulate),
         c_int(splitk))


def dense_ln_relu_f32(x, d, k, s, b, z, y, stats, rows, H, need_z=True):
    """One fp32 MLP layer (ml/models.py:107-117): z = x k, y = relu(LayerNorm(z)), stats = {mean, rstd}.  With
    matmul precision 'tf32' and H in {64, 128, 256}: ONE launch (mlb_dense_ln_relu_fwd_tf32, LayerNorm in the
    tensor-core GEMM's epilogue; z is written only if need_z); otherwise GEMM + LayerNorm kernels (z is scratch)."""
    if (_PRECISION['tf32'] and H in (64, 128, 256) and _tma_ok(x, d) and _tma_ok(k, H) and _tma_ok(y, H) and
            (z is None or _tma_ok(z, H)) and s.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0):
        call('mlb_dense_ln_relu_fwd_tf32', ptr(x), ptr(k), ptr(s), ptr(b), ptr(z if need_z else None), ptr(y),
             ptr(stats), c_ll(rows),
             c_int(d), c_int(H), c_int(d))
        return
    gemm(x, k, z, None, rows, H, d, d, H, H)
    call('mlb_ln_relu_fwd_f32', ptr(z), ptr(s), ptr(b), ptr(y), ptr(stats), c_ll(rows), c_int(H))


def gemm_tc(A, B, C, bias, M, N, K, lda, ldb, ldc, a_mn=0, b_mn=0, epi=0, splitk=1):
    """mlb_gemm_bf16_tc: C[M,N] (+)= A * B^T on tcgen05 (bf16 operands, fp32 accumulate)."""
    call('mlb_gemm_bf16_tc', ptr(A), ptr(B), ptr(C), ptr(bias), c_int(M), c_int(N), c_int(K), c_int(lda),
         c_int(ldb), c_int(ldc), c_int(a_mn), c_int(b_mn), c_int(epi), c_int(splitk))


def _splitk_tc(M, N, K):
    bn = 64 if N <= 64 else 128          # the split-K (atomic) epilogue uses N tiles <= 128
    tiles = math.ceil(M / 128) * math.ceil(N / bn)
    want = max(1, 148 // tiles)
    return int(max(1, min(want, K // 256)))


def _splitk_for(M, N, K):
    bm, bn = (128, 32) if N <= 32 else ((64, 64) if (N <= 64 or M <= 64) else (128, 128))
    tiles = math.ceil(M / bm) * math.ceil(N / bn)
    want = max(1, math.ceil(2 * 148 / tiles))
    return int(max(1, min(want, K // 512 if K >= 1024 else 1)))


class PolicyProgram:
    def __init__(self, actor_critic, obs_dim, actions_cfg, device, compute_dtype=F32):
        if not isinstance(actor_critic, ActorCritic):
            raise TypeError('policy.actor_critic must be an ActorCritic descriptor')
        if compute_dtype not in (F32, torch.bfloat16):
            raise NotImplementedError('compute_dtype must be float32 (tf32 tensor cores | exact FFMA) or bfloat16 (tcgen05)')
        self.tc = compute_dtype == torch.bfloat16
        bb = actor_critic.backbone
        if not isinstance(bb, (BackboneShared, BackboneSeparate)):
            raise NotImplementedError('backbone must be BackboneShared or BackboneSeparate')
        if bb.prefix is not None and not getattr(bb.prefix, 'is_identity', False):
            raise NotImplementedError('Backbone.prefix must be None (obs are one [N, D] tensor)')
        # BackboneSeparate (ml/actor_critic.py:247-303): two encoder towers on the same processed observations,
        # the actor head on the actor tower's features, the critic head on the critic tower's.  Lowered as NT = 2
        # MLP stacks in the one arena (flat layer index t * L + i) feeding ONE head GEMM over the concatenated
        # features [feat_actor | feat_critic] whose weight is block-diagonal: the off-diagonal blocks are zero
        # at init and their gradient is cleared before the optimiser sees it (mlb_fill_zero_2d), so they stay
        # exactly zero and the product equals the two separate Dense heads.
        self.NT = 2 if isinstance(bb, BackboneSeparate) else 1
        if self.NT == 2:
            enc, enc_c = bb.actor_encoder, bb.critic_encoder
            if not (isinstance(enc, BackboneEncoder) and isinstance(enc_c, BackboneEncoder) and
                    isinstance(enc.net, MLP) and isinstance(enc_c.net, MLP)):
                raise NotImplementedError('BackboneSeparate encoders must be BackboneEncoder(net=MLP) '
                                          '(recurrent separate towers are not lowered)')
            if (enc.net.num_channels, enc.net.num_layers) != (enc_c.net.num_channels, enc_c.net.num_layers):
                raise NotImplementedError('BackboneSeparate: actor and critic MLPs must have the same shape')
        else:
            enc = bb.encoder
        if not isinstance(enc, (BackboneEncoder, RecurrentBackboneEncoder)) or not isinstance(enc.net, MLP):
            raise NotImplementedError('encoder must be [Recurrent]BackboneEncoder(net=MLP[, rnn=LSTM])')
        self._rnn_desc = enc.rnn if isinstance(enc, RecurrentBackboneEncoder) else None
        if not isinstance(actor_critic.actor, (DenseLayerDiscreteActor, DenseLayerContinuousActor)):
            raise NotImplementedError('actor must be DenseLayerDiscreteActor or DenseLayerContinuousActor')
        if not isinstance(actor_critic.critic, (DenseLayerCritic, DreamerV3Critic, HLGaussCritic)):
            raise NotImplementedError('critic must be DenseLayerCritic, DreamerV3Critic or HLGaussCritic')
        self.ac = actor_critic
        self.device = torch.device(device)
        self.mlp = enc.net
        self.obs_dim = int(obs_dim)
        self.H = int(self.mlp.num_channels)
        self.L = int(self.mlp.num_layers)
        if self.H % 4 or self.H > 1024:
            raise NotImplementedError('MLP width must be a multiple of 4 and <= 1024')
        if self.tc and (self.H % 8 or self.obs_dim % 8):
            raise NotImplementedError('tensor-core path needs obs_dim and width multiples of 8 (TMA 16 B rows)')
        # action layout: groups in cfg.actions order, components concatenated
        self.groups = []
        buckets = []
        self.continuous = None
        for name, ac in actions_cfg.items():
            if isinstance(ac, ContinuousActionsConfig):
                # ContinuousActionDistributions (ml/dists.py:211-284): num_dims components, each a Normal with
                # mean = tanh(raw), std = (max - min) * sigmoid(raw + 2) + min; head columns means | stds
                self.continuous = (float(ac.stddev_min), float(ac.stddev_max), int(ac.num_dims))
                self.groups.append((name, 0, int(ac.num_dims)))
                continue
            if not isinstance(ac, DiscreteActionsConfig):
                raise NotImplementedError('actions must be DiscreteActionsConfig or ContinuousActionsConfig')
            self.groups.append((name, len(buckets), len(ac.actions_num_buckets)))
            buckets += list(ac.actions_num_buckets)
        if len(self.groups) != 1:
            raise NotImplementedError('the actor head drives exactly one action group')
        if (self.continuous is not None) != isinstance(actor_critic.actor, DenseLayerContinuousActor):
            raise ValueError('ContinuousActionsConfig <-> DenseLayerContinuousActor, DiscreteActionsConfig <-> '
                             'DenseLayerDiscreteActor')
        self.buckets = buckets
        if self.continuous is not None:
            self.A = self.continuous[2]
            self.sumA = 2 * self.A
        else:
            self.A = len(buckets)
            self.sumA = int(sum(buckets))
        # critic columns: 1 (plain) or num_bins two-hot logits (DreamerV3Critic, ml/models.py:157-174)
        self.twohot = isinstance(actor_critic.critic, DreamerV3Critic)
        self.hlgauss = isinstance(actor_critic.critic, HLGaussCritic)
        self.V = int(actor_critic.critic.num_bins) if (self.twohot or self.hlgauss) else 1
        if self.hlgauss:
            # HLGaussCritic (ml/models.py:253-306): centres | bounds | smoothness, the table mlb_ppo_loss_f32 reads
            # under MLB_PPO_HLGAUSS_CRITIC; the value decode (rollout) only uses the V centres in front
            cr = actor_critic.critic
            if cr.centers is None or self.V % 2 == 0 or self.V > 127:
                raise NotImplementedError('HLGaussCritic: build with HLGaussCritic.create, odd num_bins <= 127')
            tab = np.concatenate([cr.centers, cr.bounds, [cr.smoothness]]).astype(np.float32)
            self._bins_c = (ctypes.c_float * tab.size)(*tab.tolist())
        elif self.twohot:
            # SymExpTwoHotDistribution._compute_bins (ml/dists.py:127-141), float32 like the reference
            half = np.linspace(-14, 0, self.V // 2 + 1, dtype=np.float32)
            half = (np.sign(half) * np.expm1(np.abs(half))).astype(np.float32)
            bins = np.concatenate([half, -half[:-1][::-1]]).astype(np.float32)
            self._bins_c = (ctypes.c_float * self.V)(*bins.tolist())
        else:
            self._bins_c = None
        # head width: padded to 4 (fp32 path) or to 64 (tensor-core path: one 64-wide UMMA N tile
        # and one SWIZZLE_128B atom of the MN-major dhead operand)
        self.NH = _round_up(self.sumA + self.V, 64 if self.tc else 4)
        self._buckets_c = (ctypes.c_int32 * self.A)(*buckets) if self.continuous is None else None
        # ---- arena layout -------------------------------------------------------------
        off = 0
        self.layer_off = []                   # flat layer index f = tower * L + i
        self.LT = self.NT * self.L
        for _t in range(self.NT):
            d = self.obs_dim
            for _ in range(self.L):
                k_off = off
                off += d * self.H
                ln_off = off                  # scale[H] | bias[H] contiguous (one LN segment)
                off += 2 * self.H
                self.layer_off.append((k_off, ln_off, d))
                d = self.H
        self.lstm = None
        self.feat = self.H * self.NT          # BackboneSeparate: [actor features | critic features]
        if self._rnn_desc is not None:
            from .recurrent import LSTMLowering
            self.lstm = LSTMLowering(self, self._rnn_desc, self.H, off)
            off = self.lstm.end_off
            self.feat = self.lstm.RH * self.lstm.RL
        self.head_w_off = off
        off += self.feat * self.NH
        self.head_b_off = off
        off += self.NH
        self.num_params = off
        dev = self.device
        self.params = torch.zeros(off, dtype=F32, device=dev)
        self.grads = torch.zeros(off, dtype=F32, device=dev)
        self.adam_m = torch.zeros(off, dtype=F32, device=dev)
        self.adam_v = torch.zeros(off, dtype=F32, device=dev)
        self.adam_step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.grad_sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self._sumsq_ws = torch.empty(_lib.lib().mlb_sumsq_workspace(off) + 16, dtype=torch.uint8, device=dev)
        self.segments = None
        self._train_ws = None
        self._infer_ws = None
        if self.tc:
            BF = torch.bfloat16
            # bf16 operand copies of the weights: W^T [H, in] (forward, K-major B) and W [in, H]
            # (dX, K-major B); refreshed by refresh_bf16() after every optimiser step
            self.w_t = [torch.zeros(self.H, d_, dtype=BF, device=dev) for (_, _, d_) in self.layer_off]
            self.w_c = [torch.zeros(d_, self.H, dtype=BF, device=dev) for (_, _, d_) in self.layer_off]
            self.wh_t = torch.zeros(self.NH, self.feat, dtype=BF, device=dev)
            self.wh_c = torch.zeros(self.feat, self.NH, dtype=BF, device=dev)

    # ---------------------------------------------------------------------------------
    # parameters
    # ---------------------------------------------------------------------------------
    def _view(self, arena, off, *shape):
        n = int(np.prod(shape))
        return arena[off:off + n].view(*shape)

    def layer_views(self, arena, i):
        k_off, ln_off, d = self.layer_off[i]
        return (self._view(arena, k_off, d, self.H), self._view(arena, ln_off, self.H),
                self._view(arena, ln_off + self.H, self.H))

    def head_views(self, arena):
        return (self._view(arena, self.head_w_off, self.feat, self.NH),
                self._view(arena, self.head_b_off, self.NH))

    def _lname(self, f):
        """Key of flat layer f in initial_weight_norms (tower 1 = the critic encoder of BackboneSeparate)."""
        return f'Dense_{f}' if f < self.L else f'critic/Dense_{f - self.L}'

    def _net_tree(self, a, t):
        net = {}
        for i in range(self.L):
            k, s, b = self.layer_views(a, t * self.L + i)
            net[f'Dense_{i}'] = {'kernel': k}
            net[f'LayerNorm_{i}'] = {'impl': {'scale': s, 'bias': b}}
        return net

    def param_tree(self, arena=None):
        """flax-style nested dict of VIEWS into the arena (ml/train_state.py:34-40 `params`)."""
        a = self.params if arena is None else arena
        W, B = self.head_views(a)
        if self.NT == 2:                      # ml/actor_critic.py:247-250: actor_encoder / critic_encoder
            H = self.H
            return {
                'backbone': {'actor_encoder': {'net': self._net_tree(a, 0)},
                             'critic_encoder': {'net': self._net_tree(a, 1)}},
                'actor': {'impl': {'kernel': W[:H, :self.sumA], 'bias': B[:self.sumA]}},
                'critic': {'Dense_0': {'kernel': W[H:, self.sumA:self.sumA + self.V],
                                       'bias': B[self.sumA:self.sumA + self.V]}},
            }
        enc = {'net': self._net_tree(a, 0)}
        if self.lstm is not None:
            enc['rnn'] = self.lstm.param_tree(a)
        return {
            'backbone': {'encoder': enc},
            'actor': {'impl': {'kernel': W[:, :self.sumA], 'bias': B[:self.sumA]}},
            'critic': {'Dense_0': {'kernel': W[:, self.sumA:self.sumA + self.V],
                                   'bias': B[self.sumA:self.sumA + self.V]}},
        }

    def init_params(self, seed):
        """orthogonal(sqrt2) Dense kernels, LayerNorm (1, 0), orthogonal(0.01) actor,
        orthogonal(1.0) critic, zero biases (ml/models.py:103,125,144).  Done once on the
        host; init parity with jax's QR is not required (tests load identical weights)."""
        g = torch.Generator().manual_seed(int(seed) & 0x7FFFFFFF)

        def orth(rows, cols, scale):
            a = torch.randn(max(rows, cols), min(rows, cols), generator=g, dtype=torch.float64)
            q, r = torch.linalg.qr(a)
            q = q * torch.sign(torch.diagonal(r))
            if rows < cols:
                q = q.t()
            return (scale * q[:rows, :cols]).to(F32)
        host = torch.zeros(self.num_params, dtype=F32)
        for i in range(self.LT):
            k_off, ln_off, d = self.layer_off[i]
            host[k_off:k_off + d * self.H] = orth(d, self.H, self.mlp.weight_init_scale).reshape(-1)
            host[ln_off:ln_off + self.H] = 1.0
        if self.lstm is not None:
            self.lstm.init_host(host, orth)
        W = torch.zeros(self.feat, self.NH, dtype=F32)
        fa = self.H if self.NT == 2 else self.feat          # BackboneSeparate: block-diagonal head
        W[:fa, :self.sumA] = orth(fa, self.sumA, self.ac.actor.weight_init_scale)
        if not (self.twohot or self.hlgauss):      # DreamerV3Critic / HLGaussCritic are zero-initialised (ml/models.py:159,261)
            W[self.feat - fa:, self.sumA:self.sumA + 1] = orth(fa, 1, self.ac.critic.weight_init_scale)
        host[self.head_w_off:self.head_w_off + W.numel()] = W.reshape(-1)
        self.params.copy_(host)
        self.finalize_params()

    def load_oracle_params(self, p):
        """Load a parameter tree in oracle/nn.py format (tests)."""
        host = torch.zeros(self.num_params, dtype=F32)
        for f in range(self.LT):
            k_off, ln_off, d = self.layer_off[f]
            lyr = p['mlp'][f] if f < self.L else p['mlp_critic'][f - self.L]
            host[k_off:k_off + d * self.H] = torch.from_numpy(np.asarray(lyr['kernel'], np.float32)).reshape(-1)
            host[ln_off:ln_off + self.H] = torch.from_numpy(np.asarray(lyr['scale'], np.float32))
            host[ln_off + self.H:ln_off + 2 * self.H] = torch.from_numpy(np.asarray(lyr['bias'], np.float32))
        if self.lstm is not None:
            self.lstm.load_oracle(host, p['lstm'])
        W = torch.zeros(self.feat, self.NH, dtype=F32)
        fa = self.H if self.NT == 2 else self.feat
        W[:fa, :self.sumA] = torch.from_numpy(np.asarray(p['actor']['kernel'], np.float32))
        W[self.feat - fa:, self.sumA:self.sumA + self.V] = torch.from_numpy(np.asarray(p['critic']['kernel'], np.float32))
        B = torch.zeros(self.NH, dtype=F32)
        B[:self.sumA] = torch.from_numpy(np.asarray(p['actor']['bias'], np.float32))
        B[self.sumA:self.sumA + self.V] = torch.from_numpy(np.asarray(p['critic']['bias'], np.float32))
        host[self.head_w_off:self.head_w_off + W.numel()] = W.reshape(-1)
        host[self.head_b_off:self.head_b_off + self.NH] = B
        self.params.copy_(host)
        self.finalize_params()

    def to_oracle_params(self, arena=None):
        t = self.param_tree(arena)
        c = lambda x: x.detach().cpu().numpy().copy()
        layers = lambda net: [{'kernel': c(net[f'Dense_{i}']['kernel']),
                               'scale': c(net[f'LayerNorm_{i}']['impl']['scale']),
                               'bias': c(net[f'LayerNorm_{i}']['impl']['bias'])} for i in range(self.L)]
        extra = {'lstm': self.lstm.to_oracle(self.params if arena is None else arena)} if self.lstm is not None else {}
        if self.NT == 2:
            extra['mlp_critic'] = layers(t['backbone']['critic_encoder']['net'])
            net = t['backbone']['actor_encoder']['net']
        else:
            net = t['backbone']['encoder']['net']
        return {**extra, 'mlp': layers(net),
                'actor': {'kernel': c(t['actor']['impl']['kernel']), 'bias': c(t['actor']['impl']['bias'])},
                'critic': {'kernel': c(t['critic']['Dense_0']['kernel']), 'bias': c(t['critic']['Dense_0']['bias'])}}

    def initial_weight_norms_tree(self):
        """initial_weight_norms in the parameter tree's shape (ml/train_state.py:413-423): the initial L2
        norm at every backbone `kernel` leaf, None at every other leaf and under actor / critic."""
        n = self.initial_weight_norms

        def net_of(t):
            net = {}
            for i in range(self.L):
                net[f'Dense_{i}'] = {'kernel': float(n[self._lname(t * self.L + i)])}
                net[f'LayerNorm_{i}'] = {'impl': {'scale': None, 'bias': None}}
            return net
        if self.NT == 2:
            return {'backbone': {'actor_encoder': {'net': net_of(0)}, 'critic_encoder': {'net': net_of(1)}},
                    'actor': {'impl': {'kernel': None, 'bias': None}},
                    'critic': {'Dense_0': {'kernel': None, 'bias': None}}}
        enc = {'net': net_of(0)}
        if self.lstm is not None:
            cells = {}
            for li in range(self.lstm.RL):
                cell = {}
                for g in ('i', 'f', 'g', 'o'):
                    cell['i' + g] = {'kernel': n.get(f'lstm{li}/i{g}')}
                    cell['h' + g] = {'kernel': n.get(f'lstm{li}/h{g}'), 'bias': None}
                cells[f'OptimizedLSTMCell_{li}'] = cell
            enc['rnn'] = {'cell': cells}
        return {'backbone': {'encoder': enc}, 'actor': {'impl': {'kernel': None, 'bias': None}},
                'critic': {'Dense_0': {'kernel': None, 'bias': None}}}

    def finalize_params(self):
        """Record initial kernel norms (ml/train_state.py:413-423) -> device segment table."""
        self.initial_weight_norms = {}
        host = self.params.detach().cpu()
        for i in range(self.LT):
            k_off, ln_off, d = self.layer_off[i]
            n0 = float(torch.linalg.vector_norm(host[k_off:k_off + d * self.H].double()))
            self.initial_weight_norms[self._lname(i)] = n0
        self.rebuild_segments()

    def refresh_bf16(self):
        """Re-derive the bf16 operand copies from the fp32 master weights (tensor-core path)."""
        if not self.tc:
            return
        for i in range(self.LT):
            k, _, _ = self.layer_views(self.params, i)
            d = self.layer_off[i][2]
            call('mlb_cast_weight_bf16', ptr(k), ptr(self.w_t[i]), ptr(self.w_c[i]), c_int(d), c_int(self.H),
                 c_int(self.H), c_int(d), c_int(self.H))
        W, _ = self.head_views(self.params)
        call('mlb_cast_weight_bf16', ptr(W), ptr(self.wh_t), ptr(self.wh_c), c_int(self.feat), c_int(self.NH),
             c_int(self.NH), c_int(self.feat), c_int(self.NH))
        if self.lstm is not None:
            self.lstm.refresh_bf16()

    def rebuild_segments(self):
        self.refresh_bf16()
        segs, copies = [], []
        none = _lib.Bf16Copy(None, None, 0, 0, 0, 0)
        for i in range(self.LT):
            k_off, ln_off, d = self.layer_off[i]
            segs.append(_lib.Segment(k_off, d * self.H, 1, float(self.initial_weight_norms[self._lname(i)])))
            copies.append(_lib.Bf16Copy(self.w_t[i].data_ptr(), self.w_c[i].data_ptr(), d, self.H, d, self.H)
                          if self.tc else none)
            segs.append(_lib.Segment(ln_off, 2 * self.H, 2, float(self.H)))
            copies.append(none)
        if self.lstm is not None:
            lsegs = self.lstm.segments(self.params.detach().cpu(), self.initial_weight_norms)
            segs += lsegs
            copies += self.lstm.bf16_copies() if self.tc else [none] * len(lsegs)
        # the fused actor+critic head matrix is not re-projected (kind 0) but its bf16 copies are refreshed
        segs.append(_lib.Segment(self.head_w_off, self.feat * self.NH, 0, 0.0))
        copies.append(_lib.Bf16Copy(self.wh_t.data_ptr(), self.wh_c.data_ptr(), self.feat, self.NH, self.feat,
                                    self.NH) if self.tc else none)
        # tile the arena exactly, in offset order (plain tensors -- head bias, LSTM bias -- become
        # kind-0 segments): what the single-launch optimiser (mlb_optimizer_step_fused) walks
        order = sorted(range(len(segs)), key=lambda i: segs[i].offset)
        tiled, tcopies, pos = [], [], 0
        for i in order:
            if segs[i].offset > pos:
                tiled.append(_lib.Segment(pos, segs[i].offset - pos, 0, 0.0)); tcopies.append(none)
            tiled.append(segs[i]); tcopies.append(copies[i])
            pos = segs[i].offset + segs[i].length
        if pos < self.num_params:
            tiled.append(_lib.Segment(pos, self.num_params - pos, 0, 0.0)); tcopies.append(none)
        segs, copies = tiled, tcopies
        self._fused_opt = (len(segs) <= 32 and os.environ.get('MLB_FUSED_OPT', '1') != '0')
        self._opt_zero = os.environ.get('MLB_OPT_ZERO', '1') != '0'
        if getattr(self, '_opt_sync', None) is None:
            self._opt_sync = torch.zeros(2, dtype=torch.int32, device=self.device)
            self._opt_ws = torch.zeros(_lib.lib().mlb_optimizer_fused_workspace(), dtype=torch.uint8,
                                       device=self.device)
        carr = (_lib.Bf16Copy * len(copies))(*copies)
        self.copies = torch.from_numpy(np.frombuffer(bytes(carr), dtype=np.uint8).copy()).to(self.device) \
            if self.tc else None
        arr = (_lib.Segment * len(segs))(*segs)
        raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
        self.segments = torch.from_numpy(raw).to(self.device)
        self.num_segments = len(segs)

    # ---------------------------------------------------------------------------------
    # workspaces
    # ---------------------------------------------------------------------------------
    def infer_ws(self, rows):
        w = self._infer_ws
        if w is None or w['rows'] < rows:
            dev = self.device
            AT = torch.bfloat16 if self.tc else F32
            w = dict(rows=rows, z=None if self.tc else torch.empty(rows, self.H, dtype=F32, device=dev),
                     y=[torch.empty(rows, self.H, dtype=AT, device=dev) for _ in range(2 * self.NT)],
                     head=torch.empty(rows, self.NH, dtype=F32, device=dev))
            if self.tc:
                w['x'] = torch.empty(rows, self.obs_dim, dtype=AT, device=dev)
            self._infer_ws = w
        return w

    def train_ws(self, rows):
        w = self._train_ws
        if w is None or w['rows'] < rows:
            dev = self.device
            AT = torch.bfloat16 if self.tc else F32
            e = lambda *s, dtype=F32: torch.empty(*s, dtype=dtype, device=dev)
            w = dict(rows=rows, z=None if self.tc else [e(rows, self.H) for _ in range(self.LT)],
                     y=[e(rows, self.H, dtype=AT) for _ in range(self.LT)],
                     stats=None if self.tc else [e(rows, 2) for _ in range(self.LT)],
                     head=e(rows, self.NH), dhead=e(rows, self.NH, dtype=AT),
                     dy=e(rows, self.H, dtype=AT), dz=e(rows, self.H, dtype=AT),
                     loss_ws=torch.zeros(_lib.lib().mlb_ppo_loss_workspace(rows) + 16, dtype=torch.uint8, device=dev),
                     stats_out=torch.zeros(ctypes.sizeof(_lib.PPOStats), dtype=torch.uint8, device=dev))
            if self.tc:
                w['x'] = e(rows, self.obs_dim, dtype=AT)
                w['xh'] = [e(rows, self.H, dtype=AT) for _ in range(self.LT)]    # normalised pre-activations
                w['rstd'] = [e(rows) for _ in range(self.LT)]
                w['dzs'] = [w['dz'], e(rows, self.H, dtype=AT), w['dy']]       # rotating dZ buffers of the backward
                w['z'] = None                                                    # never materialised
            self._train_ws = w
        return w

    # ---------------------------------------------------------------------------------
    # forward (rollout / critic_only): ActorCritic.rollout ml/actor_critic.py:74-96
    # ---------------------------------------------------------------------------------
    def forward_infer(self, obs, rows, rnn_states=None):
        """obs f32 [rows, D] -> head f32 [rows, NH] (logits | value).  Recurrent encoders update
        `rnn_states` ([c], [h]) in place."""
        w = self.infer_ws(rows)
        if self.tc:                          # two ping-pong activation buffers per tower
            return self._forward_tc(obs, rows, w, [w['y'][2 * (f // self.L) + (f & 1)] for f in range(self.LT)],
                                    None, None, rnn_states=rnn_states)
        feats = []
        for t in range(self.NT):
            x, d = obs, self.obs_dim
            for i in range(self.L):
                k, s, b = self.layer_views(self.params, t * self.L + i)
                y = w['y'][2 * t + (i & 1)]
                dense_ln_relu_f32(x, d, k, s, b, w['z'], y, None, rows, self.H, need_z=False)
                x, d = y, self.H
            feats.append(x)
        xs = feats if self.lstm is None else self.lstm.step_infer(feats[0], rows, rnn_states, w)
        return self._head_fwd_f32(xs, w['head'], rows)

    def _head_fwd_f32(self, xs, head, rows):
        """head = concat(xs) W + b; xs: the encoder's feature slices (one per LSTM layer, else the MLP output)."""
        W, B = self.head_views(self.params)
        fw = self.feat // len(xs)
        for l, x in enumerate(xs):
            gemm(x, W[l * fw:(l + 1) * fw], head, B if l == 0 else None, rows, self.NH, fw, fw, self.NH, self.NH,
                 accumulate=0 if l == 0 else 1)
        return head

    def _head_fwd_tc(self, xs, head, rows):
        _, B = self.head_views(self.params)
        fw = self.feat // len(xs)
        for l, x in enumerate(xs):               # K-slice l of wh_t [NH, feat]: pointer offset, ldb = feat
            Bt = self.wh_t if len(xs) == 1 else self.wh_t.view(-1)[l * fw:]
            gemm_tc(x, Bt, head, B if l == 0 else None, rows, self.NH, fw, fw, self.feat, self.NH, 0, 0,
                    0 if l == 0 else 2)
        return head

    def _forward_tc(self, obs, rows, w, ys, xhs, rstds, x_ready=False, seq=None, rnn_states=None):
        """bf16 tensor-core forward: cast obs -> L x fused [tcgen05 GEMM + LayerNorm + ReLU epilogue
        out of TMEM] -> head GEMM (fp32 out + bias).  Training also stashes xhat (bf16) and rstd."""
        if not x_ready:                    # x_ready: the minibatch gather already wrote the bf16 copy into w['x']
            call('mlb_cast_f32_bf16', ptr(obs), ptr(w['x']), c_ll(rows * self.obs_dim))
        xs = []
        for t in range(self.NT):
            x, d = w['x'], self.obs_dim
            for i in range(t * self.L, (t + 1) * self.L):
                _, s, b = self.layer_views(self.params, i)
                call('mlb_dense_ln_relu_fwd_tc', ptr(x), ptr(self.w_t[i]), ptr(s), ptr(b), ptr(ys[i]),
                     ptr(None if xhs is None else xhs[i]), ptr(None if rstds is None else rstds[i]),
                     c_int(rows), c_int(d), c_int(self.H), c_int(d), c_int(d))
                x, d = ys[i], self.H
            xs.append(x)
        if self.lstm is not None:
            if seq is not None:                       # training: the whole T' sequence
                xs = self.lstm.sequence_fwd(x, seq)
            else:                                     # rollout: one step, states updated in place
                xs = self.lstm.step_infer(x, rows, rnn_states, w)
        return self._head_fwd_tc(xs, w['head'], rows)

    @property
    def fused_rollout(self):
        """True when mlb_policy_rollout_tc covers this network (bf16, feed-forward MLP encoder that
        fits one CTA's shared memory); MLB_FUSED_ROLLOUT=0 keeps the layer-by-layer path."""
        if (not self.tc or self.lstm is not None or self.continuous is not None or self.NT != 1 or
                os.environ.get('MLB_FUSED_ROLLOUT', '1') == '0'):
            return False
        if self.H > 256 or self.H % 64 or self.L > 4 or self.NH > 256 or self.obs_dim > 256:
            return False
        panels = (max(self.H, self.obs_dim + 63)) // 64
        smem = panels * 16384 + 2 * self.H * 128 + 128 * (self.NH + 1) * 4 + (4 * self.H + 1024) * 4 + 1344
        return smem <= 227 * 1024

    def rollout_step_fused(self, obs, obs_store, rows, key_in, key_out, actions, log_probs, values,
                           partitionable=False, deterministic=False, head_out=None, post_step=None):
        """One launch: key chain + obs store copy + MLP + heads + sampling (mlb_policy_rollout_tc).
        post_step: (rewards, dones, reward_slab, done_slab, env_returns, trace, gamma) of the PREVIOUS simulator
        step -- its rollout-store write rides in this launch (mlb_policy_rollout_ps_tc)."""
        if getattr(self, '_tc_desc', None) is None or self._tc_desc_base != self.params.data_ptr():
            d = _lib.MlpTcDesc()
            d.num_layers, d.obs_dim, d.hidden, d.head_width = self.L, self.obs_dim, self.H, self.NH
            for i in range(self.L):
                _, s, b = self.layer_views(self.params, i)
                d.w_t[i], d.scale[i], d.bias[i] = self.w_t[i].data_ptr(), s.data_ptr(), b.data_ptr()
            _, B = self.head_views(self.params)
            d.wh_t, d.head_bias = self.wh_t.data_ptr(), B.data_ptr()
            self._tc_desc, self._tc_desc_base = d, self.params.data_ptr()
        ps = None
        if post_step is not None:
            r, d, rs, ds, er, tr, gamma = post_step
            ps = ctypes.byref(_lib.PostStep(r.data_ptr(), d.data_ptr(), rs.data_ptr(), ds.data_ptr(), er.data_ptr(),
                                            tr.data_ptr() if tr is not None else None, float(gamma)))
        call('mlb_policy_rollout_ps_tc', ctypes.byref(self._tc_desc), ptr(obs), ptr(obs_store), c_ll(rows),
             ptr(key_in), ptr(key_out), self._buckets_c, c_int(self.A), c_int(int(partitionable)),
             c_int(int(deterministic)), ptr(actions), ptr(log_probs), ptr(values), self._bins_c,
             c_int(self.V), ptr(head_out), ps)

    def sample(self, head, rows, policy_key, actions, log_probs, values, partitionable=False,
               deterministic=False):
        if self.continuous is not None:          # actions: fp32 bit patterns in the int32 buffer
            lo, hi, n = self.continuous
            call('mlb_sample_continuous_f32', ptr(head), c_int(self.NH), ptr(policy_key), c_int(n), c_float(lo),
                 c_float(hi), c_ll(rows), c_int(int(partitionable)), c_int(int(deterministic)), ptr(actions),
                 ptr(log_probs), ptr(values), self._bins_c, c_int(self.V))
            return
        call('mlb_sample_discrete_f32', ptr(head), c_int(self.NH), ptr(policy_key), self._buckets_c,
             c_int(self.A), c_ll(rows), c_int(int(partitionable)), c_int(int(deterministic)),
             ptr(actions), ptr(log_probs), ptr(values), self._bins_c, c_int(self.V))

    # ---------------------------------------------------------------------------------
    # training forward + backward (ActorCritic.update ml/actor_critic.py:98-128 + autodiff)
    # ---------------------------------------------------------------------------------
    def forward_train(self, obs, rows, seq=None, x_ready=False):
        """seq (recurrent encoders): dict(Tp, M, ends u8 [T', M], c0, h0 [M, RH])."""
        w = self.train_ws(rows)
        if self.tc:
            return self._forward_tc(obs, rows, w, w['y'], w['xh'], w['rstd'], x_ready, seq=seq)
        feats = []
        for t in range(self.NT):
            x, d = obs, self.obs_dim
            for i in range(t * self.L, (t + 1) * self.L):
                k, s, b = self.layer_views(self.params, i)
                dense_ln_relu_f32(x, d, k, s, b, w['z'][i], w['y'][i], w['stats'][i], rows, self.H)
                x, d = w['y'][i], self.H
            feats.append(x)
        xs = feats if self.lstm is None else self.lstm.sequence_fwd(feats[0], seq)
        return self._head_fwd_f32(xs, w['head'], rows)

    def backward(self, obs, rows, seq=None):
        """Consumes train_ws['dhead']; accumulates into self.grads (pre-zeroed)."""
        w = self.train_ws(rows)
        if self.tc:
            return self._backward_tc(rows, w, seq)
        W, B = self.head_views(self.params)
        gW, gB = self.head_views(self.grads)
        if self.lstm is not None:
            lws = self.lstm.train_ws(seq['Tp'], seq['M'])
            fw = self.lstm.RH
            pairs = [(lw['h_seq'].view(rows, fw), lw['d_hseq'].view(rows, fw)) for lw in lws]
        else:
            fw = self.H
            pairs = None
        if pairs is not None:
            # dW_h = feat^T dhead ; db_h = colsum(dhead) ; dfeat = dhead W_h^T   (per feature slice)
            for l, (feat, dfeat) in enumerate(pairs):
                gemm(feat, w['dhead'], gW[l * fw:(l + 1) * fw], None, fw, self.NH, rows, fw, self.NH, self.NH,
                     ta=1, tb=0, accumulate=1, splitk=_splitk_for(fw, self.NH, rows))
                gemm(w['dhead'], W[l * fw:(l + 1) * fw], dfeat, None, rows, fw, self.NH, self.NH, self.NH, fw,
                     ta=0, tb=1)
            self.lstm.sequence_bwd(w['y'][self.L - 1], seq, w['dy'])
        for t in range(self.NT):
            base = t * self.L
            if pairs is None:                 # this tower's slice of the head: dW_h, then dfeat -> w['dy']
                feat = w['y'][base + self.L - 1]
                gemm(feat, w['dhead'], gW[t * fw:(t + 1) * fw], None, fw, self.NH, rows, fw, self.NH, self.NH,
                     ta=1, tb=0, accumulate=1, splitk=_splitk_for(fw, self.NH, rows))
                gemm(w['dhead'], W[t * fw:(t + 1) * fw], w['dy'], None, rows, fw, self.NH, self.NH, self.NH, fw,
                     ta=0, tb=1)
            for i in range(base + self.L - 1, base - 1, -1):
                k, s, b = self.layer_views(self.params, i)
                gk, gs, gb = self.layer_views(self.grads, i)
                d = self.layer_off[i][2]
                call('mlb_ln_relu_bwd_f32', ptr(w['dy']), ptr(w['z'][i]), ptr(w['stats'][i]), ptr(s), ptr(b),
                     ptr(w['dz']), ptr(gs), ptr(gb), c_ll(rows), c_int(self.H))
                x = obs if i == base else w['y'][i - 1]
                gemm(x, w['dz'], gk, None, d, self.H, rows, d, self.H, self.H, ta=1, tb=0,
                     accumulate=1, splitk=_splitk_for(d, self.H, rows))
                if i > base:
                    gemm(w['dz'], k, w['dy'], None, rows, d, self.H, self.H, self.H, d, ta=0, tb=1)
        self._mask_head_grads()

    def _backward_tc(self, rows, w, seq=None):
        """bf16 tensor-core backward.  dW products are MN-major x MN-major split-K GEMMs with
        fp32 atomic accumulation straight into the gradient arena.  The dW GEMMs only depend on
        the dZ their layer's dx kernel produced, so they run on a side stream underneath the next
        dx kernel (MLB_BWD_STREAMS=0 serialises everything on one stream)."""
        two = os.environ.get('MLB_BWD_STREAMS', '0') != '0'      # measured: -0.4 % only, off by default
        main = torch.cuda.current_stream()
        if two and getattr(self, '_side', None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        side = self._side if two else main

        def on_side(fn):
            if not two:
                return fn()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                fn()

        gW, _ = self.head_views(self.grads)       # (head bias grads were accumulated by the loss kernel)
        lws = self.lstm.train_ws(seq['Tp'], seq['M']) if self.lstm is not None else None
        lw = None if lws is None else lws[0]
        fw = self.H if lws is None else self.lstm.RH
        feats = ([w['y'][t * self.L + self.L - 1] for t in range(self.NT)] if lws is None
                 else [x['h_seq'].view(rows, fw) for x in lws])
        for l, feat in enumerate(feats):
            on_side(lambda feat=feat, l=l: gemm_tc(feat, w['dhead'], gW[l * fw:(l + 1) * fw], None, fw, self.NH, rows,
                                                   fw, self.NH, self.NH, 1, 1, 2, _splitk_tc(fw, self.NH, rows)))
        bufs = w['dzs']                           # rotating dZ buffers (3: a dW may still read the oldest)
        cur = 0
        for t in range(self.NT):                  # BackboneSeparate: the critic tower after the actor tower
            base = t * self.L
            i = base + self.L - 1
            _, s, b = self.layer_views(self.params, i)
            _, gs, gb = self.layer_views(self.grads, i)
            if lw is None:
                # dfeat = dhead Wh^T (this tower's rows of Wh) fused with the LayerNorm/ReLU backward of the
                # tower's last layer -> dZ_{L-1}
                call('mlb_dense_dx_lnbwd_tc', ptr(w['dhead']), ptr(self.wh_c[t * self.H:(t + 1) * self.H]), ptr(s),
                     ptr(b), ptr(w['xh'][i]), ptr(w['rstd'][i]), ptr(bufs[cur]), ptr(gs), ptr(gb), c_int(rows),
                     c_int(self.NH), c_int(self.H), c_int(self.NH), c_int(self.NH))
            else:
                # d(encoder output) = dhead Wh^T (fp32), BPTT through the LSTM, then the gradient to the MLP output
                # (dz_all W_i) fused with the last layer's LayerNorm/ReLU backward
                RH4 = 4 * self.lstm.RH
                for l, x in enumerate(lws):       # rows l*RH.. of wh_c [feat, NH]: the slice's own B[N = RH, K = NH]
                    gemm_tc(w['dhead'], self.wh_c[l * fw:(l + 1) * fw], x['d_hseq'], None, rows, fw, self.NH, self.NH,
                            self.NH, fw, 0, 0, 0)
                self.lstm.sequence_bwd(w['y'][i], seq, None)
                call('mlb_dense_dx_lnbwd_tc', ptr(lw['dz']), ptr(self.lstm.wi_t), ptr(s), ptr(b), ptr(w['xh'][i]),
                     ptr(w['rstd'][i]), ptr(bufs[cur]), ptr(gs), ptr(gb), c_int(rows), c_int(RH4), c_int(self.H),
                     c_int(RH4), c_int(RH4))
            for i in range(base + self.L - 1, base - 1, -1):
                gk, _, _ = self.layer_views(self.grads, i)
                d = self.layer_off[i][2]
                x = w['x'] if i == base else w['y'][i - 1]
                dz_cur = bufs[cur]
                on_side(lambda x=x, dz_cur=dz_cur, gk=gk, d=d: gemm_tc(
                    x, dz_cur, gk, None, d, self.H, rows, d, self.H, self.H, 1, 1, 2, _splitk_tc(d, self.H, rows)))
                if i > base or t + 1 < self.NT:
                    nxt = (cur + 1) % len(bufs)
                    if two and len(bufs) < 3:
                        main.wait_stream(side)    # the buffer about to be overwritten may still be read
                    if i > base:
                        _, s, b = self.layer_views(self.params, i - 1)
                        _, gs, gb = self.layer_views(self.grads, i - 1)
                        call('mlb_dense_dx_lnbwd_tc', ptr(dz_cur), ptr(self.w_c[i]), ptr(s), ptr(b), ptr(w['xh'][i - 1]),
                             ptr(w['rstd'][i - 1]), ptr(bufs[nxt]), ptr(gs), ptr(gb), c_int(rows), c_int(self.H),
                             c_int(d), c_int(self.H), c_int(self.H))
                    cur = nxt
        if two:
            main.wait_stream(side)
        self._mask_head_grads()

    def _mask_head_grads(self):
        """BackboneSeparate: the head product runs over [actor features | critic features], so the dW GEMMs also
        fill the two off-diagonal blocks (actor features x critic columns and vice versa); clearing them keeps
        those weights at their initial zero and out of the gradient norm (ml/actor_critic.py:247-303: the two
        heads never see the other tower)."""
        if self.NT != 2:
            return
        gW, _ = self.head_views(self.grads)
        H, nA = self.H, self.sumA
        base = gW.data_ptr()                   # strided blocks: explicit addresses (ptr() takes contiguous buffers)
        call('mlb_fill_zero_2d', ctypes.c_void_p(base + 4 * nA), c_int(H), c_int(self.NH - nA), c_int(self.NH))
        call('mlb_fill_zero_2d', ctypes.c_void_p(base + 4 * H * self.NH), c_int(H), c_int(nA), c_int(self.NH))

    @property
    def loss_flags(self):
        """Extra mlb_ppo_loss_f32 flags for this program (bf16 d_head on the tensor-core path)."""
        return (4 if self.tc else 0) | (8 if self.hlgauss else 0) | (16 if self.continuous is not None else 0)

    def head_bias_grad(self):
        return self.head_views(self.grads)[1]

    def zero_grads(self):
        if getattr(self, '_grads_clean', False):      # cleared by the fused optimiser kernel of the last step
            self._grads_clean = False
            return
        call('mlb_fill_zero', ptr(self.grads), c_size_t(self.num_params * 4))

    def adopt_grad_arena(self, arena):
        """Move the gradient arena into caller-provided memory (e.g. NVLink symmetric memory for
        the fused data-parallel all-reduce); every gradient view is derived from self.grads."""
        assert arena.numel() >= self.num_params and arena.dtype == F32 and arena.is_contiguous()
        arena.zero_()
        self.grads = arena

    def optimizer_step(self, lr, max_grad_norm, grad_scale=1.0, b1=0.9, b2=0.999, eps=1e-8, reduced=None):
        """clip_by_global_norm -> adam -> re-projection / LN renorm (ml/ppo.py:283-338).
        reduced: the already all-reduced gradient whose sum of squares is in self.grad_sumsq
        (mlb_allreduce_sumsq_f32); otherwise the local arena is used and its norm computed here."""
        grads = self.grads if reduced is None else reduced
        if self._fused_opt:
            call('mlb_optimizer_step_fused', ptr(self.params), ptr(grads), ptr(self.adam_m), ptr(self.adam_v),
                 c_ll(self.num_params), ptr(self.segments), c_int(self.num_segments), ptr(self.copies),
                 ptr(self.adam_step), ptr(self.grad_sumsq), c_int(0 if reduced is None else 1), c_float(lr),
                 c_float(b1), c_float(b2), c_float(eps), c_float(max_grad_norm), c_float(grad_scale),
                 ptr(self._opt_sync), ptr(self._opt_ws), c_size_t(self._opt_ws.numel()),
                 ptr(self.grads if self._opt_zero else None))
            self._grads_clean = self._opt_zero    # the kernel cleared the arena behind itself
            if self.lstm is not None:
                self.lstm.pack()
            return
        if reduced is None:
            call('mlb_sumsq_f32', ptr(grads), c_ll(self.num_params), ptr(self.grad_sumsq),
                 ptr(self._sumsq_ws), c_size_t(self._sumsq_ws.numel()))
        call('mlb_adam_step_f32', ptr(self.params), ptr(grads), ptr(self.adam_m), ptr(self.adam_v),
             c_ll(self.num_params), ptr(self.adam_step), ptr(self.grad_sumsq), c_float(lr),
             c_float(b1), c_float(b2), c_float(eps), c_float(max_grad_norm), c_float(grad_scale))
        call('mlb_renorm_segments', ptr(self.params), ptr(self.segments), c_int(self.num_segments),
             ptr(self.adam_step), ptr(self.copies))
        if self.lstm is not None:
            self.lstm.pack()

    # ---------------------------------------------------------------------------------
    # flax-style apply(method=...) entry points (ml/actor_critic.py:65-128); these allocate
    # their outputs and are meant for API parity / tests -- the training loop calls the
    # buffer-explicit methods above.
    # ---------------------------------------------------------------------------------
    def _obs2d(self, obs):
        (ob,) = obs.values() if isinstance(obs, dict) else (obs,)
        return ob.reshape(-1, self.obs_dim)

    def apply_rollout(self, prng_key, rnn_states, obs, train=False, sample_actions=True,
                      return_debug=False, partitionable=False):
        x = self._obs2d(obs)
        rows = x.shape[0]
        head = self.forward_infer(x, rows, rnn_states)
        actions = torch.empty(rows, self.A, dtype=torch.int32, device=self.device)
        log_probs = torch.empty(rows, self.A, dtype=F32, device=self.device)
        values = torch.empty(rows, 1, dtype=F32, device=self.device)
        self.sample(head, rows, prng_key, actions, log_probs if sample_actions else None, values,
                    partitionable, deterministic=not sample_actions)
        name = self.groups[0][0]
        out = {'actions': {name: actions}, 'critic': values}
        if sample_actions:
            out['log_probs'] = {name: log_probs}
        return out, rnn_states

    def apply_critic_only(self, rnn_states, obs, train=False):
        x = self._obs2d(obs)
        head = self.forward_infer(x, x.shape[0], rnn_states)
        return {'critic': head[:, self.sumA:self.sumA + self.V].clone()}, rnn_states

    def apply_actor_only(self, rnn_states, obs, train=False):
        out, rnn = self.apply_rollout(None, rnn_states, obs, sample_actions=False)
        return {'actions': out['actions']}, rnn
