"""LSTM lowering for RecurrentBackboneEncoder (ml/actor_critic.py:156-199, ml/rnn.py:10-111).

Parameters live in the program's flat arena as, per layer, W_i^T [4*RH, in] | W_h^T [4*RH, RH] |
b_h [4*RH] -- gate blocks i, f, g, o stacked along rows, so that every flax leaf
(`OptimizedLSTMCell_l/{ii,if,ig,io,hi,hf,hg,ho}/kernel`, `/bias`) is ONE contiguous segment
(its transpose) and can be re-projected to its own initial norm by the optimiser
(ml/ppo.py:303-310 applies to every leaf named `kernel`).

Sequence forward (LSTM.sequence): the input projection of all T' steps is one GEMM; each step
then adds h_{t-1} W_h (accumulating GEMM) and runs the cell kernel, which writes the unmasked
output and the carry zeroed where the step ended an episode.  BPTT walks the steps backwards
with the cell backward kernel + one GEMM per step; dW_i, dW_h, db_h and the gradient to the
layer below are single GEMMs over all steps.

compute_dtype = bfloat16 (`prog.tc`): every one of those products runs on the tcgen05 tensor cores
(mlb_gemm_bf16_tc: bf16 operands, fp32 accumulation).  The cell state c, the gate pre-activations z
and the stash stay fp32; h, the encoder output and dz are written as bf16 by the cell kernels because
they are GEMM operands.  bf16 copies of W_i (both orientations) and W_h are refreshed by the fused
optimiser, like the Dense kernels.
"""
import math

import numpy as np
import torch

from . import _lib
from ._lib import c_int, c_ll, call, ptr

F32 = torch.float32
GATES = ('i', 'f', 'g', 'o')


class _LSTMLayer:
    """One OptimizedLSTMCell of the stack (layer `li`; input = the MLP output for li = 0, else h of li - 1)."""

    def __init__(self, prog, rnn, in_dim, off, li=0):
        self.li = int(li)
        self.RH = int(rnn.num_hidden_channels)
        self.in_dim = int(in_dim)
        if self.RH % 4:
            raise NotImplementedError('LSTM width must be a multiple of 4')
        self.wi_off = off
        off += 4 * self.RH * self.in_dim
        self.wh_off = off
        off += 4 * self.RH * self.RH
        self.b_off = off
        off += 4 * self.RH
        self.end_off = off
        self.prog = prog
        self._ws = None
        self.tc = bool(getattr(prog, 'tc', False))
        if self.tc:
            if self.RH % 64 or self.in_dim % 8:
                raise NotImplementedError('tensor-core LSTM needs num_hidden_channels % 64 == 0 and in_dim % 8 == 0')
            BF, dev, RH, d = torch.bfloat16, prog.device, self.RH, self.in_dim
            self.wi_c = torch.zeros(4 * RH, d, dtype=BF, device=dev)       # z = x W_i^T : B [N = 4RH, K = in]
            self.wi_t = torch.zeros(d, 4 * RH, dtype=BF, device=dev)       # dx = dz W_i : B [N = in, K = 4RH]
            self.wh_c = torch.zeros(4 * RH, RH, dtype=BF, device=dev)      # z += h W_h^T ; dh = dz W_h (MN-major B)
            self.wh_t = torch.zeros(RH, 4 * RH, dtype=BF, device=dev)      # (the optimiser writes both orientations)
            # fused step kernel (mlb_lstm_step_tc): [W_i | W_h] with the rows permuted so that one 256-column
            # accumulator unit holds the four gates of 64 hidden units; re-packed after every optimiser step
            import os
            self.fused = d % 64 == 0 and os.environ.get('MLB_LSTM_FUSED', '1') != '0'
            if self.fused:
                self.w_packed = torch.zeros(4 * RH, d + RH, dtype=BF, device=dev)
                self.b_packed = torch.zeros(4 * RH, dtype=F32, device=dev)

    # ---- parameter views ---------------------------------------------------------------
    def views(self, arena):
        RH, d = self.RH, self.in_dim
        wi = arena[self.wi_off:self.wi_off + 4 * RH * d].view(4 * RH, d)
        wh = arena[self.wh_off:self.wh_off + 4 * RH * RH].view(4 * RH, RH)
        b = arena[self.b_off:self.b_off + 4 * RH]
        return wi, wh, b

    def param_tree(self, arena):
        wi, wh, b = self.views(arena)
        RH = self.RH
        cell = {}
        for g, name in enumerate(GATES):
            cell['i' + name] = {'kernel': wi[g * RH:(g + 1) * RH].t()}
            cell['h' + name] = {'kernel': wh[g * RH:(g + 1) * RH].t(), 'bias': b[g * RH:(g + 1) * RH]}
        return cell

    def init_host(self, host, orth):
        RH, d = self.RH, self.in_dim
        for g in range(4):
            host[self.wi_off + g * RH * d:self.wi_off + (g + 1) * RH * d] = orth(d, RH, 1.0).t().reshape(-1)
            host[self.wh_off + g * RH * RH:self.wh_off + (g + 1) * RH * RH] = orth(RH, RH, 1.0).t().reshape(-1)

    def load_oracle(self, host, lyr):
        """oracle layout: wi [in, 4RH], wh [RH, 4RH], bh [4RH]."""
        host[self.wi_off:self.wi_off + 4 * self.RH * self.in_dim] = \
            torch.from_numpy(np.ascontiguousarray(np.asarray(lyr['wi'], np.float32).T)).reshape(-1)
        host[self.wh_off:self.wh_off + 4 * self.RH * self.RH] = \
            torch.from_numpy(np.ascontiguousarray(np.asarray(lyr['wh'], np.float32).T)).reshape(-1)
        host[self.b_off:self.b_off + 4 * self.RH] = torch.from_numpy(np.asarray(lyr['bh'], np.float32))

    def to_oracle(self, arena):
        wi, wh, b = self.views(arena)
        c = lambda x: x.detach().cpu().numpy().copy()
        return {'wi': c(wi.t()), 'wh': c(wh.t()), 'bh': c(b)}

    def segments(self, host, norms):
        """One re-projection segment per gate kernel (8 per layer)."""
        RH, d = self.RH, self.in_dim
        segs = []
        for g, name in enumerate(GATES):
            for key, off, n in (('i' + name, self.wi_off + g * RH * d, RH * d),
                                ('h' + name, self.wh_off + g * RH * RH, RH * RH)):
                k = f'lstm{self.li}/{key}'
                if k not in norms:
                    norms[k] = float(torch.linalg.vector_norm(host[off:off + n].double()))
                segs.append(_lib.Segment(off, n, 1, norms[k]))
        return segs

    def refresh_bf16(self):
        wi, wh, _ = self.views(self.prog.params)
        RH, d = self.RH, self.in_dim
        call('mlb_cast_weight_bf16', ptr(wi), ptr(self.wi_t), ptr(self.wi_c), c_int(4 * RH), c_int(d), c_int(d),
             c_int(4 * RH), c_int(d))
        call('mlb_cast_weight_bf16', ptr(wh), ptr(self.wh_t), ptr(self.wh_c), c_int(4 * RH), c_int(RH), c_int(RH),
             c_int(4 * RH), c_int(RH))
        self.pack()

    def pack(self):
        """Refresh the fused step kernel's packed operand from the fp32 master weights."""
        if self.tc and self.fused:
            wi, wh, b = self.views(self.prog.params)
            call('mlb_lstm_pack_weights_bf16', ptr(wi), ptr(wh), ptr(b), ptr(self.w_packed), ptr(self.b_packed),
                 c_int(self.in_dim), c_int(self.RH))

    def bf16_copies(self):
        """One mlb_bf16_copy per segment of segments(), same order: gate block g of W_i is the [RH, in]
        row block g of wi_c (and the column block g of wi_t); likewise W_h."""
        RH, d = self.RH, self.in_dim
        out = []
        for g in range(4):
            out.append(_lib.Bf16Copy(self.wi_t.data_ptr() + 2 * g * RH, self.wi_c.data_ptr() + 2 * g * RH * d,
                                     RH, d, 4 * RH, d))
            out.append(_lib.Bf16Copy(self.wh_t.data_ptr() + 2 * g * RH, self.wh_c.data_ptr() + 2 * g * RH * RH,
                                     RH, RH, 4 * RH, RH))
        return out

    # ---- state ------------------------------------------------------------------------
    def init_states(self, N, device):
        z = lambda: torch.zeros(N, self.RH, dtype=F32, device=device)
        return z(), z()

    # ---- rollout step -------------------------------------------------------------------
    def step_infer(self, x, rows, states, z_buf, out):
        """x [rows, in] -> out [rows, RH]; states ([c], [h]) updated in place."""
        from .engine import gemm, gemm_tc
        wi, wh, b = self.views(self.prog.params)
        c, h = states[0][self.li], states[1][self.li]
        RH, d = self.RH, self.in_dim
        if self.tc:                    # x, out: bf16; states stay fp32 (h is cast for the recurrent GEMM)
            hb = self._h_bf(rows)
            call('mlb_cast_f32_bf16', ptr(h), ptr(hb), c_ll(rows * RH))
            if self.fused:             # one launch: [x | h] GEMM with the cell math in the epilogue
                call('mlb_lstm_step_tc', ptr(x), c_int(d), ptr(hb), ptr(self.w_packed), ptr(self.b_packed), ptr(c),
                     ptr(None), ptr(out), ptr(c), ptr(h), ptr(None), ptr(None), c_ll(rows), c_int(d), c_int(RH))
                return out
            gemm_tc(x, self.wi_c, z_buf, b, rows, 4 * RH, d, d, d, 4 * RH, 0, 0, 0)
            gemm_tc(hb, self.wh_c, z_buf, None, rows, 4 * RH, RH, RH, RH, 4 * RH, 0, 0, 2)
            call('mlb_lstm_cell_fwd_tc', ptr(z_buf), ptr(None), ptr(c), ptr(None), ptr(out), ptr(c), ptr(h),
                 ptr(None), ptr(None), c_ll(rows), c_int(RH))
            return out
        gemm(x, wi, z_buf, None, rows, 4 * RH, d, d, d, 4 * RH, ta=0, tb=1)
        gemm(h, wh, z_buf, None, rows, 4 * RH, RH, RH, RH, 4 * RH, ta=0, tb=1, accumulate=1)
        call('mlb_lstm_cell_fwd_f32', ptr(z_buf), ptr(b), ptr(c), ptr(None), ptr(out), ptr(c), ptr(h),
             ptr(None), c_ll(rows), c_int(RH))
        return out

    def _h_bf(self, rows):
        hb = getattr(self, '_hb', None)
        if hb is None or hb.shape[0] < rows:
            hb = self._hb = torch.empty(rows, self.RH, dtype=torch.bfloat16, device=self.prog.device)
        return hb

    def reset(self, states, dones, rows):
        for s in (states[0][self.li], states[1][self.li]):
            call('mlb_rnn_reset_f32', ptr(s), ptr(dones), c_ll(rows), c_int(self.RH))

    # ---- training sequence --------------------------------------------------------------
    def train_ws(self, Tp, M):
        w = self._ws
        if w is None or w['Tp'] != Tp or w['M'] < M:
            dev, RH = self.prog.device, self.RH
            e = lambda *s: torch.empty(*s, dtype=F32, device=dev)
            HT = torch.bfloat16 if self.tc else F32      # h / encoder output: GEMM operands on the tc path
            w = dict(Tp=Tp, M=M, z=None if (self.tc and self.fused) else e(Tp, M, 4 * RH), c_in=e(Tp + 1, M, RH),
                     h_in=torch.empty(Tp + 1, M, RH, dtype=HT, device=dev),
                     h_seq=torch.empty(Tp, M, RH, dtype=HT, device=dev), stash=e(Tp, M, 5 * RH),
                     d_hseq=e(Tp, M, RH), dh=e(M, RH), dc=[e(M, RH), e(M, RH)])
            if self.tc:
                w['dz'] = torch.empty(Tp, M, 4 * RH, dtype=HT, device=dev)
            self._ws = w
        return w

    def _load_start(self, seq, w):
        """c_in[0], h_in[0] <- this layer's chunk-start state (a strided feature slice when RL > 1)."""
        c0, h0 = seq['c0'][self.li], seq['h0'][self.li]
        M, RH = seq['M'], self.RH
        if c0.is_contiguous() and h0.is_contiguous():
            call('mlb_copy_bytes', ptr(c0), ptr(w['c_in'][0]), _lib.c_size_t(M * RH * 4))
            if self.tc:
                call('mlb_cast_f32_bf16', ptr(h0), ptr(w['h_in'][0]), c_ll(M * RH))
            else:
                call('mlb_copy_bytes', ptr(h0), ptr(w['h_in'][0]), _lib.c_size_t(M * RH * 4))
        else:
            w['c_in'][0].copy_(c0)
            w['h_in'][0].copy_(h0)

    def sequence_fwd(self, feats, seq):
        """feats [T'*M, in]; seq: dict(Tp, M, ends u8 [T', M], c0 / h0: per-layer lists of [M, RH]).
        Returns h_seq [T'*M, RH] (the encoder output)."""
        from .engine import gemm, gemm_tc
        Tp, M = seq['Tp'], seq['M']
        w = self.train_ws(Tp, M)
        wi, wh, b = self.views(self.prog.params)
        RH, d, rows = self.RH, self.in_dim, Tp * M
        if self.tc and self.fused:
            self._load_start(seq, w)
            f3 = feats.view(Tp, M, d)
            for t in range(Tp):
                call('mlb_lstm_step_tc', ptr(f3[t]), c_int(d), ptr(w['h_in'][t]), ptr(self.w_packed), ptr(self.b_packed),
                     ptr(w['c_in'][t]), ptr(seq['ends'][t]), ptr(w['h_seq'][t]), ptr(w['c_in'][t + 1]), ptr(None),
                     ptr(w['h_in'][t + 1]), ptr(w['stash'][t]), c_ll(M), c_int(d), c_int(RH))
            return w['h_seq'].view(rows, RH)
        if self.tc:
            gemm_tc(feats, self.wi_c, w['z'], b, rows, 4 * RH, d, d, d, 4 * RH, 0, 0, 0)     # + bias in the epilogue
            self._load_start(seq, w)
            for t in range(Tp):
                gemm_tc(w['h_in'][t], self.wh_c, w['z'][t], None, M, 4 * RH, RH, RH, RH, 4 * RH, 0, 0, 2)
                call('mlb_lstm_cell_fwd_tc', ptr(w['z'][t]), ptr(None), ptr(w['c_in'][t]), ptr(seq['ends'][t]),
                     ptr(w['h_seq'][t]), ptr(w['c_in'][t + 1]), ptr(None), ptr(w['h_in'][t + 1]), ptr(w['stash'][t]),
                     c_ll(M), c_int(RH))
            return w['h_seq'].view(rows, RH)
        gemm(feats, wi, w['z'], None, rows, 4 * RH, d, d, d, 4 * RH, ta=0, tb=1)
        self._load_start(seq, w)
        for t in range(Tp):
            gemm(w['h_in'][t], wh, w['z'][t], None, M, 4 * RH, RH, RH, RH, 4 * RH, ta=0, tb=1, accumulate=1)
            call('mlb_lstm_cell_fwd_f32', ptr(w['z'][t]), ptr(b), ptr(w['c_in'][t]), ptr(seq['ends'][t]),
                 ptr(w['h_seq'][t]), ptr(w['c_in'][t + 1]), ptr(w['h_in'][t + 1]), ptr(w['stash'][t]),
                 c_ll(M), c_int(RH))
        return w['h_seq'].view(rows, RH)

    def sequence_bwd(self, feats, seq, dfeats, accumulate=False):
        """Consumes ws['d_hseq'] [T', M, RH] (gradient w.r.t. this layer's output); accumulates the
        LSTM parameter gradients into the program's gradient arena and writes (accumulate: adds to)
        dfeats [T'*M, in], the gradient w.r.t. the layer's input."""
        from .engine import _splitk_for, gemm
        Tp, M = seq['Tp'], seq['M']
        w = self.train_ws(Tp, M)
        wi, wh, b = self.views(self.prog.params)
        gwi, gwh, gb = self.views(self.prog.grads)
        RH, d, rows = self.RH, self.in_dim, Tp * M
        if self.tc:
            return self._sequence_bwd_tc(feats, seq, dfeats, w, gwi, gwh, gb, accumulate)
        dz = w['z']                                   # pre-activations are dead: reuse as dz_all
        for t in range(Tp - 1, -1, -1):
            last = t == Tp - 1
            dc_in, dc_out = w['dc'][t & 1], w['dc'][(t & 1) ^ 1]
            call('mlb_lstm_cell_bwd_f32', ptr(w['d_hseq'][t]), c_int(RH), ptr(None if last else w['dh']),
                 ptr(None if last else dc_in), ptr(seq['ends'][t]), ptr(w['stash'][t]), ptr(w['c_in'][t]),
                 ptr(dz[t]), ptr(dc_out), c_ll(M), c_int(RH))
            if t > 0:                                 # dh_prev = dz_t W_h  (W_h^T stored [4RH, RH])
                gemm(dz[t], wh, w['dh'], None, M, RH, 4 * RH, 4 * RH, RH, RH)
        dz2 = dz.view(rows, 4 * RH)
        gemm(dz2, w['h_in'].view(-1, RH), gwh, None, 4 * RH, RH, rows, 4 * RH, RH, RH, ta=1, tb=0, accumulate=1,
             splitk=_splitk_for(4 * RH, RH, rows))
        gemm(dz2, feats, gwi, None, 4 * RH, d, rows, 4 * RH, d, d, ta=1, tb=0, accumulate=1,
             splitk=_splitk_for(4 * RH, d, rows))
        call('mlb_colsum_f32', ptr(dz2), c_ll(rows), c_int(4 * RH), c_int(4 * RH), ptr(gb))
        gemm(dz2, wi, dfeats, None, rows, d, 4 * RH, 4 * RH, d, d, accumulate=1 if accumulate else 0)
        return dfeats

    def _sequence_bwd_tc(self, feats, seq, dfeats, w, gwi, gwh, gb, accumulate=False):
        """Tensor-core BPTT.  feats bf16 [T'*M, in]; consumes w['d_hseq'] (f32).  dfeats None (layer 0): the
        gradient to the MLP output, dz_all W_i, is fused into the last Dense layer's LayerNorm-backward kernel by
        the caller (mlb_dense_dx_lnbwd_tc with DZ_in = w['dz'], W = wi_t); otherwise (upper layers) it is ADDED
        to dfeats = the f32 d_hseq of the layer below."""
        from .engine import _splitk_tc, gemm_tc
        Tp, M = seq['Tp'], seq['M']
        RH, d, rows = self.RH, self.in_dim, Tp * M
        dz = w['dz']
        for t in range(Tp - 1, -1, -1):
            last = t == Tp - 1
            dc_in, dc_out = w['dc'][t & 1], w['dc'][(t & 1) ^ 1]
            call('mlb_lstm_cell_bwd_tc', ptr(w['d_hseq'][t]), c_int(RH), ptr(None if last else w['dh']),
                 ptr(None if last else dc_in), ptr(seq['ends'][t]), ptr(w['stash'][t]), ptr(w['c_in'][t]),
                 ptr(dz[t]), ptr(dc_out), c_ll(M), c_int(RH))
            if t > 0:                                 # dh_prev = dz_t W_h : B = W_h^T stored [K = 4RH, N = RH] (MN-major)
                gemm_tc(dz[t], self.wh_c, w['dh'], None, M, RH, 4 * RH, 4 * RH, RH, RH, 0, 1, 0)
        dz2 = dz.view(rows, 4 * RH)
        # dW_h^T [4RH, RH] += dz^T h_in ; dW_i^T [4RH, in] += dz^T feats  (MN-major x MN-major, split-K atomics)
        gemm_tc(dz2, w['h_in'].view(-1, RH), gwh, None, 4 * RH, RH, rows, 4 * RH, RH, RH, 1, 1, 2,
                _splitk_tc(4 * RH, RH, rows))
        gemm_tc(dz2, feats, gwi, None, 4 * RH, d, rows, 4 * RH, d, d, 1, 1, 2, _splitk_tc(4 * RH, d, rows))
        call('mlb_colsum_bf16', ptr(dz2), c_ll(rows), c_int(4 * RH), c_int(4 * RH), ptr(gb))
        if dfeats is not None:                    # += dz_all W_i : B = W_i^T stored [K = 4RH, N = in] (MN-major)
            assert accumulate
            gemm_tc(dz2, self.wi_c, dfeats, None, rows, d, 4 * RH, 4 * RH, d, d, 0, 1, 2)
        return None


class LSTMLowering:
    """The LSTM stack of a RecurrentBackboneEncoder (ml/rnn.py:10-111): `num_layers` cells, layer l fed by the
    hidden state of layer l - 1, encoder output = the concatenation of every layer's hidden state
    (MultiLayerLSTMCell.__call__ :27-45).  The concatenation is never materialised: the heads read the layers'
    outputs as RL separate K-slices of the head GEMM (and scatter their gradient the same way)."""

    def __init__(self, prog, rnn, in_dim, off):
        self.RH = int(rnn.num_hidden_channels)
        self.RL = int(rnn.num_layers)
        if self.RL < 1:
            raise NotImplementedError('LSTM needs num_layers >= 1')
        self.prog = prog
        self.layers = []
        d = int(in_dim)
        for li in range(self.RL):
            lyr = _LSTMLayer(prog, rnn, d, off, li)
            self.layers.append(lyr)
            off, d = lyr.end_off, self.RH
        self.end_off = off
        self.in_dim = int(in_dim)
        self.tc = self.layers[0].tc

    # single-layer shorthands (tests, the tensor-core backward's fused dx call)
    def views(self, arena, li=0):
        return self.layers[li].views(arena)

    def __getattr__(self, name):
        if name in ('wi_c', 'wi_t', 'wh_c', 'wh_t', 'w_packed', 'b_packed', 'fused'):
            return getattr(self.layers[0], name)
        raise AttributeError(name)

    def param_tree(self, arena):
        return {'cell': {f'OptimizedLSTMCell_{l.li}': l.param_tree(arena) for l in self.layers}}

    def init_host(self, host, orth):
        for l in self.layers:
            l.init_host(host, orth)

    def load_oracle(self, host, lstm):
        assert len(lstm) == self.RL
        for l, lyr in zip(self.layers, lstm):
            l.load_oracle(host, lyr)

    def to_oracle(self, arena):
        return [l.to_oracle(arena) for l in self.layers]

    def segments(self, host, norms):
        return [sg for l in self.layers for sg in l.segments(host, norms)]

    def bf16_copies(self):
        return [c for l in self.layers for c in l.bf16_copies()]

    def refresh_bf16(self):
        for l in self.layers:
            l.refresh_bf16()

    def pack(self):
        for l in self.layers:
            l.pack()

    def init_states(self, N, device):
        st = [l.init_states(N, device) for l in self.layers]
        return [c for c, _ in st], [h for _, h in st]

    def reset(self, states, dones, rows):
        for l in self.layers:
            l.reset(states, dones, rows)

    def step_infer(self, x, rows, states, ws):
        """One rollout step through the stack; returns the per-layer outputs [rows, RH] (ws: the caller's
        workspace dict, holds the per-layer scratch)."""
        outs = []
        OT = torch.bfloat16 if self.tc else F32
        for l in self.layers:
            key = f'rout{l.li}'
            if key not in ws or ws[key].shape[0] < rows:
                ws[key] = torch.empty(rows, self.RH, dtype=OT, device=self.prog.device)
                ws['rz'] = torch.empty(rows, 4 * self.RH, dtype=F32, device=self.prog.device)
            x = l.step_infer(x, rows, states, ws['rz'], ws[key])
            outs.append(x)
        return outs

    def train_ws(self, Tp, M):
        return [l.train_ws(Tp, M) for l in self.layers]

    @staticmethod
    def _seq(seq):
        """c0 / h0 as per-layer lists (a bare tensor is accepted for a single layer)."""
        if isinstance(seq['c0'], (list, tuple)):
            return seq
        return {**seq, 'c0': [seq['c0']], 'h0': [seq['h0']]}

    def sequence_fwd(self, feats, seq):
        seq = self._seq(seq)
        outs = []
        for l in self.layers:
            feats = l.sequence_fwd(feats, seq)
            outs.append(feats)
        return outs

    def sequence_bwd(self, feats, seq, dfeats):
        """Top layer first: each layer adds the gradient w.r.t. its input to the d_hseq of the layer below
        (already holding the heads' contribution); layer 0 writes dfeats (f32 path) or leaves dz for the fused
        dx kernel (tensor-core path, dfeats None)."""
        seq = self._seq(seq)
        lws = self.train_ws(seq['Tp'], seq['M'])
        rows = seq['Tp'] * seq['M']
        for li in range(self.RL - 1, -1, -1):
            l = self.layers[li]
            if li == 0:
                l.sequence_bwd(feats, seq, dfeats)
            else:
                below = lws[li - 1]
                l.sequence_bwd(below['h_seq'].view(rows, self.RH), seq, below['d_hseq'].view(rows, self.RH),
                               accumulate=True)
