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
"""
import math

import numpy as np
import torch

from . import _lib
from ._lib import c_int, c_ll, call, ptr

F32 = torch.float32
GATES = ('i', 'f', 'g', 'o')


class LSTMLowering:
    def __init__(self, prog, rnn, in_dim, off):
        if rnn.num_layers != 1:
            raise NotImplementedError('LSTM with num_layers > 1: next (single-layer LSTM is lowered)')
        self.RH = int(rnn.num_hidden_channels)
        self.RL = 1
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
        return {'cell': {'OptimizedLSTMCell_0': cell}}

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
        return [{'wi': c(wi.t()), 'wh': c(wh.t()), 'bh': c(b)}]

    def segments(self, host, norms):
        """One re-projection segment per gate kernel (8 per layer)."""
        RH, d = self.RH, self.in_dim
        segs = []
        for g, name in enumerate(GATES):
            for key, off, n in (('i' + name, self.wi_off + g * RH * d, RH * d),
                                ('h' + name, self.wh_off + g * RH * RH, RH * RH)):
                k = f'lstm0/{key}'
                if k not in norms:
                    norms[k] = float(torch.linalg.vector_norm(host[off:off + n].double()))
                segs.append(_lib.Segment(off, n, 1, norms[k]))
        return segs

    # ---- state ------------------------------------------------------------------------
    def init_states(self, N, device):
        z = lambda: torch.zeros(N, self.RH, dtype=F32, device=device)
        return [z()], [z()]

    # ---- rollout step -------------------------------------------------------------------
    def step_infer(self, x, rows, states, z_buf, out):
        """x [rows, in] -> out [rows, RH]; states ([c], [h]) updated in place."""
        from .engine import gemm
        wi, wh, b = self.views(self.prog.params)
        c, h = states[0][0], states[1][0]
        RH, d = self.RH, self.in_dim
        gemm(x, wi, z_buf, None, rows, 4 * RH, d, d, d, 4 * RH, ta=0, tb=1)
        gemm(h, wh, z_buf, None, rows, 4 * RH, RH, RH, RH, 4 * RH, ta=0, tb=1, accumulate=1)
        call('mlb_lstm_cell_fwd_f32', ptr(z_buf), ptr(b), ptr(c), ptr(None), ptr(out), ptr(c), ptr(h),
             ptr(None), c_ll(rows), c_int(RH))
        return out

    def reset(self, states, dones, rows):
        for s in (states[0][0], states[1][0]):
            call('mlb_rnn_reset_f32', ptr(s), ptr(dones), c_ll(rows), c_int(self.RH))

    # ---- training sequence --------------------------------------------------------------
    def train_ws(self, Tp, M):
        w = self._ws
        if w is None or w['Tp'] != Tp or w['M'] < M:
            dev, RH = self.prog.device, self.RH
            e = lambda *s: torch.empty(*s, dtype=F32, device=dev)
            w = dict(Tp=Tp, M=M, z=e(Tp, M, 4 * RH), h_in=e(Tp + 1, M, RH), c_in=e(Tp + 1, M, RH),
                     h_seq=e(Tp, M, RH), stash=e(Tp, M, 5 * RH), d_hseq=e(Tp, M, RH),
                     dh=e(M, RH), dc=[e(M, RH), e(M, RH)])
            self._ws = w
        return w

    def sequence_fwd(self, feats, seq):
        """feats [T'*M, in]; seq: dict(Tp, M, ends u8 [T', M], c0 [M, RH], h0 [M, RH]).
        Returns h_seq [T'*M, RH] (the encoder output)."""
        from .engine import gemm
        Tp, M = seq['Tp'], seq['M']
        w = self.train_ws(Tp, M)
        wi, wh, b = self.views(self.prog.params)
        RH, d, rows = self.RH, self.in_dim, Tp * M
        gemm(feats, wi, w['z'], None, rows, 4 * RH, d, d, d, 4 * RH, ta=0, tb=1)
        call('mlb_copy_bytes', ptr(seq['c0']), ptr(w['c_in'][0]), _lib.c_size_t(M * RH * 4))
        call('mlb_copy_bytes', ptr(seq['h0']), ptr(w['h_in'][0]), _lib.c_size_t(M * RH * 4))
        for t in range(Tp):
            gemm(w['h_in'][t], wh, w['z'][t], None, M, 4 * RH, RH, RH, RH, 4 * RH, ta=0, tb=1, accumulate=1)
            call('mlb_lstm_cell_fwd_f32', ptr(w['z'][t]), ptr(b), ptr(w['c_in'][t]), ptr(seq['ends'][t]),
                 ptr(w['h_seq'][t]), ptr(w['c_in'][t + 1]), ptr(w['h_in'][t + 1]), ptr(w['stash'][t]),
                 c_ll(M), c_int(RH))
        return w['h_seq'].view(rows, RH)

    def sequence_bwd(self, feats, seq, dfeats):
        """Consumes ws['d_hseq'] [T', M, RH] (gradient w.r.t. the encoder output); accumulates the
        LSTM parameter gradients into the program's gradient arena and writes dfeats [T'*M, in]."""
        from .engine import _splitk_for, gemm
        Tp, M = seq['Tp'], seq['M']
        w = self.train_ws(Tp, M)
        wi, wh, b = self.views(self.prog.params)
        gwi, gwh, gb = self.views(self.prog.grads)
        RH, d, rows = self.RH, self.in_dim, Tp * M
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
        gemm(dz2, wi, dfeats, None, rows, d, 4 * RH, 4 * RH, d, d)
        return dfeats
