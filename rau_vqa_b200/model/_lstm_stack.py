"""Shared implementation of the two LSTM cell factories (model/ATTLSTM.lua, model/DeepLSTM.lua): a stack of
layers, each one `rau_lstm_cell_fwd/bwd` call, with nn.Dropout between them (`rau_dropout`)."""
from __future__ import annotations

import itertools

import torch

from .. import core, nn
from .._ffi import check, ffi

_stream_counter = itertools.count(1)


class LSTMStack(nn.Container):
    def __init__(self, input_size, rnn_size, num_layers, dropout, gate_order, packed_state, dropout_on_first):
        super().__init__()
        self.input_size, self.rnn_size, self.num_layers, self.dropout = input_size, rnn_size, num_layers, float(dropout)
        self.gate_order, self.packed_state, self.dropout_on_first = gate_order, packed_state, dropout_on_first
        for L in range(num_layers):
            self.modules.append(nn.Linear(input_size if L == 0 else rnn_size, 4 * rnn_size))   # i2h  (A:6 / D:43)
            self.modules.append(nn.Linear(rnn_size, 4 * rnn_size))                              # h2h  (A:7 / D:44)
        self.noise_override = None      # optional list of uint8 keep masks per dropout site (parity tests)
        self._stream = 0
        self._saved = None

    # ---- state layout: ATTLSTM = separate c/h tensors narrowed per layer (A:43-44); DeepLSTM = [c1|h1|c2|h2] (D:23-24)
    def _views(self, state, L):
        H = self.rnn_size
        if self.packed_state:
            return state[:, 2 * L * H: 2 * L * H + H], state[:, 2 * L * H + H: 2 * (L + 1) * H]
        c, h = state
        return c[:, L * H:(L + 1) * H], h[:, L * H:(L + 1) * H]

    def _has_dropout(self, L):
        return self.dropout > 0 and (L > 0 or self.dropout_on_first)

    def _drop(self, ctx, x, L, site_out):
        if not self._has_dropout(L):
            return x
        mask = None if self.noise_override is None else self.noise_override[L]
        check(ctx.lib.rau_dropout(ctx.h, x.numel(), core.fptr(x), self.dropout, int(self.train), core.bptr(mask),
                                  self._stream * 64 + L, core.fptr(site_out)))
        return site_out

    def _desc(self, B, in_size, ldx, ldc_prev, ldh_prev, ldc, ldh):
        d = ffi.new("rau_lstm_desc*")
        d.B, d.in_size, d.H, d.gate_order = B, in_size, self.rnn_size, self.gate_order
        d.ldx, d.ldc_prev, d.ldh_prev, d.ldc, d.ldh = ldx, ldc_prev, ldh_prev, ldc, ldh
        return d

    def _split_input(self, input):
        if self.packed_state:
            x, state = input
            return x, state
        x, c, h = input
        return x, (c, h)

    def updateOutput(self, input):
        x, state = self._split_input(input)
        x = x.contiguous()
        B, H, n = x.shape[0], self.rnn_size, self.num_layers
        ctx = nn.context(x.device)
        self._stream = next(_stream_counter)
        f = dict(dtype=torch.float32, device=x.device)
        if self.packed_state:
            out = torch.empty(B, 2 * n * H, **f)
            out_state = out
        else:
            out_state = (torch.empty(B, n * H, **f), torch.empty(B, n * H, **f))
            out = list(out_state)
        saved = []
        u = x
        for L in range(n):
            in_size = self.input_size if L == 0 else H
            ud = self._drop(ctx, u, L, torch.empty(B, in_size, **f)) if u.is_contiguous() else None
            if ud is None:   # u is a strided view of the previous layer's h
                ud = self._drop(ctx, u.contiguous(), L, torch.empty(B, in_size, **f))
            ud = ud if ud.is_contiguous() else ud.contiguous()
            c_prev, h_prev = self._views(state, L)
            c_new, h_new = self._views(out_state, L)
            sv = torch.empty(5, B, H, **f)
            d = self._desc(B, in_size, ud.stride(0), c_prev.stride(0), h_prev.stride(0), c_new.stride(0), h_new.stride(0))
            i2h, h2h = self.modules[2 * L], self.modules[2 * L + 1]
            check(ctx.lib.rau_lstm_cell_fwd(ctx.h, d, _p(ud), _p(c_prev), _p(h_prev), _p(i2h.weight), _p(i2h.bias),
                                            _p(h2h.weight), _p(h2h.bias), _p(c_new), _p(h_new), _p(sv)))
            saved.append((ud, sv))
            u = h_new
        self._saved = saved
        self.output = out
        return self.output

    def _bwd(self, input, gradOutput, scale, want_input, want_params):
        x, state = self._split_input(input)
        B, H, n = x.shape[0], self.rnn_size, self.num_layers
        ctx = nn.context(x.device)
        f = dict(dtype=torch.float32, device=x.device)
        if self.packed_state:
            g_state = gradOutput
            d_state = torch.zeros(B, 2 * n * H, **f)
        else:
            g_state = (gradOutput[0], gradOutput[1])
            d_state = (torch.zeros(B, n * H, **f), torch.zeros(B, n * H, **f))
        dx_from_above = None
        dx = None
        for L in range(n - 1, -1, -1):
            in_size = self.input_size if L == 0 else H
            ud, sv = self._saved[L]
            c_prev, h_prev = self._views(state, L)
            dc_out, dh_out = self._views(g_state, L)
            dc_prev, dh_prev = self._views(d_state, L)
            du = torch.empty(B, in_size, **f)
            d = self._desc(B, in_size, ud.stride(0), c_prev.stride(0), h_prev.stride(0), H, H)
            i2h, h2h = self.modules[2 * L], self.modules[2 * L + 1]
            gp = (lambda t: _p(t)) if want_params else (lambda t: ffi.NULL)
            check(ctx.lib.rau_lstm_cell_bwd(ctx.h, d, _p(ud), _p(c_prev), _p(h_prev), _p(i2h.weight), _p(h2h.weight), _p(sv),
                                            _p(dc_out), _p(dh_out), dc_out.stride(0), dh_out.stride(0), _p(dx_from_above),
                                            _p(du), _p(dc_prev), _p(dh_prev), du.stride(0), dc_prev.stride(0), dh_prev.stride(0),
                                            gp(i2h.gradWeight), gp(i2h.gradBias), gp(h2h.gradWeight), gp(h2h.gradBias),
                                            float(scale)))
            if self._has_dropout(L):      # Dropout backward = same mask on the gradient
                mask = None if self.noise_override is None else self.noise_override[L]
                dud = torch.empty(B, in_size, **f)
                check(ctx.lib.rau_dropout(ctx.h, du.numel(), _p(du), self.dropout, int(self.train), core.bptr(mask),
                                          self._stream * 64 + L, _p(dud)))
                du = dud
            if L > 0:
                dx_from_above = du
            else:
                dx = du
        if want_input:
            self.gradInput = [dx, d_state] if self.packed_state else [dx, d_state[0], d_state[1]]
        return self.gradInput

    def updateGradInput(self, input, gradOutput):
        return self._bwd(input, gradOutput, 0.0, True, False)

    def accGradParameters(self, input, gradOutput, scale=1.0):
        self._bwd(input, gradOutput, scale, False, True)

    def backward(self, input, gradOutput, scale=1.0):
        return self._bwd(input, gradOutput, scale, True, True)


def _p(t):
    """float* of a float32 CUDA tensor whose rows are contiguous (row pitch passed separately)."""
    if t is None:
        return ffi.NULL
    assert t.dtype == torch.float32 and (t.dim() < 2 or t.stride(-1) == 1)
    return ffi.cast("float*", t.data_ptr())
