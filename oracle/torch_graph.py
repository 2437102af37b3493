"""Independent second derivation of the RAU step used ONLY to validate oracle/rau_oracle.py.

TEST INFRASTRUCTURE ONLY (see the header of rau_oracle.py).  The graph is rebuilt here out of
torch.nn modules (the PyTorch descendants of the Torch7 nn modules the reference composes with
nngraph), one module per reference node, in float64, and differentiated with torch.autograd.
The oracle's hand-written backward passes must agree with this to ~1e-10.

Cited lines: F: experiments/Ours_Full/LstmAttCtrlGradNoiseDontSelect.lua, A: model/ATTLSTM.lua,
D: model/DeepLSTM.lua.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as Fn

from . import rau_oracle as O


def _lin(w, b):
    m = nn.Linear(w.shape[1], w.shape[0]).double()
    m.weight = nn.Parameter(w)
    m.bias = nn.Parameter(b)
    return m


class TorchRAU:
    """Holds leaf tensors for the three flat groups and evaluates feval's loss with autograd."""

    def __init__(self, cfg: O.RauConfig, params: dict):
        self.cfg = cfg
        self.flat = {g: torch.tensor(np.asarray(params[g], dtype=np.float64), requires_grad=True) for g in O.GROUPS}
        self.P = {}
        for g in O.GROUPS:
            off = 0
            for name, shp in O.group_shapes(cfg, g):
                n = int(np.prod(shp))
                self.P[name] = self.flat[g][off:off + n].view(*shp)
                off += n

    @staticmethod
    def _drop(x, mask, p):
        if mask is None or p <= 0:
            return x
        return x * torch.as_tensor(mask, dtype=torch.float64) / (1.0 - p)   # nn.Dropout v2

    def word_embed(self, x_t, mask):
        e = Fn.embedding(torch.as_tensor(np.asarray(x_t), dtype=torch.long) - 1, self.P["E"])  # F:204
        return torch.tanh(self._drop(e, mask, self.cfg.p_embed))                              # F:205-206

    def deeplstm(self, x, s, mask):
        cfg, P, H = self.cfg, self.P, self.cfg.Hq
        outs, u = [], x
        for L in range(1, cfg.nlayer + 1):
            prev_c = s.narrow(1, 2 * (L - 1) * H, H)          # D:23
            prev_h = s.narrow(1, 2 * (L - 1) * H + H, H)      # D:24
            if L > 1:
                u = self._drop(outs[-1], mask, cfg.p_rnn)     # D:38-39
            sums = Fn.linear(u, P[f"l{L}.Wi"], P[f"l{L}.bi"]) + Fn.linear(prev_h, P[f"l{L}.Wh"], P[f"l{L}.bh"])  # D:43-45
            sg = torch.sigmoid(sums.narrow(1, 0, 3 * H))      # D:47-48
            i, f, o = sg.narrow(1, 0, H), sg.narrow(1, H, H), sg.narrow(1, 2 * H, H)  # D:49-51
            g = torch.tanh(sums.narrow(1, 3 * H, H))          # D:53-54
            c = f * prev_c + i * g                            # D:56-59
            h = o * torch.tanh(c)                             # D:61
            outs += [c, h]
        return torch.cat(outs, 1)                             # D:68

    def attlstm(self, x, c, h):
        P, H = self.P, self.cfg.H
        gates = Fn.linear(x, P["Wx"], P["bx"]) + Fn.linear(h, P["Whh"], P["bhh"])    # A:6-8
        r = gates.view(-1, 4, H)                              # A:12
        i = torch.sigmoid(r[:, 0]); g = torch.tanh(r[:, 1])   # A:16-17
        f = torch.sigmoid(r[:, 2]); o = torch.sigmoid(r[:, 3])  # A:18-19
        c2 = f * c + i * g                                    # A:21-24
        return c2, o * torch.tanh(c2)                         # A:25

    def hop(self, q, X, c, h, masks):
        cfg, P = self.cfg, self.P
        mq = mX = mm = None
        if masks is not None:
            mq, mX, mm = masks.get("q"), masks.get("X"), masks.get("m")
        B = q.shape[0]
        gw, gh = (14, 14) if cfg.S == 196 else (cfg.S, 1)     # cnnout_w x cnnout_h (F:217-219); toy grids are S x 1
        qf = torch.tanh(Fn.linear(self._drop(q, mq, cfg.p_q), P["Wq"], P["bq"]) + Fn.linear(h, P["Wh"], P["bh"]))  # F:233-235
        Xd = self._drop(X.reshape(B, cfg.C, gw, gh), None if mX is None else np.asarray(mX).reshape(B, cfg.C, gw, gh), cfg.p_x)
        I = torch.tanh(Fn.conv2d(Xd, P["Wi"].view(cfg.M, cfg.C, 1, 1), P["bi"])).reshape(B, cfg.M, cfg.S)  # F:239-242
        qatt = Fn.linear(qf, P["Wqa"], P["bqa"]).unsqueeze(2).expand(B, cfg.A, cfg.S)      # F:246 Replicate
        proj = Fn.conv2d(I.reshape(B, cfg.M, cfg.S, 1), P["Wa"].view(cfg.A, cfg.M, 1, 1), P["ba"]).reshape(B, cfg.A, cfg.S)  # F:247-249
        add = torch.tanh(proj + qatt).reshape(B, cfg.A, cfg.S, 1)                          # F:250
        att = Fn.conv2d(add, P["ws"].view(1, cfg.A, 1, 1), P["bs"]).reshape(B, cfg.S)      # F:251
        p = torch.softmax(att + Fn.linear(h, P["Wm"], P["bm"]), dim=1)                     # F:287-289
        a = (I * p.unsqueeze(1).expand(B, cfg.M, cfg.S)).sum(2)                            # F:254-263
        j = (qf + a) + Fn.linear(p, P["Wp"], P["bp"])                                      # F:270-272
        c2, h2 = self.attlstm(j, c, h)                                                     # F:273
        m = self._drop(j + Fn.linear(h2, P["Wo"], P["bo"]), mm, cfg.p_m)                   # F:277-279
        score = Fn.linear(m, P["Ws"], P["bso"])                                            # F:280
        do_pred = torch.sigmoid(Fn.linear(m, P["wd"], P["bd"])).sum(1)                     # F:281
        return score, do_pred, p, c2, h2

    def forward(self, X, x, x_len, y, masks=None, hop_mask=None):
        """Returns (joint loss tensor, per-hop scores) following feval F:460-537, F:585-589."""
        cfg = self.cfg
        T, B = x.shape
        Xt = torch.as_tensor(X, dtype=torch.float64)
        max_len = int(np.max(x_len))
        state = torch.zeros(B, cfg.Q, dtype=torch.float64)
        rnn_out = torch.zeros(B, cfg.Q, dtype=torch.float64)
        lens = torch.as_tensor(np.asarray(x_len))
        for t in range(max_len):
            me = None if masks is None else masks["embed"][t]
            mr = None if masks is None else masks["rnn"][t]
            state = self.deeplstm(self.word_embed(x[t], me), state, mr)
            selm = (lens == t + 1).unsqueeze(1)
            rnn_out = torch.where(selm, state, rnn_out)          # rnn_out[k] = lst[k]  F:472-478
        c = torch.zeros(B, cfg.H, dtype=torch.float64)
        h = torch.zeros(B, cfg.H, dtype=torch.float64)
        yt = torch.as_tensor(np.asarray(y), dtype=torch.long) - 1
        total = torch.zeros((), dtype=torch.float64)
        scores = []
        hop_mask = [True] * cfg.nHop if hop_mask is None else hop_mask
        for hp in range(cfg.nHop):
            hm = None if masks is None else masks["hops"][hp]
            score, do_pred, p, c, h = self.hop(rnn_out, Xt, c, h, hm)
            scores.append(score)
            if hop_mask[hp]:
                total = total + Fn.cross_entropy(score, yt)      # CrossEntropyCriterion, mean (F:535)
            # do_pred BCE gradient is multiplied by 0 (F:583); attprob gradient is zeros (F:592)
        return total, scores, rnn_out

    def grads(self, X, x, x_len, y, masks=None, hop_mask=None):
        for g in O.GROUPS:
            self.flat[g].grad = None
        total, scores, rnn_out = self.forward(X, x, x_len, y, masks, hop_mask)
        total.backward()
        return ({g: (self.flat[g].grad.numpy().copy() if self.flat[g].grad is not None
                     else np.zeros(self.flat[g].shape)) for g in O.GROUPS},
                [s.detach().numpy() for s in scores], rnn_out.detach().numpy())
