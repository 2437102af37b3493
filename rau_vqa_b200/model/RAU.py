"""The inline model graph of the experiment scripts as two modules with the nn.Module surface:

* ``WordEmbed``  = protos.word_embed, LookupTable -> Dropout(0.5) -> Tanh (F:203-206)
* ``Multimodal`` = protos.multimodal (F:224-307): {q, feats, c, h} -> {score, do_pred, attprob, c', h'},
  one recurrent answering unit, ATTLSTM (model/ATTLSTM.lua) inside.  One native call per direction
  (rau_hop_fwd / rau_hop_bwd).

Parameters are held by child modules in the flat order of include/rau.h so that ``getParameters()`` yields the
``mult`` group layout the fused step (core.feval / core.train_step) consumes.
"""
from __future__ import annotations

import itertools

import torch

from .. import core, nn
from .._ffi import check, ffi

_stream_counter = itertools.count(1)


def _p(t):
    return core.fptr(t) if t is not None else ffi.NULL


class WordEmbed(nn.Module):
    def __init__(self, cfg: core.RauConfig, device="cpu"):
        super().__init__()
        self.cfg = cfg
        self.weight = torch.zeros(cfg.V, cfg.embed, dtype=torch.float32, device=device)   # nn.LookupTable.weight
        self.gradWeight = torch.zeros_like(self.weight)
        self.noise_override = None
        self._stream = 0

    def updateOutput(self, input):
        ids = input.contiguous()
        ctx = nn.context(ids.device)
        self._stream = next(_stream_counter)
        n = ids.numel()
        out = torch.empty(n, self.cfg.embed, dtype=torch.float32, device=ids.device)
        check(ctx.lib.rau_embed_fwd(ctx.h, self.cfg.c(), n, _p(ids), _p(self.weight), int(self.train),
                                    core.bptr(self.noise_override), self._stream, _p(out)))
        self.output = out
        return out

    def updateGradInput(self, input, gradOutput):
        self.gradInput = None          # LookupTable has no gradInput
        return None

    def accGradParameters(self, input, gradOutput, scale=1.0):
        ids = input.contiguous()
        ctx = nn.context(ids.device)
        g = gradOutput.contiguous() if scale == 1.0 else (gradOutput * scale).contiguous()
        check(ctx.lib.rau_embed_bwd(ctx.h, self.cfg.c(), ids.numel(), _p(ids), _p(self.output), int(self.train),
                                    core.bptr(self.noise_override), self._stream, _p(g), _p(self.gradWeight)))


# flat layout of the mult group (include/rau.h): the two scalar biases bs, bd come last so that every tensor above them
# starts on a 16-byte boundary; ws and wd are therefore bias-less Linear holders followed by two nn.Add-style holders
_MULT_LAYERS = (("Wq", "bq"), ("Wh", "bh"), ("Wi", "bi"), ("Wqa", "bqa"), ("Wa", "ba"), ("ws", None), ("Wm", "bm"),
                ("Wp", "bp"), ("Wx", "bx"), ("Whh", "bhh"), ("Wo", "bo"), ("Ws", "bso"), ("wd", None))
_MULT_TAIL_BIASES = ("bs", "bd")


def mult_shapes(cfg: core.RauConfig):
    return dict(Wq=(cfg.M, cfg.Q), Wh=(cfg.M, cfg.H), Wi=(cfg.M, cfg.C), Wqa=(cfg.A, cfg.M), Wa=(cfg.A, cfg.M),
                ws=(1, cfg.A), Wm=(cfg.S, cfg.H), Wp=(cfg.M, cfg.S), Wx=(4 * cfg.H, cfg.M), Whh=(4 * cfg.H, cfg.H),
                Wo=(cfg.M, cfg.H), Ws=(cfg.N, cfg.M), wd=(1, cfg.M))


class Multimodal(nn.Container):
    def __init__(self, cfg: core.RauConfig, device="cpu"):
        super().__init__()
        self.cfg = cfg
        shp = mult_shapes(cfg)
        for w, b in _MULT_LAYERS:
            self.modules.append(nn.Linear(shp[w][1], shp[w][0], device=device, bias=b is not None))
        for _ in _MULT_TAIL_BIASES:
            self.modules.append(nn.Add(1, device=device))
        self._flat = self._gflat = None
        self.masks = None              # optional dict(q=, x=, m=) of uint8 keep masks (parity tests)
        self._stream = 0
        self._saved = None

    def getParameters(self):
        self._flat, self._gflat = super().getParameters()
        return self._flat, self._gflat

    def _flats(self):
        """The flat mult-group buffers the native call reads; valid when the fields still view one storage in
        layout order (after getParameters() / share()), rebuilt otherwise."""
        slots = self._param_slots()
        first = getattr(slots[0][0], slots[0][1])
        base, gbase = first.data_ptr(), getattr(slots[0][0], slots[0][2]).data_ptr()
        off = 0
        ok = True
        for o, w, g in slots:
            t, gt = getattr(o, w), getattr(o, g)
            ok &= t.data_ptr() == base + 4 * off and gt.data_ptr() == gbase + 4 * off
            off += t.numel()
        if not ok:
            raise RuntimeError("Multimodal parameters are not one flat storage: call getParameters() first (F:324)")
        n = off
        st, gst = first.untyped_storage(), getattr(slots[0][0], slots[0][2]).untyped_storage()
        flat = torch.empty(0, dtype=torch.float32, device=first.device).set_(st, first.storage_offset(), (n,))
        gflat = torch.empty(0, dtype=torch.float32, device=first.device).set_(gst, getattr(slots[0][0], slots[0][2]).storage_offset(), (n,))
        return flat, gflat

    def updateOutput(self, input):
        q, X, c, h = (t.contiguous() for t in input)
        cfg, B = self.cfg, q.shape[0]
        ctx = nn.context(q.device)
        flat, _ = self._flats()
        self._stream = next(_stream_counter)
        f = dict(dtype=torch.float32, device=q.device)
        score, dop, p = torch.empty(B, cfg.N, **f), torch.empty(B, **f), torch.empty(B, cfg.S, **f)
        c2, h2 = torch.empty(B, cfg.H, **f), torch.empty(B, cfg.H, **f)
        nbytes = int(ctx.lib.rau_hop_saved_bytes(cfg.c(), B))
        self._saved = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
        mk = self.masks or {}
        check(ctx.lib.rau_hop_fwd(ctx.h, cfg.c(), B, _p(flat), _p(q), _p(X), _p(c), _p(h), int(self.train),
                                  core.bptr(mk.get("q")), core.bptr(mk.get("x")), core.bptr(mk.get("m")), self._stream,
                                  _p(score), _p(dop), _p(p), _p(c2), _p(h2), ffi.cast("void*", self._saved.data_ptr())))
        self.output = [score, dop, p, c2, h2]
        return self.output

    def backward(self, input, gradOutput, scale=1.0, want_dX=False):
        assert scale == 1.0
        q, X, c, h = (t.contiguous() for t in input)
        cfg, B = self.cfg, q.shape[0]
        ctx = nn.context(q.device)
        flat, gflat = self._flats()
        f = dict(dtype=torch.float32, device=q.device)
        dscore, ddop, dp, dc2, dh2 = (None if g is None else g.contiguous() for g in gradOutput)
        dq, dc, dh = torch.empty(B, cfg.Q, **f), torch.empty(B, cfg.H, **f), torch.empty(B, cfg.H, **f)
        dX = torch.empty(B, cfg.C, cfg.S, **f) if want_dX else None
        check(ctx.lib.rau_hop_bwd(ctx.h, cfg.c(), B, _p(flat), _p(gflat), _p(q), _p(X), _p(c), _p(h), int(self.train),
                                  ffi.cast("void*", self._saved.data_ptr()), _p(dscore), _p(ddop), _p(dp), _p(dc2), _p(dh2),
                                  _p(dq), _p(dX), _p(dc), _p(dh)))
        self.gradInput = [dq, dX, dc, dh]
        return self.gradInput

    def updateGradInput(self, input, gradOutput):
        raise NotImplementedError("Multimodal fuses updateGradInput and accGradParameters: call backward() (F:590)")
