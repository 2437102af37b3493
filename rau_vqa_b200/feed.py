"""Host-side wrappers of the batch feed (include/rau.h, rau_feed_* / rau_feat_cache_*).

`Feed` is the pinned, double-buffered replacement of the reference's per-step `feats:float():cuda()` (F:452-456) behind
the loader's return tuple (feats, x, x_len, y; LD:1009); `FeatCache` keeps a split's features resident in HBM as fp16."""
from __future__ import annotations

import numpy as np
import torch

from ._ffi import check, ffi
from .core import Context, RauConfig, fptr

FEED_F32, FEED_F16, FEED_F16_DIRECT = 0, 1, 2   # (DIRECT: no widening pass; the training step reads the fp16 upload)


class Feed:
    def __init__(self, ctx: Context, cfg: RauConfig, B: int, fmt: int = FEED_F32, depth: int = 2):
        self.ctx, self.cfg, self.B, self.fmt, self.depth = ctx, cfg, B, fmt, depth
        out = ffi.new("rau_feed**")
        check(ctx.lib.rau_feed_create(ctx.h, cfg.c(), B, fmt, depth, out))
        self.h = out[0]
        self.bytes_per_batch = int(ctx.lib.rau_feed_host_bytes(self.h))

    def close(self):
        if getattr(self, "h", None) is not None:
            self.ctx.lib.rau_feed_destroy(self.h)
            self.h = None

    def host_slot(self, slot: int):
        """numpy views of the slot's pinned staging: (feats [B,C,S] float32|float16, tokens [T,B], lengths [B], labels [B])"""
        pf, pt, pl, py = ffi.new("void**"), ffi.new("float**"), ffi.new("float**"), ffi.new("float**")
        check(self.ctx.lib.rau_feed_host_slot(self.h, slot, pf, pt, pl, py))
        B, C, S, T = self.B, self.cfg.C, self.cfg.S, self.cfg.T
        fdt = np.float32 if self.fmt == FEED_F32 else np.float16
        feats = np.frombuffer(ffi.buffer(pf[0], B * C * S * np.dtype(fdt).itemsize), dtype=fdt).reshape(B, C, S)
        tok = np.frombuffer(ffi.buffer(pt[0], T * B * 4), dtype=np.float32).reshape(T, B)
        ln = np.frombuffer(ffi.buffer(pl[0], B * 4), dtype=np.float32)
        lab = np.frombuffer(ffi.buffer(py[0], B * 4), dtype=np.float32)
        return feats, tok, ln, lab

    def fill(self, slot: int, feats, tokens, lengths, labels):
        """the loader side: cast one batch (host arrays; feats float64 or float32 like LD:1009 delivers) into the staging"""
        f, t, ln, y = self.host_slot(slot)
        src = np.ascontiguousarray(feats)
        if src.dtype not in (np.float32, np.float64):
            src = src.astype(np.float32)
        check(self.ctx.lib.rau_feed_convert(self.h, ffi.cast("const void*", src.ctypes.data), int(src.dtype == np.float64),
                                            src.size, ffi.cast("void*", f.ctypes.data)))
        t[...] = tokens
        ln[...] = lengths
        y[...] = labels

    def submit(self, slot: int):
        check(self.ctx.lib.rau_feed_submit(self.h, slot))

    def acquire(self, slot: int, B_global: int = 0):
        """-> rau_batch* (cffi) whose device pointers stay valid until release(slot)"""
        b = ffi.new("rau_batch*")
        check(self.ctx.lib.rau_feed_acquire(self.h, slot, b))
        if B_global:
            b.B_global = B_global
        return b

    def release(self, slot: int):
        check(self.ctx.lib.rau_feed_release(self.h, slot))


class FeatCache:
    def __init__(self, ctx: Context, n_images: int, C: int, S: int = 196):
        self.ctx, self.n, self.C, self.S = ctx, n_images, C, S
        out = ffi.new("rau_feat_cache**")
        check(ctx.lib.rau_feat_cache_create(ctx.h, n_images, C, S, out))
        self.h = out[0]

    def close(self):
        if getattr(self, "h", None) is not None:
            self.ctx.lib.rau_feat_cache_destroy(self.h)
            self.h = None

    def put(self, first: int, feats):
        a = np.ascontiguousarray(feats, dtype=np.float32)
        check(self.ctx.lib.rau_feat_cache_put(self.h, first, a.shape[0], ffi.cast("const float*", a.ctypes.data)))

    def gather(self, image_index: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        B = image_index.numel()
        if out is None:
            out = torch.empty(B, self.C, self.S, dtype=torch.float32, device=image_index.device)
        check(self.ctx.lib.rau_feat_cache_gather(self.h, fptr(image_index), B, fptr(out)))
        return out

    def gather_f16(self, image_index: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """the batch's features as fp16 [B, C, S] for rau_batch.feats_f16 (core.train_step(..., feats_f16=...))"""
        B = image_index.numel()
        if out is None:
            out = torch.empty(B, self.C, self.S, dtype=torch.float16, device=image_index.device)
        check(self.ctx.lib.rau_feat_cache_gather_f16(self.h, fptr(image_index), B, ffi.cast("void*", out.data_ptr())))
        return out


def train_step_batch(ctx: Context, cfg: RauConfig, params, grads, opt_state, batch, out, optim=5, lrs=(3e-3, 3e-3, 3e-4),
                     hyper=(0.9, 0.999, 1e-8), eta=0.01, gamma=0.55, clip=0.1, step_t=0, opt_t=0):
    """rau_train_step on a rau_batch* handed out by Feed.acquire (same arguments as core.train_step otherwise)."""
    keep = []
    hp = ffi.new("rau_train_hparams*")
    hp.optim = optim
    for g in range(3):
        hp.lr[g] = lrs[g]
    hp.h0, hp.h1, hp.h2 = hyper
    hp.eta, hp.gamma, hp.clip = eta, gamma, clip
    hp.opt_t = int(opt_t)
    st = ffi.new("float*[3][2]")
    for g in range(3):
        for k in range(2):
            st[g][k] = fptr(opt_state[g][k]) if opt_state is not None and opt_state[g][k] is not None else ffi.NULL
    P = ffi.new("float*[3]", [fptr(t) for t in params])
    G = ffi.new("float*[3]", [fptr(t) for t in grads])
    check(ctx.lib.rau_train_step(ctx.h, cfg.c(), batch, P, G, st, ffi.NULL, ffi.NULL, step_t, hp, out.c(keep)))
