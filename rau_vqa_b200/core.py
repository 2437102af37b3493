"""Thin host-side wrappers over the C ABI: configuration struct, context and the fused step calls.

torch is used for device memory and streams only (caller-owned tensors whose raw pointers cross the ABI); no torch
operator is on the compute path.  Every function raises RauError on a non-zero status -- there is no fallback.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass

import torch

from ._ffi import RauError, check, ffi, load

GATES_IFOG, GATES_IGFO = 0, 1
PREC_F32, PREC_BF16, PREC_BF16X3, PREC_MIXED, PREC_F16IMG = 0, 1, 2, 3, 4
OPT_SGD, OPT_SGDM, OPT_SGDMOM, OPT_ADAGRAD, OPT_RMSPROP, OPT_ADAM = range(6)
GROUPS = ("embed", "rnn", "mult")


@dataclass
class RauConfig:
    """The constants of the experiment scripts (F:202-229); defaults are Ours_Full."""
    V: int = 16384
    embed: int = 200
    Hq: int = 512
    nlayer: int = 2
    C: int = 512
    S: int = 196
    M: int = 512
    A: int = 256
    H: int = 512
    N: int = 2000
    nHop: int = 8
    T: int = 26
    p_embed: float = 0.5
    p_rnn: float = 0.5
    p_q: float = 0.5
    p_x: float = 0.5
    p_m: float = 0.5

    @property
    def Q(self) -> int:
        return 2 * self.Hq * self.nlayer

    def c(self):
        s = ffi.new("rau_config*")
        for f in dataclasses.fields(self):
            setattr(s, f.name, getattr(self, f.name))
        return s

    def group_size(self, group) -> int:
        g = GROUPS.index(group) if isinstance(group, str) else group
        return int(load().rau_group_size(self.c(), g))

    def param_offset(self, group, name: str) -> int:
        g = GROUPS.index(group) if isinstance(group, str) else group
        return int(load().rau_param_offset(self.c(), g, name.encode()))


def fptr(t):
    """float* of a contiguous float32 CUDA tensor (None -> NULL)."""
    if t is None:
        return ffi.NULL
    if not (isinstance(t, torch.Tensor) and t.dtype == torch.float32 and t.is_contiguous()):
        raise TypeError("expected a contiguous float32 tensor")
    return ffi.cast("float*", t.data_ptr())


def bptr(t):
    """uint8_t* of a contiguous uint8 tensor (None -> NULL)."""
    if t is None:
        return ffi.NULL
    if not (isinstance(t, torch.Tensor) and t.dtype == torch.uint8 and t.is_contiguous()):
        raise TypeError("expected a contiguous uint8 tensor")
    return ffi.cast("uint8_t*", t.data_ptr())


class Context:
    """rau_ctx: one per (device, stream).  Work is enqueued on the stream; call sync() before reading results
    on the host (torch's own synchronisation also covers it when the ctx uses torch's current stream)."""

    def __init__(self, device: int = 0, stream=None, precision: int | None = None, seed: int | None = None):
        self.lib = load()
        out = ffi.new("rau_ctx**")
        if stream is None:
            stream = torch.cuda.current_stream(device).cuda_stream if torch.cuda.is_available() else 0
        check(self.lib.rau_ctx_create(out, device, ffi.cast("void*", stream)))
        self.h = out[0]
        self.device = device
        if precision is not None:
            self.set_precision(precision)
        if seed is not None:
            self.set_seed(seed)

    def close(self):
        if getattr(self, "h", None) is not None:
            self.lib.rau_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_precision(self, p: int):
        check(self.lib.rau_set_precision(self.h, p))

    def set_seed(self, s: int):
        check(self.lib.rau_set_seed(self.h, s))

    def sync(self):
        check(self.lib.rau_sync(self.h))

    @property
    def launches(self) -> int:
        return int(self.lib.rau_launch_count(self.h))

    # ---- data parallel
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        check(self.lib.rau_comm_init(self.h, ffi.from_buffer("uint8_t[]", bytearray(unique_id)), rank, world))

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = ffi.new("uint8_t[128]")
        check(load().rau_comm_unique_id(buf))
        return bytes(ffi.buffer(buf, 128))


def _batch_struct(keep, B, feats, tokens, lengths, labels, max_len, B_global):
    b = ffi.new("rau_batch*")
    b.B = B
    b.B_global = B_global or B
    if feats.dtype == torch.float16:     # fp16 features go in as they are (rau_batch.feats_f16; the training step only)
        b.feats = ffi.NULL
        b.feats_f16 = ffi.cast("const void*", feats.data_ptr())
    else:
        b.feats = fptr(feats)
    b.tokens = fptr(tokens)
    b.lengths = fptr(lengths)
    b.max_len = int(max_len)
    b.labels = fptr(labels)
    keep.append(b)
    return b


def _masks_struct(keep, masks):
    if masks is None:
        return ffi.NULL
    m = ffi.new("rau_masks*")
    for k in ("embed", "rnn", "q", "x", "m"):
        setattr(m, k, bptr(masks.get(k)))
    keep.append(m)
    return m


class StepBuffers:
    """Device outputs of one feval (tab_loss, answers, ...), allocated once per (cfg, B)."""

    def __init__(self, cfg: RauConfig, B: int, device, want_scores=True):
        f = dict(dtype=torch.float32, device=device)
        self.loss = torch.zeros(cfg.nHop + 2, **f)
        self.loss_do_pred = torch.zeros(cfg.nHop, **f)
        self.answers = torch.zeros(cfg.nHop + 2, B, **f)
        self.scores = torch.zeros(cfg.nHop, B, cfg.N, **f) if want_scores else None
        self.attprob = torch.zeros(cfg.nHop, B, cfg.S, **f) if want_scores else None
        self.do_pred = torch.zeros(cfg.nHop, B, **f) if want_scores else None
        self.norms = torch.zeros(3, **f)

    def c(self, keep):
        o = ffi.new("rau_step_out*")
        o.loss = fptr(self.loss)
        o.loss_do_pred = fptr(self.loss_do_pred)
        o.answers = fptr(self.answers)
        o.scores = fptr(self.scores)
        o.attprob = fptr(self.attprob)
        o.do_pred = fptr(self.do_pred)
        o.norms = fptr(self.norms)
        keep.append(o)
        return o


def _triple(keep, tensors):
    arr = ffi.new("float*[3]", [fptr(t) for t in tensors])
    keep.append(arr)
    return arr


def feval(ctx: Context, cfg: RauConfig, params, grads, feats, tokens, lengths, labels, out: StepBuffers,
          hop_mask=None, masks=None, step_t: int = 0, max_len: int = 0, B_global: int = 0):
    """feval (F:445-615) on flat parameter/gradient triples (embed, rnn, mult).  Gradients are left before
    noise/clip.  tokens [T,B], lengths [B], labels [B] are float tensors of 1-based ids like the reference."""
    keep = []
    B = feats.shape[0]
    b = _batch_struct(keep, B, feats, tokens, lengths, labels, max_len, B_global)
    hm = ffi.new("float[]", [float(v) for v in hop_mask]) if hop_mask is not None else ffi.NULL
    check(ctx.lib.rau_feval(ctx.h, cfg.c(), b, _triple(keep, params), _triple(keep, grads), hm,
                            _masks_struct(keep, masks), step_t, out.c(keep)))


def noise_clip(ctx, cfg, grads, step_t, eta, gamma, clip, noise=None, norms=None):
    keep = []
    nz = ffi.new("float*[3]", [fptr(t) for t in noise]) if noise is not None else ffi.NULL
    check(ctx.lib.rau_noise_clip(ctx.h, cfg.c(), _triple(keep, grads), step_t, eta, gamma, clip,
                                 ffi.cast("const float* const*", nz) if noise is not None else ffi.NULL, fptr(norms)))


def optim_step(ctx, optim, x, dx, lr, h0=0.0, h1=0.0, h2=0.0, state0=None, state1=None, t=1):
    check(ctx.lib.rau_optim_step(ctx.h, optim, x.numel(), fptr(x), fptr(dx), lr, h0, h1, h2, fptr(state0), fptr(state1), t))


def train_step(ctx: Context, cfg: RauConfig, params, grads, opt_state, feats, tokens, lengths, labels, out: StepBuffers,
               optim=OPT_ADAM, lrs=(3e-3, 3e-3, 3e-4), hyper=(0.9, 0.999, 1e-8), eta=0.01, gamma=0.55, clip=0.1,
               hop_mask=None, masks=None, noise=None, step_t: int = 0, max_len: int = 0, B_global: int = 0, opt_t: int = 0):
    """feval + [all-reduce] + noise/clip + the three optimizer calls (F:786-791) as one enqueue.
    step_t = the reference's `it` (keys dropout / noise, noise variance eta/((it+1) gamma), F:617); opt_t = adam's own step
    count after its increment (OU:79; 0 = step_t + 1)."""
    keep = []
    B = feats.shape[0]
    b = _batch_struct(keep, B, feats, tokens, lengths, labels, max_len, B_global)
    hm = ffi.new("float[]", [float(v) for v in hop_mask]) if hop_mask is not None else ffi.NULL
    hp = ffi.new("rau_train_hparams*")
    hp.optim = optim
    for g in range(3):
        hp.lr[g] = lrs[g]
    hp.h0, hp.h1, hp.h2 = hyper
    hp.eta, hp.gamma, hp.clip = eta, gamma, clip
    hp.opt_t = int(opt_t)
    if noise is not None:
        nz = ffi.new("float*[3]", [fptr(t) for t in noise])
        keep.append(nz)
        hp.noise_override = ffi.cast("const float* const*", nz)
    st = ffi.new("float*[3][2]")
    for g in range(3):
        for k in range(2):
            st[g][k] = fptr(opt_state[g][k]) if opt_state is not None and opt_state[g][k] is not None else ffi.NULL
    check(ctx.lib.rau_train_step(ctx.h, cfg.c(), b, _triple(keep, params), _triple(keep, grads), st, hm,
                                 _masks_struct(keep, masks), step_t, hp, out.c(keep)))


def draw_masks(ctx: Context, cfg: RauConfig, B: int, step_t: int, device=None):
    """The keep masks the step with iteration number step_t draws from this context's Philox streams, as uint8 tensors
    in the rau_masks layouts (dict embed / rnn / q / x / m).  Test hook (rau_draw_masks)."""
    dev = device if device is not None else torch.device("cuda", ctx.device)
    u8 = dict(dtype=torch.uint8, device=dev)
    out = dict(embed=torch.empty(cfg.T, B, cfg.embed, **u8), rnn=torch.empty(cfg.T, B, cfg.Hq, **u8),
               q=torch.empty(cfg.nHop, B, cfg.Q, **u8), x=torch.empty(cfg.nHop, B, cfg.C, cfg.S, **u8),
               m=torch.empty(cfg.nHop, B, cfg.M, **u8))
    mo = ffi.new("rau_masks_out*")
    for k, t in out.items():
        setattr(mo, k, bptr(t))
    check(ctx.lib.rau_draw_masks(ctx.h, cfg.c(), B, step_t, mo))
    return out


def predict(ctx: Context, cfg: RauConfig, params, feats, tokens, lengths, max_len: int = 0):
    """predict_result (F:652-724): returns (pred[nHop+2,B,N], att[nHop+2,B,S])."""
    keep = []
    B = feats.shape[0]
    b = _batch_struct(keep, B, feats, tokens, lengths, None, max_len, 0)
    pred = torch.empty(cfg.nHop + 2, B, cfg.N, dtype=torch.float32, device=feats.device)
    att = torch.empty(cfg.nHop + 2, B, cfg.S, dtype=torch.float32, device=feats.device)
    check(ctx.lib.rau_predict(ctx.h, cfg.c(), b, _triple(keep, params), fptr(pred), fptr(att)))
    return pred, att


def predict_answers(ctx: Context, cfg: RauConfig, params, feats, tokens, lengths, mc_choices=None, max_len: int = 0,
                    want_tables: bool = False):
    """predict_result + the test loop's answer extraction (F:903-918) on the device: returns (oe_answers[nHop+2,B],
    mc_answers[nHop+2,B] or None[, pred, att]); mc_choices = ans_mc [B, nmc] float, 1-based ids, 0 = empty."""
    keep = []
    B = feats.shape[0]
    b = _batch_struct(keep, B, feats, tokens, lengths, None, max_len, 0)
    f = dict(dtype=torch.float32, device=feats.device)
    oe = torch.empty(cfg.nHop + 2, B, **f)
    mc = torch.empty(cfg.nHop + 2, B, **f) if mc_choices is not None else None
    pred = torch.empty(cfg.nHop + 2, B, cfg.N, **f) if want_tables else None
    att = torch.empty(cfg.nHop + 2, B, cfg.S, **f) if want_tables else None
    check(ctx.lib.rau_predict_answers(ctx.h, cfg.c(), b, _triple(keep, params), fptr(mc_choices),
                                      0 if mc_choices is None else int(mc_choices.shape[1]), fptr(oe), fptr(mc), fptr(pred),
                                      fptr(att)))
    return (oe, mc, pred, att) if want_tables else (oe, mc)


__all__ = ["predict_answers", "RauConfig", "Context", "StepBuffers", "feval", "noise_clip", "optim_step", "train_step", "predict", "draw_masks",
           "RauError", "fptr", "bptr", "GROUPS"]
