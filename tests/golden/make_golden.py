"""Generates tests/golden/*.npz from the float64 oracle (oracle/rau_oracle.py).

The reference ships no golden vectors and Torch7 cannot run in this image, so these fixtures pin the
ORACLE (against drift) and give the CUDA path a fixed target; they are not outputs of the reference itself.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import rau_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

TOY = dict(V=50, embed=12, Hq=16, nlayer=2, C=24, S=20, M=16, A=8, H=16, N=30, nHop=2, T=5)


def pack_masks(masks):
    return dict(mask_embed=masks["embed"].astype(np.uint8), mask_rnn=masks["rnn"].astype(np.uint8),
                mask_q=np.stack([h["q"] for h in masks["hops"]]).astype(np.uint8),
                mask_x=np.stack([h["X"] for h in masks["hops"]]).astype(np.uint8),
                mask_m=np.stack([h["m"] for h in masks["hops"]]).astype(np.uint8))


def step_fixture(name, cfg_kw, B, seed, optim):
    cfg = O.RauConfig(**cfg_kw)
    params = O.init_params(cfg, seed=seed)
    p0 = {g: params[g].copy() for g in O.GROUPS}
    X, x, x_len, y = O.synth_batch(cfg, B, seed=seed + 1, min_len=2)
    masks = O.synth_masks(cfg, B, seed=seed + 2)
    rng = np.random.default_rng(seed + 3)
    std = O.noise_std(cfg, 0)
    noise = {g: rng.standard_normal(params[g].size) * std for g in O.GROUPS}
    raw = O.feval(cfg, params, X, x, x_len, y, masks=masks, clip=False)
    opt_state = {}
    res = O.train_step(cfg, params, opt_state, X, x, x_len, y, masks=masks, noise=noise, optim=optim)
    out = dict(cfg=np.array([cfg_kw[k] for k in sorted(cfg_kw)]), cfg_keys=np.array(sorted(cfg_kw)), B=B, seed=seed,
               optim=optim, X=X, x=x, x_len=x_len, y=y, loss=res.loss, loss_do_pred=res.loss_do_pred,
               answers=res.answers, scores=np.stack(res.scores), attprob=np.stack(res.attprob),
               do_pred=np.stack(res.do_pred), norms=np.array([res.norms[g] for g in O.GROUPS]))
    out.update(pack_masks(masks))
    for g in O.GROUPS:
        out[f"p0_{g}"] = p0[g]
        out[f"noise_{g}"] = noise[g]
        out[f"graw_{g}"] = raw.grads[g]          # feval gradients before noise/clip
        out[f"g_{g}"] = res.grads[g]             # after noise + clip
        out[f"p1_{g}"] = params[g]               # after the optimizer step
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: float(v) for k, v in zip(O.GROUPS, out["norms"])}, res.loss)


def cells_fixture():
    """One ATTLSTM step and one 2-layer DeepLSTM step with explicit upstream gradients (a1/a2)."""
    cfg = O.RauConfig(**TOY)
    rng = np.random.default_rng(11)
    B, H = 3, cfg.H
    Wx, Whh = rng.uniform(-.3, .3, (4 * H, cfg.M)), rng.uniform(-.3, .3, (4 * H, H))
    bx, bhh = rng.uniform(-.3, .3, 4 * H), rng.uniform(-.3, .3, 4 * H)
    x, c, h = rng.standard_normal((B, cfg.M)), rng.standard_normal((B, H)), rng.standard_normal((B, H))
    c2, h2, cache = O.attlstm_fwd(Wx, bx, Whh, bhh, x, c, h)
    dc2, dh2 = rng.standard_normal((B, H)), rng.standard_normal((B, H))
    g = dict(Wx=np.zeros_like(Wx), bx=np.zeros_like(bx), Whh=np.zeros_like(Whh), bhh=np.zeros_like(bhh))
    dx, dc, dh = O.attlstm_bwd(Wx, Whh, g, cache, dc2, dh2)
    out = dict(att_Wx=Wx, att_Whh=Whh, att_bx=bx, att_bhh=bhh, att_x=x, att_c=c, att_h=h, att_c2=c2, att_h2=h2,
               att_dc2=dc2, att_dh2=dh2, att_dx=dx, att_dc=dc, att_dh=dh, att_gWx=g["Wx"], att_gbx=g["bx"],
               att_gWhh=g["Whh"], att_gbhh=g["bhh"])
    Pr = {n: rng.uniform(-.3, .3, s) for n, s in O.rnn_param_shapes(cfg)}
    xe, s0 = rng.standard_normal((B, cfg.embed)), rng.standard_normal((B, cfg.Q))
    mk = (rng.random((B, cfg.Hq)) >= 0.5).astype(np.float64)
    s1, lcache = O.deeplstm_fwd(Pr, cfg, xe, s0, [mk])
    ds1 = rng.standard_normal((B, cfg.Q))
    gP = {n: np.zeros(s) for n, s in O.rnn_param_shapes(cfg)}
    dxe, ds0 = O.deeplstm_bwd(Pr, gP, cfg, lcache, [mk], ds1)
    out.update(deep_x=xe, deep_s0=s0, deep_mask=mk.astype(np.uint8), deep_s1=s1, deep_ds1=ds1, deep_dx=dxe, deep_ds0=ds0)
    for n in Pr:
        out[f"deep_P_{n}"] = Pr[n]
        out[f"deep_g_{n}"] = gP[n]
    np.savez_compressed(os.path.join(HERE, "cells_toy.npz"), **out)
    print("cells_toy.npz")


def digest(a, k=24):
    """a tensor's fingerprint in a fixture that must stay small: sum, l2 norm, max |.|, and k evenly strided entries"""
    a = np.asarray(a, dtype=np.float64).ravel()
    idx = np.linspace(0, a.size - 1, min(k, a.size)).astype(np.int64)
    return np.concatenate([[a.sum(), np.linalg.norm(a), np.abs(a).max()], a[idx]])


def hop_fixture(name, C, seed):
    """SURVEY 8c `rau_hop_{C512,C2048}`: ONE answering unit at the reference's dimensions (M 512, A 256, H 512, S 196; a
    64-way head and batch 2 keep it small), forward and backward with explicit upstream gradients.  Inputs and weights
    are regenerated from the seeds (float32-rounded, as the GPU sees them); outputs are stored in full, gradients as digests."""
    cfg = O.RauConfig(V=10, C=C, N=64, nHop=1)
    B = 2
    rng = np.random.default_rng(seed)
    f32 = lambda a: np.asarray(a, np.float32).astype(np.float64)
    pm = f32(rng.uniform(-0.08, 0.08, O.group_size(cfg, "mult")))
    Pm = O.views(cfg, "mult", pm)
    q, X = f32(rng.standard_normal((B, cfg.Q))), f32(np.maximum(rng.standard_normal((B, C, cfg.S)), 0))
    c, h = f32(rng.standard_normal((B, cfg.H)) * 0.5), f32(np.tanh(rng.standard_normal((B, cfg.H))))
    mk = dict(q=(rng.random((B, cfg.Q)) >= 0.5).astype(np.uint8), X=(rng.random((B, C, cfg.S)) >= 0.5).astype(np.uint8),
              m=(rng.random((B, cfg.M)) >= 0.5).astype(np.uint8))
    score, dop, p, c2, h2, cache = O.hop_fwd(Pm, cfg, q, X, c, h, mk)
    ups = [f32(rng.standard_normal(t.shape) * s) for t, s in ((score, 0.01), (dop, 0.01), (p, 0.01), (c2, 0.01), (h2, 0.01))]
    gflat = np.zeros(pm.size)
    dq, _, dc, dh = O.hop_bwd(Pm, O.views(cfg, "mult", gflat), cfg, cache, *ups)
    out = dict(C=C, B=B, seed=seed, N=cfg.N, score=score, do_pred=dop, p=p, c2=c2, h2=h2, dq=digest(dq, 64), dc=dc, dh=dh)
    for n_, v in O.views(cfg, "mult", gflat).items():
        out[f"g_{n_}"] = digest(v)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, float(np.abs(score).max()), float(np.linalg.norm(gflat)))


def hop_fixture_inputs(C, seed):
    """the inputs of hop_fixture again (tests regenerate them; the draw order is the generator's)"""
    cfg = O.RauConfig(V=10, C=C, N=64, nHop=1)
    B = 2
    rng = np.random.default_rng(seed)
    f32 = lambda a: np.asarray(a, np.float32).astype(np.float64)
    pm = f32(rng.uniform(-0.08, 0.08, O.group_size(cfg, "mult")))
    q, X = f32(rng.standard_normal((B, cfg.Q))), f32(np.maximum(rng.standard_normal((B, C, cfg.S)), 0))
    c, h = f32(rng.standard_normal((B, cfg.H)) * 0.5), f32(np.tanh(rng.standard_normal((B, cfg.H))))
    mk = dict(q=(rng.random((B, cfg.Q)) >= 0.5).astype(np.uint8), X=(rng.random((B, C, cfg.S)) >= 0.5).astype(np.uint8),
              m=(rng.random((B, cfg.M)) >= 0.5).astype(np.uint8))
    shapes = [(B, cfg.N), (B,), (B, cfg.S), (B, cfg.H), (B, cfg.H)]
    ups = [f32(rng.standard_normal(s) * 0.01) for s in shapes]
    return cfg, pm, q, X, c, h, mk, ups


def encoder_fixture():
    """SURVEY 8c `encoder_varlen`: the question encoder (F:460-479 / F:600-615) on ragged lengths incl. 1 and T, with an
    explicit gradient into rnn_out."""
    cfg = O.RauConfig(**dict(TOY, T=7))
    B = 6
    rng = np.random.default_rng(77)
    params = O.init_params(cfg, seed=78)
    x_len = np.array([1, 7, 3, 7, 2, 5])
    x = rng.integers(2, cfg.V + 1, (cfg.T, B))
    for b in range(B):
        x[x_len[b]:, b] = 1
    masks = dict(embed=(rng.random((cfg.T, B, cfg.embed)) >= 0.5).astype(np.float64),
                 rnn=(rng.random((cfg.T, B, cfg.Hq)) >= 0.5).astype(np.float64))
    Pe, Pr = O.views(cfg, "embed", params["embed"]), O.views(cfg, "rnn", params["rnn"])
    rnn_out, caches = O.encoder_fwd(Pe, Pr, cfg, x, x_len, masks)
    dq = rng.standard_normal(rnn_out.shape)
    gE, gR = np.zeros_like(params["embed"]), np.zeros_like(params["rnn"])
    O.encoder_bwd(Pr, O.views(cfg, "embed", gE), O.views(cfg, "rnn", gR), cfg, x_len, caches, dq)
    np.savez_compressed(os.path.join(HERE, "encoder_varlen.npz"), cfg=np.array([dict(TOY, T=7)[k] for k in sorted(TOY)]),
                        cfg_keys=np.array(sorted(TOY)), p_embed=params["embed"], p_rnn=params["rnn"], x=x, x_len=x_len,
                        mask_embed=masks["embed"].astype(np.uint8), mask_rnn=masks["rnn"].astype(np.uint8), rnn_out=rnn_out,
                        dq=dq, g_embed=gE, g_rnn=gR)
    print("encoder_varlen.npz", float(np.linalg.norm(gR)))


def noise_clip_fixture():
    """SURVEY 8c `noise_clip`: F:617-648 on three small groups -- one far above the clip threshold, one just above, one below."""
    cfg = O.RauConfig(**TOY)
    rng = np.random.default_rng(91)
    sizes = dict(embed=700, rnn=1300, mult=900)
    scale = dict(embed=1.0, rnn=0.1 / np.sqrt(1300) * 1.05, mult=1e-4)
    out = dict(step_t=np.array(4))
    for g in O.GROUPS:
        grad = rng.standard_normal(sizes[g]) * scale[g]
        noise = rng.standard_normal(sizes[g]) * O.noise_std(cfg, 4) * 1e-3
        clipped = grad.copy()
        n = O.noise_and_clip(cfg, clipped, noise)
        out.update({f"grad_{g}": grad, f"noise_{g}": noise, f"out_{g}": clipped, f"norm_{g}": np.array(n)})
    np.savez_compressed(os.path.join(HERE, "noise_clip.npz"), **out)
    print("noise_clip.npz", {g: float(out[f"norm_{g}"]) for g in O.GROUPS})


if __name__ == "__main__":
    step_fixture("step_toy_adam.npz", TOY, B=3, seed=123, optim="adam")
    step_fixture("step_toy_rmsprop.npz", dict(TOY, nHop=1), B=4, seed=321, optim="rmsprop")
    cells_fixture()
    # SURVEY.md 8c's list
    hop_fixture("rau_hop_C512.npz", 512, seed=512)
    hop_fixture("rau_hop_C2048.npz", 2048, seed=2048)
    for nh in (1, 3, 8):
        step_fixture(f"joint_step_nHop{nh}.npz", dict(TOY, nHop=nh), B=4, seed=1000 + nh, optim="adam")
    encoder_fixture()
    noise_clip_fixture()
