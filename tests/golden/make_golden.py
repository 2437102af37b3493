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


if __name__ == "__main__":
    step_fixture("step_toy_adam.npz", TOY, B=3, seed=123, optim="adam")
    step_fixture("step_toy_rmsprop.npz", dict(TOY, nHop=1), B=4, seed=321, optim="rmsprop")
    cells_fixture()
