"""Shared glue for the parity tests: oracle <-> librau conversions."""
import os

import numpy as np

from oracle import rau_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    cfg = O.RauConfig(**{str(k): int(v) for k, v in zip(z["cfg_keys"], z["cfg"])})
    return z, cfg


def oracle_masks_from_golden(z, cfg):
    return dict(embed=z["mask_embed"].astype(np.float64), rnn=z["mask_rnn"].astype(np.float64),
                hops=[dict(q=z["mask_q"][h].astype(np.float64), X=z["mask_x"][h].astype(np.float64),
                           m=z["mask_m"][h].astype(np.float64)) for h in range(cfg.nHop)])


def lib_cfg(cfg):
    from rau_vqa_b200 import RauConfig
    return RauConfig(V=cfg.V, embed=cfg.embed, Hq=cfg.Hq, nlayer=cfg.nlayer, C=cfg.C, S=cfg.S, M=cfg.M, A=cfg.A, H=cfg.H,
                     N=cfg.N, nHop=cfg.nHop, T=cfg.T, p_embed=cfg.p_embed, p_rnn=cfg.p_rnn, p_q=cfg.p_q, p_x=cfg.p_x,
                     p_m=cfg.p_m)


def dev(a, dtype=None):
    import torch
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is None:
        dtype = torch.uint8 if t.dtype == torch.uint8 else torch.float32
    return t.to(dtype).cuda().contiguous()


def lib_masks(masks):
    """oracle mask dict -> dict of uint8 CUDA tensors in the rau_masks layout."""
    if masks is None:
        return None
    return dict(embed=dev(masks["embed"].astype(np.uint8)), rnn=dev(masks["rnn"].astype(np.uint8)),
                q=dev(np.stack([h["q"] for h in masks["hops"]]).astype(np.uint8)),
                x=dev(np.stack([h["X"] for h in masks["hops"]]).astype(np.uint8)),
                m=dev(np.stack([h["m"] for h in masks["hops"]]).astype(np.uint8)))


def rel_err(a, b):
    """max |a-b| / max |b|: the 'relative' of north_star's 1e-3 (a tensor-level relative error)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


# gradients that are ZERO by construction in the reference's step: the do_pred head receives a zeroed gradient
# (F:582-583 -> wd, bd) and the scalar bias of the 256 -> 1 attention convolution shifts every logit of the softmax alike
# (F:251, F:289 -> bs = sum_s ds_s = 0).  Their float64 value is rounding noise, so "relative to max|b|" is meaningless;
# they are held to tol * (largest gradient of the group) instead.
STRUCTURAL_ZERO = {"mult": ("wd", "bd", "bs")}


def per_tensor_rel_err(cfg, group, got_flat, ref_flat):
    """north_star's "gradients within 1e-3 relative", per NAMED tensor: {name: max|a-b| / max|b|} over the views of one
    flat group (O.views = what getParameters() hands the script, F:322-324)."""
    got = O.views(cfg, group, np.asarray(got_flat, dtype=np.float64))
    ref = O.views(cfg, group, np.asarray(ref_flat, dtype=np.float64))
    gmax = max(float(np.abs(ref_flat).max()), 1e-30)
    out = {}
    for name, r in ref.items():
        denom = float(np.abs(r).max())
        if name in STRUCTURAL_ZERO.get(group, ()):
            denom = gmax
        out[name] = float(np.abs(got[name] - r).max() / max(denom, 1e-30))
    return out


def assert_grads_per_tensor(cfg, grads, ref_grads, tol, report=None):
    """every named tensor of the three groups within tol; returns the worst (error, group.name)."""
    worst = (0.0, "")
    for g in O.GROUPS:
        errs = per_tensor_rel_err(cfg, g, grads[g], ref_grads[g])
        for name, e in errs.items():
            if report is not None:
                report[f"{g}.{name}"] = e
            worst = max(worst, (e, f"{g}.{name}"))
    bad = {k: v for k, v in (report or {}).items() if isinstance(v, float) and v > tol} if report is not None else None
    assert worst[0] <= tol, (worst, bad)
    return worst


def run_lib_feval(ctx, cfg, params, X, x, x_len, y, masks=None, hop_mask=None, B_global=0, step_t=0):
    """One rau_feval on the GPU from numpy inputs; returns (numpy grads dict, StepBuffers)."""
    import rau_vqa_b200 as R
    lc = lib_cfg(cfg)
    P = [dev(params[g]) for g in O.GROUPS]
    G = [t.clone().zero_() for t in P]
    out = R.StepBuffers(lc, X.shape[0], P[0].device)
    R.feval(ctx, lc, P, G, dev(X), dev(x), dev(x_len), dev(y), out, hop_mask=hop_mask, masks=lib_masks(masks),
            step_t=step_t, max_len=int(np.max(x_len)), B_global=B_global)
    ctx.sync()
    return {g: G[i].cpu().numpy().astype(np.float64) for i, g in enumerate(O.GROUPS)}, out
