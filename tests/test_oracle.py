"""The oracle against an independent derivation (torch.autograd over torch.nn-style modules that mirror the
nngraph, oracle/torch_graph.py), finite differences, and hand-computed optimizer steps.  CPU only."""
import numpy as np
import pytest

from conftest import small_cfg
from oracle import rau_oracle as O
from oracle.torch_graph import TorchRAU


def _setup(cfg, B, seed=0, dropout=True):
    params = O.init_params(cfg, seed=seed)
    X, x, x_len, y = O.synth_batch(cfg, B, seed=seed + 1, min_len=2)
    masks = O.synth_masks(cfg, B, seed=seed + 2) if dropout else None
    return params, X, x, x_len, y, masks


@pytest.mark.parametrize("dropout", [False, True])
@pytest.mark.parametrize("nHop", [1, 3])
def test_feval_matches_autograd(dropout, nHop):
    cfg = small_cfg(nHop=nHop)
    params, X, x, x_len, y, masks = _setup(cfg, B=4, dropout=dropout)
    res = O.feval(cfg, params, X, x, x_len, y, masks=masks, clip=False)
    tg, tscores, trnn = TorchRAU(cfg, params).grads(X, x, x_len, y, masks=masks)
    np.testing.assert_allclose(res.rnn_out, trnn, rtol=1e-12, atol=1e-14)
    for h in range(nHop):
        np.testing.assert_allclose(res.scores[h], tscores[h], rtol=1e-10, atol=1e-12)
    for g in O.GROUPS:
        scale = np.abs(tg[g]).max()
        assert np.abs(res.grads[g] - tg[g]).max() <= 1e-10 * max(scale, 1e-30), g


def test_hop_mask_zeroes_that_hops_loss_gradient():
    cfg = small_cfg(nHop=3)
    params, X, x, x_len, y, masks = _setup(cfg, B=3)
    hm = [True, False, True]
    res = O.feval(cfg, params, X, x, x_len, y, masks=masks, hop_mask=hm, clip=False)
    tg, _, _ = TorchRAU(cfg, params).grads(X, x, x_len, y, masks=masks, hop_mask=hm)
    for g in O.GROUPS:
        np.testing.assert_allclose(res.grads[g], tg[g], rtol=1e-9, atol=1e-13)
    assert res.loss[1] > 0          # the loss itself is still reported (F:535)


def test_feval_finite_differences():
    cfg = small_cfg(nHop=2)
    params, X, x, x_len, y, masks = _setup(cfg, B=2)
    res = O.feval(cfg, params, X, x, x_len, y, masks=masks, clip=False)
    rng = np.random.default_rng(5)

    def joint(p):
        return O.feval(cfg, p, X, x, x_len, y, masks=masks, clip=False).loss[:cfg.nHop].sum()

    for g in O.GROUPS:
        used = np.flatnonzero(res.grads[g])
        for i in rng.choice(used, size=4, replace=False):
            eps = 1e-6
            pp = {k: v.copy() for k, v in params.items()}
            pm = {k: v.copy() for k, v in params.items()}
            pp[g][i] += eps
            pm[g][i] -= eps
            fd = (joint(pp) - joint(pm)) / (2 * eps)
            assert abs(fd - res.grads[g][i]) <= 1e-6 * max(1.0, abs(fd)), (g, i, fd, res.grads[g][i])


def test_lengths_select_the_right_state_and_pad_rows_get_no_gradient():
    cfg = small_cfg(nHop=1)
    params, X, x, x_len, y, _ = _setup(cfg, B=4, dropout=False)
    x_len[:] = [2, 5, 3, 5]
    for b in range(4):
        x[:, b] = 7
        x[x_len[b]:, b] = 1
    res = O.feval(cfg, params, X, x, x_len, y, clip=False)
    # a question cut after 2 tokens encodes like the same 2 tokens run alone
    x2 = x[:, :1].copy()
    rnn2, _ = O.encoder_fwd(O.views(cfg, "embed", params["embed"]), O.views(cfg, "rnn", params["rnn"]), cfg, x2,
                            np.array([2]), None)
    np.testing.assert_allclose(res.rnn_out[0], rnn2[0], rtol=1e-13)
    # the pad token's embedding row only receives gradient through steps <= len, which never see token 1 here
    gE = O.views(cfg, "embed", res.grads["embed"])["E"]
    assert np.all(gE[0] == 0.0) and np.any(gE[6] != 0.0)


def test_dX_branch_matches_autograd():
    import torch
    cfg = small_cfg(nHop=1)
    params, X, x, x_len, y, masks = _setup(cfg, B=2)
    Pm = O.views(cfg, "mult", params["mult"])
    rng = np.random.default_rng(1)
    q, c, h = rng.standard_normal((2, cfg.Q)), rng.standard_normal((2, cfg.H)), rng.standard_normal((2, cfg.H))
    score, dop, p, c2, h2, cache = O.hop_fwd(Pm, cfg, q, X, c, h, masks["hops"][0])
    ups = [rng.standard_normal(t.shape) for t in (score, dop, p, c2, h2)]
    gP = O.views(cfg, "mult", np.zeros_like(params["mult"]))
    dq, dX, dc, dh = O.hop_bwd(Pm, gP, cfg, cache, *ups, want_dX=True)
    T = TorchRAU(cfg, params)
    tq, tX, tc, th = (torch.tensor(a, requires_grad=True) for a in (q, X, c, h))
    outs = T.hop(tq, tX, tc, th, masks["hops"][0])
    sum((o * torch.tensor(u)).sum() for o, u in zip(outs, ups)).backward()
    for a, b in ((dq, tq.grad), (dX, tX.grad), (dc, tc.grad), (dh, th.grad)):
        np.testing.assert_allclose(a, b.numpy().reshape(a.shape), rtol=1e-9, atol=1e-12)
    o = cfg_off(cfg, "wd")     # the do_pred head does get gradient when the caller supplies one
    np.testing.assert_allclose(gP["wd"], T.flat["mult"].grad.numpy()[o:o + cfg.M][None, :], rtol=1e-9, atol=1e-12)


def cfg_off(cfg, name):
    off = 0
    for n, shp in O.mult_param_shapes(cfg):
        if n == name:
            return off
        off += int(np.prod(shp))
    raise KeyError(name)


def test_adam_and_rmsprop_follow_optim_updates_lua():
    rng = np.random.default_rng(0)
    x0, g = rng.standard_normal(7), rng.standard_normal(7)
    x, st = x0.copy(), {}
    O.adam(x, g, 1e-2, st)
    # first step: m = .1 g, v = .001 g^2, step = lr*sqrt(1-.999)/(1-.9)   (OU:76-86)
    exp = x0 - 1e-2 * np.sqrt(1 - 0.999) / (1 - 0.9) * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-8)
    np.testing.assert_allclose(x, exp, rtol=1e-14)
    assert st["t"] == 1
    x, st = x0.copy(), {}
    O.rmsprop(x, g, 1e-2, 0.99, 1e-8, st)
    np.testing.assert_allclose(x, x0 - 1e-2 * g / (np.sqrt(0.01 * g * g) + 1e-8), rtol=1e-14)   # OU:52-56


def test_noise_formula_is_a_product_not_a_power():
    cfg = small_cfg()
    assert O.noise_std(cfg, 9) == pytest.approx(np.sqrt(0.01 / (10 * 0.55)))   # F:617-618


def test_clip_per_group():
    cfg = small_cfg()
    g = np.full(100, 0.05)
    n = O.noise_and_clip(cfg, g, None)
    assert n == pytest.approx(0.5) and np.linalg.norm(g) == pytest.approx(0.1)   # F:628-630
    g = np.full(100, 0.001)
    O.noise_and_clip(cfg, g, None)
    assert np.allclose(g, 0.001)


def test_predict_merges_like_predict_result():
    cfg = small_cfg(nHop=3)
    params, X, x, x_len, y, _ = _setup(cfg, B=5, dropout=False)
    preds, atts = O.predict(cfg, params, X, x, x_len)
    assert len(preds) == cfg.nHop + 2
    np.testing.assert_allclose(preds[cfg.nHop], sum(preds[:cfg.nHop]) / cfg.nHop, rtol=1e-13)   # F:718
    # select = exactly one hop's score per row (the last hop is forced, F:704)
    for b in range(5):
        assert any(np.array_equal(preds[cfg.nHop + 1][b], preds[h][b]) for h in range(cfg.nHop))
    np.testing.assert_allclose(atts[cfg.nHop].sum(1), 1.0, rtol=1e-12)
