"""The oracle reproduces the committed golden fixtures (pins the oracle against drift).  CPU only."""
import numpy as np
import pytest

from helpers import load_golden, oracle_masks_from_golden
from oracle import rau_oracle as O


@pytest.mark.parametrize("name", ["step_toy_adam.npz", "step_toy_rmsprop.npz"])
def test_step_fixture(name):
    z, cfg = load_golden(name)
    params = {g: z[f"p0_{g}"].copy() for g in O.GROUPS}
    masks = oracle_masks_from_golden(z, cfg)
    noise = {g: z[f"noise_{g}"] for g in O.GROUPS}
    raw = O.feval(cfg, params, z["X"], z["x"], z["x_len"], z["y"], masks=masks, clip=False)
    for g in O.GROUPS:
        np.testing.assert_allclose(raw.grads[g], z[f"graw_{g}"], rtol=1e-12, atol=1e-15)
    res = O.train_step(cfg, params, {}, z["X"], z["x"], z["x_len"], z["y"], masks=masks, noise=noise, optim=str(z["optim"]))
    np.testing.assert_allclose(res.loss, z["loss"], rtol=1e-12)
    np.testing.assert_array_equal(res.answers, z["answers"])
    for g in O.GROUPS:
        np.testing.assert_allclose(res.grads[g], z[f"g_{g}"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(params[g], z[f"p1_{g}"], rtol=1e-12, atol=1e-15)


def test_cells_fixture():
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "cells_toy.npz"))
    c2, h2, cache = O.attlstm_fwd(z["att_Wx"], z["att_bx"], z["att_Whh"], z["att_bhh"], z["att_x"], z["att_c"], z["att_h"])
    np.testing.assert_allclose(c2, z["att_c2"], rtol=1e-13)
    np.testing.assert_allclose(h2, z["att_h2"], rtol=1e-13)
    # gate order (i, g, f, o): c' = sigmoid(G3) c + sigmoid(G1) tanh(G2)   (A:12-24)
    G = z["att_x"] @ z["att_Wx"].T + z["att_bx"] + z["att_h"] @ z["att_Whh"].T + z["att_bhh"]
    H = c2.shape[1]
    sig = lambda v: 1 / (1 + np.exp(-v))
    np.testing.assert_allclose(c2, sig(G[:, 2 * H:3 * H]) * z["att_c"] + sig(G[:, :H]) * np.tanh(G[:, H:2 * H]), rtol=1e-12)
    np.testing.assert_allclose(h2, sig(G[:, 3 * H:]) * np.tanh(c2), rtol=1e-12)


def _digest(a, k=24):
    a = np.asarray(a, dtype=np.float64).ravel()
    idx = np.linspace(0, a.size - 1, min(k, a.size)).astype(np.int64)
    return np.concatenate([[a.sum(), np.linalg.norm(a), np.abs(a).max()], a[idx]])


@pytest.mark.parametrize("name", ["joint_step_nHop1.npz", "joint_step_nHop3.npz", "joint_step_nHop8.npz"])
def test_joint_step_fixtures(name):
    test_step_fixture(name)


@pytest.mark.parametrize("C", [512, 2048])
def test_hop_fixture_at_reference_dims(C):
    """SURVEY 8c rau_hop_{C512,C2048}: one answering unit at M 512 / A 256 / H 512 / S 196, inputs regenerated from the seed."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(__import__("helpers").GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    z = np.load(os.path.join(__import__("helpers").GOLDEN, f"rau_hop_C{C}.npz"))
    cfg, pm, q, X, c, h, mk, ups = mg.hop_fixture_inputs(C, int(z["seed"]))
    Pm = O.views(cfg, "mult", pm)
    score, dop, p, c2, h2, cache = O.hop_fwd(Pm, cfg, q, X, c, h, mk)
    for got, key in ((score, "score"), (dop, "do_pred"), (p, "p"), (c2, "c2"), (h2, "h2")):
        np.testing.assert_allclose(got, z[key], rtol=1e-11, atol=1e-14)
    g = np.zeros(pm.size)
    dq, _, dc, dh = O.hop_bwd(Pm, O.views(cfg, "mult", g), cfg, cache, *ups)
    np.testing.assert_allclose(dc, z["dc"], rtol=1e-10, atol=1e-15)
    np.testing.assert_allclose(dh, z["dh"], rtol=1e-10, atol=1e-15)
    np.testing.assert_allclose(_digest(dq, 64), z["dq"], rtol=1e-9, atol=1e-14)
    for n, v in O.views(cfg, "mult", g).items():
        np.testing.assert_allclose(_digest(v), z[f"g_{n}"], rtol=1e-9, atol=1e-14, err_msg=n)


def test_encoder_varlen_fixture():
    z, cfg = load_golden("encoder_varlen.npz")
    Pe, Pr = O.views(cfg, "embed", z["p_embed"]), O.views(cfg, "rnn", z["p_rnn"])
    masks = dict(embed=z["mask_embed"].astype(np.float64), rnn=z["mask_rnn"].astype(np.float64))
    rnn_out, caches = O.encoder_fwd(Pe, Pr, cfg, z["x"], z["x_len"], masks)
    np.testing.assert_allclose(rnn_out, z["rnn_out"], rtol=1e-12)
    # F:472-478: row k holds the state after ITS last token -- length 1 rows equal a one-step run
    one, _ = O.encoder_fwd(Pe, Pr, cfg, z["x"][:, :1], np.array([1]), dict(embed=masks["embed"][:, :1], rnn=masks["rnn"][:, :1]))
    np.testing.assert_allclose(rnn_out[0], one[0], rtol=1e-12)
    gE, gR = np.zeros_like(z["p_embed"]), np.zeros_like(z["p_rnn"])
    O.encoder_bwd(Pr, O.views(cfg, "embed", gE), O.views(cfg, "rnn", gR), cfg, z["x_len"], caches, z["dq"])
    np.testing.assert_allclose(gE, z["g_embed"], rtol=1e-11, atol=1e-15)
    np.testing.assert_allclose(gR, z["g_rnn"], rtol=1e-11, atol=1e-15)


def test_noise_clip_fixture():
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "noise_clip.npz"))
    cfg = O.RauConfig()
    for g in O.GROUPS:
        grad = z[f"grad_{g}"].copy()
        n = O.noise_and_clip(cfg, grad, z[f"noise_{g}"])
        assert n == pytest.approx(float(z[f"norm_{g}"]), rel=1e-13)
        np.testing.assert_allclose(grad, z[f"out_{g}"], rtol=1e-13)
    assert float(z["norm_embed"]) > 0.1 > float(z["norm_mult"])          # one group clipped, one untouched
    assert np.linalg.norm(z["out_embed"]) == pytest.approx(0.1, rel=1e-12)
    np.testing.assert_array_equal(z["out_mult"], z["grad_mult"] + z["noise_mult"])
