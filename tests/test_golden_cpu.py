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
