"""The nn.Module surface (model.ATTLSTM, model.DeepLSTM, model.RAU, utils.model_utils) against the oracle and the
committed cell fixtures: construction path of the scripts (create -> getParameters -> clone(...)), table-shaped
inputs/outputs, gradient accumulation across shared clones."""
import os

import numpy as np
import pytest

from conftest import small_cfg
from helpers import GOLDEN, dev, lib_cfg, rel_err
from oracle import rau_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    import torch
    assert torch.cuda.is_available(), "the gpu tests need a B200"


def _set_linear(lin, W, b):
    lin.weight.copy_(dev(W))
    lin.bias.copy_(dev(b))


def test_attlstm_module_matches_fixture():
    from rau_vqa_b200.model import ATTLSTM
    z = np.load(os.path.join(GOLDEN, "cells_toy.npz"))
    H, M = z["att_Whh"].shape[1], z["att_Wx"].shape[1]
    m = ATTLSTM.LSTM.create(M, H, 1, 0.0).cuda()
    _set_linear(m.modules[0], z["att_Wx"], z["att_bx"])
    _set_linear(m.modules[1], z["att_Whh"], z["att_bhh"])
    inp = [dev(z["att_x"]), dev(z["att_c"]), dev(z["att_h"])]
    c2, h2 = m.forward(inp)
    assert rel_err(c2.cpu().numpy(), z["att_c2"]) <= TOL and rel_err(h2.cpu().numpy(), z["att_h2"]) <= TOL
    dx, dc, dh = m.backward(inp, [dev(z["att_dc2"]), dev(z["att_dh2"])])
    for got, want in ((dx, "att_dx"), (dc, "att_dc"), (dh, "att_dh")):
        assert rel_err(got.cpu().numpy(), z[want]) <= TOL, want
    for lin, gw, gb in ((m.modules[0], "att_gWx", "att_gbx"), (m.modules[1], "att_gWhh", "att_gbhh")):
        assert rel_err(lin.gradWeight.cpu().numpy(), z[gw]) <= TOL
        assert rel_err(lin.gradBias.cpu().numpy(), z[gb]) <= TOL
    # accGradParameters accumulates (clones share gradWeight, F:339-347): a second backward doubles it
    m.backward(inp, [dev(z["att_dc2"]), dev(z["att_dh2"])])
    assert rel_err(m.modules[0].gradWeight.cpu().numpy(), 2 * z["att_gWx"]) <= TOL


def test_deeplstm_module_matches_fixture_with_dropout_between_layers():
    from rau_vqa_b200.model import DeepLSTM
    z = np.load(os.path.join(GOLDEN, "cells_toy.npz"))
    cfg = small_cfg()
    m = DeepLSTM.LSTM.create(cfg.embed, cfg.Hq, 2, 0.5).cuda()
    for L in (1, 2):
        _set_linear(m.modules[2 * (L - 1)], z[f"deep_P_l{L}.Wi"], z[f"deep_P_l{L}.bi"])
        _set_linear(m.modules[2 * (L - 1) + 1], z[f"deep_P_l{L}.Wh"], z[f"deep_P_l{L}.bh"])
    m.noise_override = [None, dev(z["deep_mask"])]
    m.training()
    inp = [dev(z["deep_x"]), dev(z["deep_s0"])]
    s1 = m.forward(inp)
    assert rel_err(s1.cpu().numpy(), z["deep_s1"]) <= TOL
    dx, ds0 = m.backward(inp, dev(z["deep_ds1"]))
    assert rel_err(dx.cpu().numpy(), z["deep_dx"]) <= TOL
    assert rel_err(ds0.cpu().numpy(), z["deep_ds0"]) <= TOL
    for L in (1, 2):
        assert rel_err(m.modules[2 * (L - 1)].gradWeight.cpu().numpy(), z[f"deep_g_l{L}.Wi"]) <= TOL
        assert rel_err(m.modules[2 * (L - 1) + 1].gradBias.cpu().numpy(), z[f"deep_g_l{L}.bh"]) <= TOL
    # evaluate(): dropout is the identity (F:667-668)
    m.evaluate()
    s_eval = m.forward(inp).cpu().numpy()
    Pr = {k[len("deep_P_"):]: z[k] for k in z.files if k.startswith("deep_P_")}
    ref, _ = O.deeplstm_fwd(Pr, cfg, z["deep_x"], z["deep_s0"], None)
    assert rel_err(s_eval, ref) <= TOL


def test_getparameters_clone_share_path():
    """F:322-347: getParameters() flattens and re-points; clone('weight','bias','gradWeight','gradBias') shares."""
    import torch
    from rau_vqa_b200.model import DeepLSTM
    from rau_vqa_b200.utils import model_utils
    cfg = small_cfg()
    proto = DeepLSTM.LSTM.create(cfg.embed, cfg.Hq, 2, 0.0).cuda()
    flat, gflat = proto.getParameters()
    assert flat.numel() == O.group_size(cfg, "rnn")
    flat.uniform_(-0.08, 0.08)                                    # F:352-354 writes through to every view
    assert proto.modules[0].weight.data_ptr() == flat.data_ptr()
    clones = [proto.clone("weight", "bias", "gradWeight", "gradBias") for _ in range(3)]
    clones += model_utils.clone_many_times(proto, 2)
    for c in clones:
        assert c.modules[3].weight.data_ptr() == proto.modules[3].weight.data_ptr()
        assert c.modules[3].gradBias.data_ptr() == proto.modules[3].gradBias.data_ptr()
    # unrolled BPTT over 3 shared clones accumulates into the ONE flat gradient, same as the oracle's loop
    rng = np.random.default_rng(0)
    B = 4
    xs = [rng.standard_normal((B, cfg.embed)) for _ in range(3)]
    s = torch.zeros(B, cfg.Q, device="cuda")
    states = [s]
    for t in range(3):
        states.append(clones[t].forward([dev(xs[t]), states[-1]]))
    ds = dev(rng.standard_normal((B, cfg.Q)))
    for t in (2, 1, 0):
        dx, ds = clones[t].backward([dev(xs[t]), states[t]], ds)
    Pr = O.views(cfg, "rnn", flat.cpu().numpy().astype(np.float64))
    gP = O.views(cfg, "rnn", np.zeros(flat.numel()))
    st, caches = np.zeros((B, cfg.Q)), []
    for t in range(3):
        st, c = O.deeplstm_fwd(Pr, cfg, xs[t].astype(np.float32).astype(np.float64), st, None)
        caches.append(c)
    assert rel_err(states[-1].cpu().numpy(), st) <= TOL
    # (the oracle's gradient needs the same upstream ds; recompute with the identical random draw)
    rng = np.random.default_rng(0)
    [rng.standard_normal((B, cfg.embed)) for _ in range(3)]
    d = rng.standard_normal((B, cfg.Q)).astype(np.float32).astype(np.float64)
    for t in (2, 1, 0):
        _, d = O.deeplstm_bwd(Pr, gP, cfg, caches[t], None, d)
    ref = np.concatenate([gP[n].ravel() for n, _ in O.rnn_param_shapes(cfg)])
    assert rel_err(gflat.cpu().numpy(), ref) <= TOL
    f2, g2 = model_utils.combine_all_parameters(proto, clones[0])
    assert f2.numel() == flat.numel()                              # shared storages are laid out once (MU:88-109)


def test_multimodal_module_hop_matches_oracle_including_dX():
    from rau_vqa_b200.model.RAU import Multimodal
    cfg = small_cfg(nHop=1)
    lc = lib_cfg(cfg)
    rng = np.random.default_rng(3)
    B = 3
    mod = Multimodal(lc).cuda()
    flat, gflat = mod.getParameters()
    pm = rng.uniform(-0.3, 0.3, flat.numel())
    flat.copy_(dev(pm))
    Pm = O.views(cfg, "mult", pm.astype(np.float32).astype(np.float64))
    q, X = rng.standard_normal((B, cfg.Q)), np.maximum(rng.standard_normal((B, cfg.C, cfg.S)), 0)
    c, h = rng.standard_normal((B, cfg.H)), rng.standard_normal((B, cfg.H))
    mk = O.synth_masks(cfg, B, seed=9)["hops"][0]
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    score, dop, p, c2, h2, cache = O.hop_fwd(Pm, cfg, f32(q), f32(X), f32(c), f32(h), mk)
    mod.masks = dict(q=dev(mk["q"].astype(np.uint8)), x=dev(mk["X"].astype(np.uint8)), m=dev(mk["m"].astype(np.uint8)))
    mod.training()
    inp = [dev(q), dev(X), dev(c), dev(h)]
    out = mod.forward(inp)
    for got, want in zip(out, (score, dop, p, c2, h2)):
        assert rel_err(got.cpu().numpy(), want) <= TOL
    ups = [f32(rng.standard_normal(t.shape)) for t in (score, dop, p, c2, h2)]
    gP = O.views(cfg, "mult", np.zeros(pm.size))
    dq, dX, dc, dh = O.hop_bwd(Pm, gP, cfg, cache, *ups, want_dX=True)
    gi = mod.backward(inp, [dev(u) for u in ups], want_dX=True)
    for got, want, name in zip(gi, (dq, dX, dc, dh), "dq dX dc dh".split()):
        assert rel_err(got.cpu().numpy(), want) <= TOL, name
    ref = np.concatenate([gP[n].ravel() for n, _ in O.mult_param_shapes(cfg)])
    assert rel_err(gflat.cpu().numpy(), ref) <= TOL
    # evaluate(): no dropout anywhere
    mod.evaluate()
    s_eval = mod.forward(inp)[0].cpu().numpy()
    assert rel_err(s_eval, O.hop_fwd(Pm, cfg, f32(q), f32(X), f32(c), f32(h), None)[0]) <= TOL


def test_word_embed_module_scatter_adds_repeated_tokens():
    from rau_vqa_b200.model.RAU import WordEmbed
    cfg = small_cfg()
    lc = lib_cfg(cfg)
    rng = np.random.default_rng(5)
    we = WordEmbed(lc).cuda()
    E = rng.uniform(-1, 1, (cfg.V, cfg.embed))
    we.weight.copy_(dev(E))
    ids = np.array([3, 7, 3, 1, 50, 3], dtype=np.float64)        # repeats and both ends of the vocabulary
    mask = (rng.random((6, cfg.embed)) >= 0.5).astype(np.uint8)
    we.noise_override = dev(mask)
    we.training()
    e = we.forward(dev(ids))
    ref, cache = O.word_embed_fwd(E.astype(np.float32).astype(np.float64), ids, mask.astype(np.float64), 0.5)
    assert rel_err(e.cpu().numpy(), ref) <= 1e-5
    de = rng.standard_normal((6, cfg.embed)).astype(np.float32)
    we.backward(dev(ids), dev(de))
    gE = np.zeros_like(E)
    O.word_embed_bwd(gE, cache, mask.astype(np.float64), 0.5, de.astype(np.float64))
    assert rel_err(we.gradWeight.cpu().numpy(), gE) <= 1e-5
    assert np.all(we.gradWeight.cpu().numpy()[10] == 0)
