"""Oracle parity ON THE BASELINE CONFIGURATIONS, through the schedule bench.py measures.

Every case runs rau_train_step with DRAWN (Philox) dropout masks -- the all-hops feature pack, the batched mask launches,
three streams, and after two eager calls the captured whole-step CUDA graph, replayed -- with lr = 0, no noise and no
clipping, so that the raw gradients stay in place and the parameters do not move.  rau_draw_masks then exports the keep
masks that step drew, and the float64 oracle (F:445-650 restated, oracle/rau_oracle.py) runs on the same inputs, weights and
masks.  Tolerance: north_star's 1e-3 relative, on EVERY named tensor (26 of the answering units, 8 of the encoder, E), on the
logits, attention maps and losses; argmax answers bit-exact wherever the oracle's top-1 / top-2 margin clears the tolerance.
The per-tensor worst cases are written to gpurun_out/parity_<case>.json (copied to profiles/ by hand after a GPU run)."""
import json
import os

import numpy as np
import pytest

from helpers import assert_grads_per_tensor, dev, rel_err
from oracle import rau_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-3
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib_cfg(cfg):
    import rau_vqa_b200 as R
    return R.RauConfig(V=cfg.V, embed=cfg.embed, Hq=cfg.Hq, nlayer=cfg.nlayer, C=cfg.C, S=cfg.S, M=cfg.M, A=cfg.A, H=cfg.H,
                       N=cfg.N, nHop=cfg.nHop, T=cfg.T)


def _graph_step_vs_oracle(name, cfg, B, seed, precision=None, replays=4, step_t=3, hop_mask=None):
    import torch
    import rau_vqa_b200 as R
    from rau_vqa_b200 import core
    lc = _lib_cfg(cfg)
    params = O.init_params(cfg, seed=seed)
    X, x, x_len, y = O.synth_batch(cfg, B, seed=seed + 1)
    ctx = R.Context(0, seed=seed + 2)
    if precision is not None:
        ctx.set_precision(precision)
    P = [dev(params[g]) for g in O.GROUPS]
    P0 = [p.clone() for p in P]
    G = [torch.zeros_like(p) for p in P]
    ST = [[torch.zeros_like(p), torch.zeros_like(p)] for p in P]
    out = R.StepBuffers(lc, B, P[0].device)
    Xd, xd, ld, yd = dev(X), dev(x), dev(x_len), dev(y)
    l0 = ctx.launches
    per_call = []
    for rep in range(replays):   # calls 1-2 eager, call 3 captures + launches the graph, call 4 replays it
        core.train_step(ctx, lc, P, G, ST, Xd, xd, ld, yd, out, optim=core.OPT_ADAM, lrs=(0.0, 0.0, 0.0),
                        hyper=(0.9, 0.999, 1e-8), eta=0.0, gamma=0.55, clip=1e9, hop_mask=hop_mask, step_t=step_t,
                        max_len=int(x_len.max()), B_global=B)
        ctx.sync()
        per_call.append(ctx.launches - l0)
        l0 = ctx.launches
    for p, p0 in zip(P, P0):
        assert torch.equal(p, p0)          # lr = 0: the parameters did not move
    mk = core.draw_masks(ctx, lc, B, step_t)
    masks = dict(embed=mk["embed"].cpu().numpy(), rnn=mk["rnn"].cpu().numpy(),
                 hops=[dict(q=mk["q"][h].cpu().numpy(), X=mk["x"][h].cpu().numpy(), m=mk["m"][h].cpu().numpy())
                       for h in range(cfg.nHop)])
    keep = float(np.mean([hm["X"].mean() for hm in masks["hops"]]))
    assert abs(keep - (1 - cfg.p_x)) < 0.01, keep
    p32 = {g: params[g].astype(np.float32).astype(np.float64) for g in O.GROUPS}
    res = O.feval(cfg, p32, X.astype(np.float32).astype(np.float64), x, x_len, y, masks=masks, hop_mask=hop_mask, clip=False)
    grads = {g: G[i].cpu().numpy().astype(np.float64) for i, g in enumerate(O.GROUPS)}
    report = {}
    sc = out.scores.cpu().numpy()
    att = out.attprob.cpu().numpy()
    dp = out.do_pred.cpu().numpy()
    ans = out.answers.cpu().numpy().astype(np.int64)
    worst_fwd = 0.0
    for h in range(cfg.nHop):
        e = (rel_err(sc[h], res.scores[h]), rel_err(att[h], res.attprob[h]), rel_err(dp[h], res.do_pred[h]))
        report[f"hop{h}.score/attprob/do_pred"] = e
        worst_fwd = max(worst_fwd, *e)
    from helpers import per_tensor_rel_err
    for g in O.GROUPS:
        for tname, e in per_tensor_rel_err(cfg, g, grads[g], res.grads[g]).items():
            report[f"{g}.{tname}"] = e
    # the measured errors are written out BEFORE anything is asserted (tools/precision_table.py also records the modes that fail)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    mode = {0: "f32", 1: "bf16", 2: "bf16x3", 3: "mixed", 4: "f16img"}.get(int(ctx.lib.rau_get_precision(ctx.h)), "?")
    with open(os.path.join(ROOT, "gpurun_out", f"parity_{name}_{mode}.json"), "w") as f:
        json.dump(dict(case=name, precision=mode, B=B, nHop=cfg.nHop, C=cfg.C, N=cfg.N, launches_per_call=per_call,
                       worst_forward=worst_fwd, per_tensor=report), f, indent=1, default=float)
    try:
        for h in range(cfg.nHop):
            top2 = np.sort(res.scores[h], axis=1)[:, -2:]
            safe = (top2[:, 1] - top2[:, 0]) > 4 * TOL * np.abs(res.scores[h]).max()
            assert safe.sum() >= max(1, B // 2)
            np.testing.assert_array_equal(ans[h][safe], res.answers[h][safe])      # argmax bit-exact
        loss = out.loss.cpu().numpy()
        np.testing.assert_allclose(loss, res.loss, rtol=TOL)
        np.testing.assert_allclose(out.loss_do_pred.cpu().numpy(), res.loss_do_pred, rtol=10 * TOL, atol=1e-6)
        norms = out.norms.cpu().numpy()
        for i, g in enumerate(O.GROUPS):
            assert norms[i] == pytest.approx(np.linalg.norm(res.grads[g]), rel=TOL)
        worst = assert_grads_per_tensor(cfg, grads, res.grads, TOL, report={})
        assert worst_fwd <= TOL, report
        # the 3rd and 4th call ran as ONE graph launch each: the library counts the captured kernels, and the eager calls
        # before them launched the same number
        assert per_call[-1] == per_call[-2] and per_call[-1] > 0
    finally:
        ctx.close()
    return worst, report


@pytest.mark.parametrize("mode", ["default", "bf16x3"])
def test_ours_full_b256_graph_schedule_matches_oracle(mode):
    """BASELINE.json configs[2] -- the configuration bench.py reports: Ours_Full, nHop 8, C 512, batch 256 -- in the default
    precision mode (RAU_PREC_MIXED) and in bf16x3."""
    from rau_vqa_b200 import core
    cfg = O.RauConfig(V=16384, C=512, nHop=8, N=2000)
    _graph_step_vs_oracle("ours_full_b256", cfg, 256, seed=2301, precision=None if mode == "default" else core.PREC_BF16X3)


def test_ours_ms_b64_matches_oracle():
    """BASELINE.json configs[1]: Ours_MS, 3 answering units, batch 64."""
    cfg = O.RauConfig(V=16384, C=512, nHop=3, N=2000)
    _graph_step_vs_oracle("ours_ms_b64", cfg, 64, seed=2311)


def test_ours_resnet_nhop8_matches_oracle():
    """BASELINE.json configs[3] at a batch the oracle finishes in seconds: ResNet-101 features (C = 2048), nHop 8; 160 rows
    = two row tiles of the persistent recurrence, the second one ragged."""
    cfg = O.RauConfig(V=16384, C=2048, nHop=8, N=2000)
    _graph_step_vs_oracle("ours_resnet_b160", cfg, 160, seed=2321)


def test_c1024_and_early_stopped_hops_match_oracle():
    """The attention sweep's middle width (C = 1024, configs[4]) with Ours_Full's early-stop table switched on for two hops
    (F:414-428, F:587-589) and a 1000-way head (F:222's comment)."""
    cfg = O.RauConfig(V=4000, C=1024, nHop=4, N=1000)
    _graph_step_vs_oracle("c1024_b48", cfg, 48, seed=2331, hop_mask=[1, 0, 1, 0])


def test_ours_ss_b8_matches_oracle():
    """BASELINE.json configs[0]: Ours_SS, one answering unit, batch 8 (the CPU-runnable case)."""
    cfg = O.RauConfig(V=16384, C=512, nHop=1, N=2000)
    _graph_step_vs_oracle("ours_ss_b8", cfg, 8, seed=2341)


def test_predict_at_reference_dims_uses_the_hoisted_image_side():
    """predict_result (F:652-724) at the reference's dimensions: the hop-invariant image side (feature transpose, I, Z) is
    formed once for all hops (rau_step.cu rau_predict); logits / attention vs the oracle, argmax bit-exact."""
    import rau_vqa_b200 as R
    cfg = O.RauConfig(V=16384, C=512, nHop=8, N=2000)
    lc = _lib_cfg(cfg)
    B = 40
    params = O.init_params(cfg, seed=2351)
    X, x, x_len, _ = O.synth_batch(cfg, B, seed=2352)
    ctx = R.Context(0)
    p32 = {g: params[g].astype(np.float32).astype(np.float64) for g in O.GROUPS}
    preds, atts = O.predict(cfg, p32, X.astype(np.float32).astype(np.float64), x, x_len)
    l0 = ctx.launches
    pred, att = R.predict(ctx, lc, [dev(params[g]) for g in O.GROUPS], dev(X), dev(x), dev(x_len), max_len=int(x_len.max()))
    ctx.sync()
    launches = ctx.launches - l0
    for k in range(cfg.nHop + 2):
        assert rel_err(pred[k].cpu().numpy(), preds[k]) <= TOL, k
        assert rel_err(att[k].cpu().numpy(), atts[k]) <= TOL, k
        top2 = np.sort(preds[k], axis=1)[:, -2:]
        safe = (top2[:, 1] - top2[:, 0]) > 4 * TOL * np.abs(preds[k]).max()
        assert safe.sum() >= B // 2
        np.testing.assert_array_equal(pred[k].cpu().numpy().argmax(1)[safe], preds[k].argmax(1)[safe])
    assert launches < 400, launches      # (one image-side pass, not nHop of them)
    ctx.close()


def test_predict_answers_open_ended_and_multiple_choice_bit_exact():
    """rau_predict_answers: the test loop's answer extraction (F:903-918) on the device at the reference's dimensions.  The
    open-ended answer is the argmax of each of the nHop+2 tables; the multiple-choice answer is the argmax of pred * mask
    with the reference's multiplicative 0/1 mask over the 18 candidates (ans_mc, 0 = empty slot).  Both are compared bit for
    bit with numpy on the library's own tables, and with the oracle's tables wherever its top-2 margin clears the tolerance."""
    import rau_vqa_b200 as R
    cfg = O.RauConfig(V=16384, C=512, nHop=3, N=2000)
    lc = _lib_cfg(cfg)
    B, nmc = 24, 18
    params = O.init_params(cfg, seed=2371)
    X, x, x_len, _ = O.synth_batch(cfg, B, seed=2372)
    rng = np.random.default_rng(2373)
    mc = np.zeros((B, nmc))
    for b in range(B):
        k = rng.integers(1, nmc + 1)
        mc[b, :k] = rng.choice(cfg.N, size=k, replace=False) + 1          # 1-based candidate ids, the rest stays 0 = empty
    ctx = R.Context(0)
    P = [dev(params[g]) for g in O.GROUPS]
    oe, mca, pred, att = R.predict_answers(ctx, lc, P, dev(X), dev(x), dev(x_len), mc_choices=dev(mc), max_len=int(x_len.max()),
                                           want_tables=True)
    oe2, none = R.predict_answers(ctx, lc, P, dev(X), dev(x), dev(x_len), max_len=int(x_len.max()))
    ctx.sync()
    assert none is None
    pred = pred.cpu().numpy()
    mask = np.zeros((B, cfg.N), dtype=np.float32)
    for b in range(B):
        for v in mc[b]:
            if v != 0:
                mask[b, int(v) - 1] = 1
    np.testing.assert_array_equal(oe.cpu().numpy(), pred.argmax(2) + 1)
    np.testing.assert_array_equal(oe2.cpu().numpy(), pred.argmax(2) + 1)
    np.testing.assert_array_equal(mca.cpu().numpy(), (pred * mask[None]).argmax(2) + 1)
    p32 = {g: params[g].astype(np.float32).astype(np.float64) for g in O.GROUPS}
    preds, _ = O.predict(cfg, p32, X.astype(np.float32).astype(np.float64), x, x_len)
    for k in range(cfg.nHop + 2):
        top2 = np.sort(preds[k], axis=1)[:, -2:]
        safe = (top2[:, 1] - top2[:, 0]) > 4 * TOL * np.abs(preds[k]).max()
        assert safe.sum() >= B // 2
        np.testing.assert_array_equal(oe.cpu().numpy()[k][safe], preds[k].argmax(1)[safe] + 1)
        mcs = preds[k] * mask
        t2 = np.sort(mcs, axis=1)[:, -2:]
        safe = (t2[:, 1] - t2[:, 0]) > 4 * TOL * np.abs(preds[k]).max()
        np.testing.assert_array_equal(mca.cpu().numpy()[k][safe], mcs.argmax(1)[safe] + 1)
    ctx.close()


def test_train_graph_survives_a_larger_validation_batch():
    """a captured training step bakes arena addresses in; a validation pass with a larger batch re-allocates
    arena buffers.  train (graph) -> predict (larger B) -> train must equal train -> train."""
    import torch
    import rau_vqa_b200 as R
    from rau_vqa_b200 import core
    cfg = O.RauConfig(V=2000, C=128, nHop=2, N=300)
    lc = _lib_cfg(cfg)
    B = 16
    params = O.init_params(cfg, seed=2361)
    X, x, x_len, y = O.synth_batch(cfg, B, seed=2362)
    Xv, xv, lv, _ = O.synth_batch(cfg, 4 * B, seed=2363)
    res = []
    for with_predict in (False, True):
        ctx = R.Context(0, seed=5)
        P = [dev(params[g]) for g in O.GROUPS]
        G = [torch.zeros_like(p) for p in P]
        ST = [[torch.zeros_like(p), torch.zeros_like(p)] for p in P]
        out = R.StepBuffers(lc, B, P[0].device, want_scores=False)      # (the step then uses its own arena buffers)
        args = (dev(X), dev(x), dev(x_len), dev(y))
        for it in range(1, 7):
            core.train_step(ctx, lc, P, G, ST, *args, out, step_t=it, opt_t=it, max_len=26)
            if with_predict and it == 4:      # the step has been captured by now (calls 3+ replay)
                R.predict(ctx, lc, P, dev(Xv), dev(xv), dev(lv), max_len=26)
        ctx.sync()
        res.append(([p.cpu().numpy() for p in P], out.loss.cpu().numpy().copy()))
        ctx.close()
    # (two contexts, atomics in different orders: 1e-5-level noise; a step that ran on freed buffers differs by O(1))
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=1e-4)
    for a, b in zip(res[0][0], res[1][0]):
        assert rel_err(a, b) <= 1e-4
