"""The batch feed (rau_feed_*, rau_feat_cache_*) against the direct call with device-resident
tensors: what arrives through the pinned double-buffered upload -- float32 staging, float16 staging, or a gather from the
HBM-resident fp16 feature cache -- trains exactly like the same batch handed over as device pointers (F:452-456, LD:1009)."""
import numpy as np
import pytest

from helpers import dev
from oracle import rau_oracle as O

pytestmark = pytest.mark.gpu


def _setup(B=12, seed=3101):
    import rau_vqa_b200 as R
    cfg = O.RauConfig(V=2000, C=128, nHop=2, N=300)
    lc = R.RauConfig(V=cfg.V, C=cfg.C, nHop=cfg.nHop, N=cfg.N)
    params = O.init_params(cfg, seed=seed)
    batches = [O.synth_batch(cfg, B, seed=seed + 1 + i) for i in range(5)]
    return R, cfg, lc, params, batches


def _run(R, lc, params, batches, mode):
    """five adam steps; mode: 'direct' | FEED_F32 | FEED_F16 | 'direct_f16' (direct call on fp16-rounded features)"""
    import torch
    from rau_vqa_b200 import core, feed as F
    ctx = R.Context(0, seed=9)
    P = [dev(params[g]) for g in O.GROUPS]
    G = [torch.zeros_like(p) for p in P]
    ST = [[torch.zeros_like(p), torch.zeros_like(p)] for p in P]
    B = batches[0][0].shape[0]
    out = R.StepBuffers(lc, B, P[0].device, want_scores=False)
    losses = []
    if mode in ("direct", "direct_f16"):
        for it, (X, x, x_len, y) in enumerate(batches, start=1):
            Xs = X.astype(np.float32)
            if mode == "direct_f16":
                Xs = Xs.astype(np.float16).astype(np.float32)
            core.train_step(ctx, lc, P, G, ST, dev(Xs), dev(x), dev(x_len), dev(y), out, step_t=it, opt_t=it,
                            max_len=int(x_len.max()))
            ctx.sync()
            losses.append(out.loss.cpu().numpy().copy())
    else:
        fd = F.Feed(ctx, lc, B, fmt=mode, depth=2)
        assert fd.bytes_per_batch == B * lc.C * lc.S * (4 if mode == F.FEED_F32 else 2) + 4 * (lc.T * B + 2 * B)
        fd.fill(0, *batches[0])          # (float64 features, as the loader returns them: LD:1009)
        fd.submit(0)
        for it in range(1, len(batches) + 1):
            slot = (it - 1) % 2
            if it < len(batches):        # the next batch crosses PCIe under this step
                fd.fill(1 - slot, *batches[it])
                fd.submit(1 - slot)
            b = fd.acquire(slot)
            assert b.max_len == int(batches[it - 1][2].max()) and b.B == B
            F.train_step_batch(ctx, lc, P, G, ST, b, out, step_t=it, opt_t=it)
            fd.release(slot)
            ctx.sync()
            losses.append(out.loss.cpu().numpy().copy())
        fd.close()
    res = ([p.cpu().numpy() for p in P], np.stack(losses))
    ctx.close()
    return res


NOISE = 1e-4     # two runs of the SAME path differ by fp32 summation order (atomics, split-K reduce-adds): ~1e-6 .. 1e-5


def _same(a, b):
    from helpers import rel_err
    np.testing.assert_allclose(a[1], b[1], rtol=NOISE)
    for x, y in zip(a[0], b[0]):
        assert rel_err(x, y) <= NOISE


def test_f32_feed_equals_direct_call():
    R, cfg, lc, params, batches = _setup()
    from rau_vqa_b200 import feed as F
    _same(_run(R, lc, params, batches, "direct"), _run(R, lc, params, batches, F.FEED_F32))


def test_f16_feed_equals_direct_call_on_fp16_rounded_features():
    """the fp16 staging delivers exactly fp16(x); at toy size (fp32 feature path) that equals the direct call on rounded x"""
    R, cfg, lc, params, batches = _setup()
    from rau_vqa_b200 import feed as F
    _same(_run(R, lc, params, batches, "direct_f16"), _run(R, lc, params, batches, F.FEED_F16))


def test_f16_feed_changes_no_bit_of_what_the_tensor_pipe_sees_in_the_default_mode():
    """In RAU_PREC_MIXED the features enter the tensor pipe as fp16(x / (1 - p)) and the reference's dropout scale
    1 / (1 - 0.5) = 2 commutes with the rounding, so the fp16 staging loses nothing: the packed operand of the i_embed product
    (rau_feature_pack, a deterministic kernel) is bit-identical whether it is formed from x or from fp16(x) -- and a training
    run through the fp16 feed equals the direct float32 call to summation-order noise."""
    import torch
    import rau_vqa_b200 as R
    from rau_vqa_b200 import core, feed as F
    from rau_vqa_b200._ffi import check, ffi
    ctx = R.Context(0, seed=5)
    assert int(ctx.lib.rau_get_precision(ctx.h)) == core.PREC_MIXED
    Bq, C, S, nHop = 3, 128, 196, 2
    rng = np.random.default_rng(7)
    X = np.maximum(rng.standard_normal((Bq, C, S)).astype(np.float32), 0)
    outs = []
    for Xin in (X, X.astype(np.float16).astype(np.float32)):
        out = torch.empty((nHop, Bq * S, C), dtype=torch.float32, device="cuda")
        Xd = torch.from_numpy(Xin).cuda()
        check(ctx.lib.rau_feature_pack(ctx.h, ffi.cast("const float*", Xd.data_ptr()), Bq, C, S, nHop, 0.5, 0x4200, 1,
                                       ffi.cast("float*", out.data_ptr())))
        ctx.sync()
        outs.append(out.cpu().numpy())
    np.testing.assert_array_equal(outs[0], outs[1])
    assert (outs[0] != 0).mean() > 0.15
    ctx.close()
    cfg = O.RauConfig(V=3000, C=512, nHop=2, N=2000)
    lc = R.RauConfig(V=cfg.V, C=cfg.C, nHop=cfg.nHop, N=cfg.N)
    params = O.init_params(cfg, seed=3201)
    batches = [O.synth_batch(cfg, 6, seed=3202 + i) for i in range(3)]
    _same(_run(R, lc, params, batches, "direct"), _run(R, lc, params, batches, F.FEED_F16))


def test_f16_direct_feed_trains_like_the_f16_feed():
    """RAU_FEED_F16_DIRECT: no widening pass on the copy stream -- the all-hops feature pack reads the uploaded fp16 buffer
    (rau_batch.feats_f16, feats == NULL).  half -> float is exact, so the packed operand is the one the F16 feed produces and
    the run equals it to summation-order noise; the upload is the same number of bytes."""
    import rau_vqa_b200 as R
    from rau_vqa_b200 import feed as F
    cfg = O.RauConfig(V=3000, C=512, nHop=2, N=2000)
    lc = R.RauConfig(V=cfg.V, C=cfg.C, nHop=cfg.nHop, N=cfg.N)
    params = O.init_params(cfg, seed=3301)
    batches = [O.synth_batch(cfg, 6, seed=3302 + i) for i in range(4)]
    _same(_run(R, lc, params, batches, F.FEED_F16), _run(R, lc, params, batches, F.FEED_F16_DIRECT))


def test_f16_only_batch_is_refused_where_float32_features_are_needed():
    """a batch that carries only feats_f16 is accepted by the training step on the tcgen05 rows path; the toy shape (CUDA-core
    feature path) and the inference entry points say so instead of reading a NULL pointer"""
    import torch
    import rau_vqa_b200 as R
    from rau_vqa_b200 import feed as F
    from rau_vqa_b200._ffi import RauError, ffi
    from rau_vqa_b200 import core
    Rm, cfg, lc, params, batches = _setup(B=4)
    ctx = R.Context(0, seed=9, precision=core.PREC_F32)   # (the CUDA-core engine: no all-hops feature pack)
    P = [dev(params[g]) for g in O.GROUPS]
    G = [torch.zeros_like(p) for p in P]
    ST = [[torch.zeros_like(p), torch.zeros_like(p)] for p in P]
    out = R.StepBuffers(lc, 4, P[0].device, want_scores=False)
    fd = F.Feed(ctx, lc, 4, fmt=F.FEED_F16_DIRECT, depth=2)
    fd.fill(0, *batches[0])
    fd.submit(0)
    b = fd.acquire(0)
    assert b.feats == ffi.NULL and b.feats_f16 != ffi.NULL
    with pytest.raises(RauError, match="feats is NULL"):
        F.train_step_batch(ctx, lc, P, G, ST, b, out, step_t=1, opt_t=1)
    pred = torch.empty(lc.nHop + 2, 4, lc.N, device="cuda")
    Pp = ffi.new("float*[3]", [ffi.cast("float*", t.data_ptr()) for t in P])
    with pytest.raises(RauError, match="float32 features"):
        from rau_vqa_b200._ffi import check
        check(ctx.lib.rau_predict(ctx.h, lc.c(), b, Pp, ffi.cast("float*", pred.data_ptr()), ffi.NULL))
    fd.release(0)
    fd.close()
    ctx.close()


def test_feature_cache_gather():
    import torch
    import rau_vqa_b200 as R
    from rau_vqa_b200 import feed as F
    ctx = R.Context(0)
    n, C, S = 150, 64, 196
    rng = np.random.default_rng(1)
    feats = np.maximum(rng.standard_normal((n, C, S), dtype=np.float32), 0)
    cache = F.FeatCache(ctx, n, C, S)
    cache.put(0, feats[:100])
    cache.put(100, feats[100:])
    idx = rng.integers(1, n + 1, 37)
    got = cache.gather(dev(idx))
    ctx.sync()
    np.testing.assert_array_equal(got.cpu().numpy(), feats[idx - 1].astype(np.float16).astype(np.float32))
    got16 = cache.gather_f16(dev(idx))      # the same rows without widening (rau_batch.feats_f16)
    ctx.sync()
    assert got16.dtype == torch.float16
    np.testing.assert_array_equal(got16.cpu().numpy(), feats[idx - 1].astype(np.float16))
    from rau_vqa_b200._ffi import RauError
    with pytest.raises(RauError):
        cache.put(140, feats[:20])       # past the end
    cache.close()
    ctx.close()
