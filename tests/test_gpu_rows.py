"""The persistent rows-layout tcgen05 engine (k_rows_tc.cu) against numpy float64 through rau_rows_gemm: both operand
major-ness, ragged extents (rows / K not multiples of the 128 x 256 x BK tile, TMA zero-fill and store clipping),
several work items per CTA (accumulator double buffering), and the split-K TMA reduce-add used by the weight gradients."""
import numpy as np
import pytest

from helpers import dev, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import torch
    assert torch.cuda.is_available()
    import rau_vqa_b200 as R
    return R


def _rows_gemm(ctx, A, B, a_mn, b_mn, D0=None):
    from rau_vqa_b200._ffi import check
    from rau_vqa_b200.core import fptr
    M, K = (A.shape[1], A.shape[0]) if a_mn else A.shape
    N = B.shape[1] if b_mn else B.shape[0]
    a, b = dev(A), dev(B)
    d = dev(np.zeros((M, N)) if D0 is None else D0)
    check(ctx.lib.rau_rows_gemm(ctx.h, M, N, K, fptr(a), A.shape[1], a_mn, fptr(b), B.shape[1], b_mn, fptr(d), N,
                                0 if D0 is None else 1))
    ctx.sync()
    return d.cpu().numpy()


# (M, N, K): pitches (the contiguous extent of each stored operand) must be multiples of 8
SHAPES = [(128, 256, 64), (784, 512, 512), (200, 64, 96), (8, 8, 8), (1568, 256, 512), (256, 512, 784), (512, 2048, 784),
          (40000, 512, 64), (512, 512, 6272)]


@pytest.mark.parametrize("mode,tol", [("bf16x3", 1e-4), ("bf16", 1.5e-2)])
@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("shape", SHAPES)
def test_rows_gemm_matches_numpy(R, mode, tol, a_mn, b_mn, shape):
    from rau_vqa_b200 import core
    M, N, K = shape
    if (a_mn and M % 8) or (b_mn and N % 8) or ((not a_mn or not b_mn) and K % 8):
        pytest.skip("pitch not a multiple of 8")
    ctx = R.Context(0, precision=dict(bf16x3=core.PREC_BF16X3, bf16=core.PREC_BF16)[mode])
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((K, M) if a_mn else (M, K)).astype(np.float32)
    B = rng.standard_normal((K, N) if b_mn else (N, K)).astype(np.float32)
    ref = (A.T if a_mn else A).astype(np.float64) @ (B if b_mn else B.T).astype(np.float64)
    got = _rows_gemm(ctx, A, B, a_mn, b_mn)
    assert rel_err(got, ref) <= tol
    D0 = rng.standard_normal((M, N)).astype(np.float32)
    got = _rows_gemm(ctx, A, B, a_mn, b_mn, D0)          # split-K reduce-add into an existing buffer
    assert rel_err(got, ref + D0) <= tol
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["bf16x3", "bf16", "mixed"])
def test_feature_pack_inline_masks(R, mode):
    """rau_feature_pack: dropout (keep bits drawn inline) + transpose to rows + bf16 split of the image features.
    Kept cells hold x/(1-p), dropped ones 0, the keep rate is 1-p, hops draw different masks, and the one-launch
    all-hops form draws exactly the bits of the per-hop launches."""
    import torch
    from rau_vqa_b200 import core
    from rau_vqa_b200._ffi import check, ffi
    ctx = R.Context(0, seed=5, precision=dict(bf16x3=core.PREC_BF16X3, bf16=core.PREC_BF16, mixed=core.PREC_MIXED)[mode])
    B, C, S, nHop, p = 5, 128, 196, 3, 0.3
    rng = np.random.default_rng(11)
    X = (rng.standard_normal((B, C, S)).astype(np.float32) + 3.0)      # no zeros: a zero output means "dropped"
    Xd = torch.from_numpy(X).cuda()
    outs = []
    for all_hops in (1, 0):
        out = torch.empty((nHop, B * S, C), dtype=torch.float32, device="cuda")
        check(ctx.lib.rau_feature_pack(ctx.h, ffi.cast("const float*", Xd.data_ptr()), B, C, S, nHop, p, 0x1234500, all_hops,
                                       ffi.cast("float*", out.data_ptr())))
        ctx.sync()
        outs.append(out.cpu().numpy())
    np.testing.assert_array_equal(outs[0], outs[1])
    got = outs[0]
    ref = np.transpose(X, (0, 2, 1)).reshape(B * S, C) / (1.0 - p)
    keep = got != 0.0
    tol = dict(bf16x3=1e-5, bf16=1e-2, mixed=6e-4)[mode]      # (mixed: one fp16 plane, half an ulp = 2^-12 relative)
    for h in range(nHop):
        np.testing.assert_allclose(got[h][keep[h]], ref[keep[h]], rtol=tol)
        rate = keep[h].mean()
        n = keep[h].size
        assert abs(rate - (1.0 - p)) < 5.0 * np.sqrt(p * (1.0 - p) / n) + 1e-4, rate
        # both channels of a pair and all four cells of a Philox call are used independently
        assert abs(keep[h][:, 0::2].mean() - keep[h][:, 1::2].mean()) < 0.01
    assert (keep[0] != keep[1]).mean() > 0.3 and (keep[1] != keep[2]).mean() > 0.3
    # neighbouring channels / cells are not copies of each other
    assert (keep[0][:, 0] != keep[0][:, 1]).mean() > 0.3
    assert (keep[0][0::4] != keep[0][1::4]).mean() > 0.3
    ctx.close()
