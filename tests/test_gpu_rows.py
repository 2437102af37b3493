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
