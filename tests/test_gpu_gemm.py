"""The tcgen05 contraction engine against numpy float64 through rau_gemm: every operand major-ness, ragged
extents (not multiples of the 128 x BN x 64 tile), accumulate, and both tensor-core precision modes."""
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


def _gemm(ctx, A, B, ta, tb, C0=None):
    from rau_vqa_b200._ffi import check
    from rau_vqa_b200.core import fptr
    M, K = (A.shape[1], A.shape[0]) if ta else A.shape
    N = B.shape[1] if tb else B.shape[0]
    a, b = dev(A), dev(B)
    c = dev(np.zeros((M, N)) if C0 is None else C0)
    check(ctx.lib.rau_gemm(ctx.h, M, N, K, fptr(a), A.shape[1], ta, fptr(b), B.shape[1], tb, fptr(c), N, 0 if C0 is None else 1))
    ctx.sync()
    return c.cpu().numpy()


SHAPES = [(128, 128, 64), (256, 208, 512), (4, 2048, 512), (2048, 4, 200), (130, 70, 96), (1, 1, 8), (512, 196, 520),
          (300, 2000, 33), (64, 512, 6656)]


@pytest.mark.parametrize("mode,tol", [("bf16x3", 1e-4), ("bf16", 1.5e-2)])
@pytest.mark.parametrize("ta", [0, 1])
@pytest.mark.parametrize("tb", [0, 1])
@pytest.mark.parametrize("shape", SHAPES)
def test_tc_gemm_matches_numpy(R, mode, tol, ta, tb, shape):
    from rau_vqa_b200 import core
    M, N, K = shape
    ctx = R.Context(0, precision=dict(bf16x3=core.PREC_BF16X3, bf16=core.PREC_BF16)[mode])
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    B = rng.standard_normal((K, N) if tb else (N, K)).astype(np.float32)
    ref = (A.T if ta else A).astype(np.float64) @ (B if tb else B.T).astype(np.float64)
    got = _gemm(ctx, A, B, ta, tb)
    assert rel_err(got, ref) <= tol
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    got = _gemm(ctx, A, B, ta, tb, C0)
    assert rel_err(got, ref + C0) <= tol
    ctx.close()


def test_f32_mode_is_the_cuda_core_engine(R):
    from rau_vqa_b200 import core
    ctx = R.Context(0, precision=core.PREC_F32)
    rng = np.random.default_rng(0)
    A, B = rng.standard_normal((100, 300)).astype(np.float32), rng.standard_normal((50, 300)).astype(np.float32)
    assert rel_err(_gemm(ctx, A, B, 0, 0), A.astype(np.float64) @ B.astype(np.float64).T) <= 1e-6
    ctx.close()
