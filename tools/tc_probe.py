"""Diagnostic sweep of the tcgen05 engine through rau_gemm: prints the error of every case instead of stopping."""
import os, sys
os.environ.setdefault("RAU_TC_MIN_WORK", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rau_vqa_b200 as R
from rau_vqa_b200 import core
from rau_vqa_b200._ffi import check
from rau_vqa_b200.core import fptr

def run(ctx, M, N, K, ta, tb):
    rng = np.random.default_rng(1)
    A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    B = rng.standard_normal((K, N) if tb else (N, K)).astype(np.float32)
    ref = (A.T if ta else A).astype(np.float64) @ (B if tb else B.T).astype(np.float64)
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    c = torch.zeros(M, N, device="cuda")
    check(ctx.lib.rau_gemm(ctx.h, M, N, K, fptr(a), A.shape[1], ta, fptr(b), B.shape[1], tb, fptr(c), N, 0))
    ctx.sync()
    got = c.cpu().numpy()
    return np.abs(got - ref).max() / np.abs(ref).max()

shapes = [(128, 128, 64), (128, 128, 256), (256, 256, 128), (256, 208, 512), (4, 2048, 512), (2048, 4, 200), (130, 70, 96), (512, 196, 520)]
for mode in (core.PREC_BF16, core.PREC_BF16X3):
    ctx = R.Context(0, precision=mode)
    for (M, N, K) in shapes:
        for ta in (0, 1):
            for tb in (0, 1):
                try:
                    e = run(ctx, M, N, K, ta, tb)
                    print(f"mode={mode} M={M} N={N} K={K} ta={ta} tb={tb} rel_err={e:.3e}", flush=True)
                except Exception as ex:
                    print(f"mode={mode} M={M} N={N} K={K} ta={ta} tb={tb} FAILED {ex}", flush=True)
    ctx.close()
