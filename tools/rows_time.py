"""Time the rows-layout tcgen05 engine on the answering unit's product shapes (B = 256 images of 196 cells)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rau_vqa_b200 as R
from rau_vqa_b200 import core
from rau_vqa_b200._ffi import check
from rau_vqa_b200.core import fptr

Rr = 256 * 196
SHAPES = [  # (M, N, K, a_mn, b_mn, reduce, what)
    (Rr, 512, 512, 0, 0, 0, "i_embed fwd  I = Xd Wi^T          (C=512)"),
    (Rr, 256, 512, 0, 0, 0, "attbycontent E = I Wa^T"),
    (Rr, 512, 256, 0, 1, 0, "dI = dZ Wa"),
    (256, 512, Rr, 1, 1, 1, "gWa += dZ^T I      (split-K)"),
    (512, 512, Rr, 1, 1, 1, "gWi += dY^T Xd     (split-K, C=512)"),
    (Rr, 512, 2048, 0, 0, 0, "i_embed fwd                      (C=2048)"),
    (512, 2048, Rr, 1, 1, 1, "gWi += dY^T Xd     (split-K, C=2048)"),
    (8192, 8192, 2048, 0, 0, 0, "square-ish peak check"),
]
for mode, name in ((core.PREC_BF16X3, "bf16x3"), (core.PREC_BF16, "bf16")):
    ctx = R.Context(0, precision=mode)
    for (M, N, K, a_mn, b_mn, red, what) in SHAPES:
        a = torch.randn((K, M) if a_mn else (M, K), device="cuda")
        b = torch.randn((K, N) if b_mn else (N, K), device="cuda")
        d = torch.zeros(M, N, device="cuda")
        def run():
            check(ctx.lib.rau_rows_gemm(ctx.h, M, N, K, fptr(a), a.shape[1], a_mn, fptr(b), b.shape[1], b_mn, fptr(d), N, red))
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 5
        e0.record()
        for _ in range(it):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / it
        print(f"{name} M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: {us:.1f} us incl. operand packing, "
              f"{2.0 * M * N * K / us / 1e6:.1f} algorithmic TFLOP/s  # {what}", flush=True)
    ctx.close()
