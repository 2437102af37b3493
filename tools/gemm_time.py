"""Kernel-only timing of the tcgen05 engine on the step's contraction shapes.  Run plain (prints the shape order), then
under `ncu --metrics gpu__time_duration.sum` to read the per-launch device time of tc_gemm_kernel (pack kernels apart)."""
import os, sys
os.environ.setdefault("RAU_TC_MIN_WORK", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rau_vqa_b200 as R
from rau_vqa_b200 import core
from rau_vqa_b200._ffi import check
from rau_vqa_b200.core import fptr

SHAPES = [  # (M, N, K, ta, tb, what)
    (50176, 512, 512, 0, 0, "i_embed rows=(image,cell): K-major x K-major"),
    (50176, 512, 512, 0, 1, "same, B MN-major"),
    (50176, 512, 512, 1, 0, "same, A MN-major"),
    (50176, 256, 512, 0, 0, "Wa.I rows layout"),
    (512, 512, 50176, 1, 1, "gWi wgrad, both MN-major, K = rows"),
    (256, 2048, 512, 0, 0, "LSTM recurrent step"),
    (256, 2048, 1024, 0, 0, "LSTM step, K = in+H"),
    (6656, 2048, 512, 0, 0, "hoisted encoder projection"),
    (8192, 8192, 2048, 0, 0, "square-ish peak check"),
]
PICK = [int(i) for i in os.environ.get("GT_SHAPES", "").split(",") if i] or list(range(len(SHAPES)))
MODES = [m for m in ((core.PREC_BF16, "bf16"), (core.PREC_BF16X3, "bf16x3")) if m[1] in os.environ.get("GT_MODES", "bf16,bf16x3").split(",")]
for mode, name in MODES:
    ctx = R.Context(0, precision=mode)
    for (M, N, K, ta, tb, what) in [SHAPES[i] for i in PICK]:
        a = torch.randn((K, M) if ta else (M, K), device="cuda")
        b = torch.randn((K, N) if tb else (N, K), device="cuda")
        c = torch.zeros(M, N, device="cuda")
        def go():
            check(ctx.lib.rau_gemm(ctx.h, M, N, K, fptr(a), a.shape[1], ta, fptr(b), b.shape[1], tb, fptr(c), N, 0))
        go(); ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): go()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{name} M={M} N={N} K={K} ta={ta} tb={tb}: {ms*1e3:.1f} us incl. packing, {2.0*M*N*K/ms/1e9:.1f} TFLOP/s  # {what}", flush=True)
    ctx.close()
