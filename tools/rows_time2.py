"""GPU time per launch of rows-engine products, replayed from a CUDA graph (no host launch cost in the number)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rau_vqa_b200 as R
from rau_vqa_b200 import core
from rau_vqa_b200._ffi import check, ffi

Rr = 256 * 196
SHAPES = [(256, 2048, 512, 0, 0, 0, "LSTM-like gate product (plain epilogue)"),
          (256, 512, 2048, 0, 1, 1, "dgrad K=4H split-K reduce (128 CTAs)"),
          (256, 512, 2048, 0, 1, 0, "dgrad K=4H single pass (16 CTAs)"),
          (256, 512, 512, 0, 0, 0, "Linear 512->512"),
          (256, 2000, 512, 0, 0, 0, "answer head"),
          (2048, 512, 2048, 1, 1, 1, "deferred wgrad over 8 hops"),
          (Rr, 512, 512, 0, 0, 0, "i_embed product, plain epilogue"),
          (512, 512, Rr, 1, 1, 1, "gWi split-K")]
modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["bf16x3", "bf16"]
for name in modes:
    ctx = R.Context(0, precision=dict(bf16x3=core.PREC_BF16X3, bf16=core.PREC_BF16)[name])
    for (M, N, K, a_mn, b_mn, red, what) in SHAPES:
        us = ffi.new("float*")
        check(ctx.lib.rau_rows_gemm_time(ctx.h, M, N, K, a_mn, b_mn, red, 50, us))
        print(f"{name} M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} red={red}: {us[0]:.2f} us/launch  "
              f"{2.0 * M * N * K / us[0] / 1e6:.1f} alg TFLOP/s  # {what}", flush=True)
    ctx.close()
