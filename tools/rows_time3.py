"""GPU time per launch of the hop's image-side products with their real epilogues (graph-replayed), at full width and
under the side stream's SM cap (RAU_TIME_CAP)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rau_vqa_b200 as R
from rau_vqa_b200 import core
from rau_vqa_b200._ffi import check, ffi

Rr = 256 * 196
SHAPES = [(Rr, 512, 256, 0, 1, 2, "dY epilogue + column sums"),
          (Rr, 512, 256, 0, 1, 3, "dY epilogue, no column sums"),
          (Rr, 512, 256, 0, 1, 0, "dY shape, plain fp32 epilogue"),
          (Rr, 512, 512, 0, 0, 4, "i_embed, tanh epilogue"),
          (Rr, 256, 512, 0, 0, 0, "Z = I Wa^T, plain epilogue"),
          (256, 512, Rr, 1, 1, 1, "gWa split-K"),
          (512, 512, Rr, 1, 1, 1, "gWi split-K")]
ctx = R.Context(0, precision=core.PREC_BF16X3)
for (M, N, K, a_mn, b_mn, red, what) in SHAPES:
    us = ffi.new("float*")
    check(ctx.lib.rau_rows_gemm_time(ctx.h, M, N, K, a_mn, b_mn, red, 30, us))
    print(f"cap={os.environ.get('RAU_TIME_CAP', '148')} M={M} N={N} K={K} epi={red}: {us[0]:.2f} us/launch  "
          f"{2.0 * M * N * K / us[0] / 1e6:.1f} alg TFLOP/s  # {what}", flush=True)
ctx.close()
