"""The encoder's hoisted input projections as rows-engine launches ([T*B, in] x [in, 4H] -> fp32): time per launch
(rau_rows_gemm_time) and in-kernel clock stamps of a launch (RAU_ROWS_TRACE=1)."""
import os, sys
os.environ["RAU_ROWS_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rau_vqa_b200 as R
from rau_vqa_b200 import core
from rau_vqa_b200._ffi import check, ffi

names = ["start", "prologue", "tma0", "stage0", "stage1", "mma_done", "acc_ready", "epi_issued", "stores_drained", "end",
         "issued0", "issued1", "mma0_issued", "issued7", "mma7_issued"]
ctx = R.Context(0, precision=core.PREC_BF16X3)
for (M, N, K, mode) in [(6656, 2048, 200, 0), (6656, 2048, 200, 5), (6656, 2048, 512, 0), (6656, 2048, 512, 5)]:
    us = ffi.new("float*")
    check(ctx.lib.rau_rows_gemm_time(ctx.h, M, N, K, 0, 0, mode, 10, us))
    out = ffi.new("uint64_t[]", 16 * 148)
    check(ctx.lib.rau_rows_trace(ctx.h, out, 16 * 148))
    t = np.array(list(out), dtype=np.int64).reshape(148, 16)
    act = t[:, 0] > 0
    rel = t[act][:, :15] - t[act][:, :1]
    ns = t[act][:, 15]
    print(f"M={M} N={N} K={K} mode={mode} (0 plain, 5 nn.Linear + bias): {us[0]:.1f} us per launch; {act.sum()} CTAs; lifetime median {np.median(ns) / 1e3:.1f} us max {ns.max() / 1e3:.1f} us")
    print("   " + "  ".join(f"{n}={int(np.median(rel[:, i][t[act][:, i] > 0]))}" for i, n in enumerate(names) if (t[act][:, i] > 0).any()))
ctx.close()
