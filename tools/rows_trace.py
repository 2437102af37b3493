"""RAU_ROWS_TRACE=1: clock-stamp timeline of one skinny rows-engine launch (where do the microseconds go?)."""
import os, sys
os.environ["RAU_ROWS_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rau_vqa_b200 as R
from rau_vqa_b200 import core
from rau_vqa_b200._ffi import check, ffi
from rau_vqa_b200.core import fptr

names = ["start", "prologue", "tma0", "stage0", "stage1", "mma_done", "acc_ready", "epi_issued", "stores_drained", "end", "issued0", "issued1", "mma0_issued", "issued7", "mma7_issued"]
for (M, N, K, a_mn, b_mn, red) in [(256, 512, 512, 0, 0, 0), (256, 256, 512, 0, 0, 0), (256, 2048, 1024, 0, 0, 0), (256, 512, 2048, 0, 1, 0), (256, 512, 2048, 0, 1, 1)]:
    ctx = R.Context(0, precision=core.PREC_BF16X3)
    a = torch.randn((K, M) if a_mn else (M, K), device="cuda")
    b = torch.randn((K, N) if b_mn else (N, K), device="cuda")
    d = torch.zeros(M, N, device="cuda")
    for _ in range(3):
        check(ctx.lib.rau_rows_gemm(ctx.h, M, N, K, fptr(a), a.shape[1], a_mn, fptr(b), b.shape[1], b_mn, fptr(d), N, red))
    out = ffi.new("uint64_t[]", 16 * 148)
    check(ctx.lib.rau_rows_trace(ctx.h, out, 16 * 148))
    t = np.array(list(out), dtype=np.int64).reshape(148, 16)
    act = t[:, 0] > 0
    rel = (t[act][:, :15] - t[act][:, :1])
    print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: {act.sum()} CTAs; median cycles since CTA start:")
    print("   " + "  ".join(f"{n}={int(np.median(rel[:, i]))}" for i, n in enumerate(names)))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        check(ctx.lib.rau_rows_gemm(ctx.h, M, N, K, fptr(a), a.shape[1], a_mn, fptr(b), b.shape[1], b_mn, fptr(d), N, red))
    e1.record()
    torch.cuda.synchronize()
    print(f"   {e0.elapsed_time(e1) * 50:.1f} us per call (2 packs + product, back to back)")
    ctx.close()
