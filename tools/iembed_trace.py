"""RAU_ROWS_TRACE=1 python tools/iembed_trace.py [precision]: in-kernel clock stamps of the i_embed product
(rows_gemm_kernel<EPI_TANH>), first item of every CTA and the CTA's end."""
import os, sys
os.environ["RAU_ROWS_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rau_vqa_b200 as R
from rau_vqa_b200._ffi import check, ffi
from rau_vqa_b200.core import fptr

prec = sys.argv[1] if len(sys.argv) > 1 else "mixed"
C, B = 512, 256
cfg = R.RauConfig(V=16384, C=C, nHop=8, N=2000)
ctx = R.Context(0, seed=1, precision=dict(f32=0, bf16=1, bf16x3=2, mixed=3, f16img=4)[prec])
dev = torch.device("cuda", 0)
mult = torch.rand(cfg.group_size(2), device=dev) * 0.16 - 0.08
X = torch.relu(torch.randn(B, C, 196, device=dev))
ms = ffi.new("float*")
check(ctx.lib.rau_time_iembed(ctx.h, cfg.c(), B, fptr(mult), fptr(X), 3, ms))
out = ffi.new("uint64_t[]", 16 * 148)
check(ctx.lib.rau_rows_trace(ctx.h, out, 16 * 148))
t = np.array(list(out), dtype=np.int64).reshape(148, 16)
names = ["start", "prologue", "tma0", "stage0", "stage1", "mma_done", "acc_ready", "epi_issued", "stores_drained", "end",
         "issued0", "issued1", "mma0_issued", "issued7", "mma7_issued"]
act = t[:, 0] > 0
rel = t[act][:, :15] - t[act][:, :1]
print(f"i_embed {prec}: {ms[0] * 1e3:.1f} us per launch; {act.sum()} CTAs; median / max cycles since CTA start:")
for i, n in enumerate(names):
    col = rel[:, i][t[act][:, i] > 0]
    if len(col):
        print(f"   {n:15s} median {int(np.median(col)):7d}   min {int(col.min()):7d}   max {int(col.max()):7d}   (n={len(col)})")
s = t[act][:, 0]
ns = t[act][:, 15]
print(f"   CTA lifetime: median {np.median(ns) / 1e3:.2f} us, max {ns.max() / 1e3:.2f} us; clock = {np.median((t[act][:, 9] - t[act][:, 0]) / np.maximum(ns, 1)) * 1e3:.0f} MHz")
print("   CTA start spread:", int(s.max() - s.min()), "cycles")
ctx.close()
