"""python tools/ncu_sweep.py [precision] [C] [B]: one pass of rau_sweep_attention (every image-side kernel of an answering
unit launched alone, twice each) -- the target of ncu captures, e.g.
  ncu --set full --import-source on -k regex:rows_gemm_kernel -s 4 -c 1 ... python tools/ncu_sweep.py mixed   # the dY product
(rows_gemm launches in order: i_embed x2, Z x2, dY x2, gWa x2, gWi x2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rau_vqa_b200 as R
from rau_vqa_b200._ffi import check, ffi
from rau_vqa_b200.core import fptr

prec = sys.argv[1] if len(sys.argv) > 1 else "mixed"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 512
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
cfg = R.RauConfig(V=16384, C=C, nHop=1, N=2000)
ctx = R.Context(0, seed=1, precision=dict(f32=0, bf16=1, bf16x3=2, mixed=3, f16img=4)[prec])
dev = torch.device("cuda", 0)
mult = torch.rand(cfg.group_size(2), device=dev) * 0.16 - 0.08
X = torch.relu(torch.randn(B, C, 196, device=dev))
us = ffi.new("float[9]")
check(ctx.lib.rau_sweep_attention(ctx.h, cfg.c(), B, fptr(mult), fptr(X), int(os.environ.get("ITERS", "1")),
                                  int(os.environ.get("FLUSH", "0")), us))
names = ["pack", "i_embed", "Z", "score", "softmax_sum", "bwd_dp_dz", "dY", "gWa", "gWi"]
print(prec, f"C={C} B={B}", {n: round(us[k], 1) for k, n in enumerate(names)})
ctx.close()
