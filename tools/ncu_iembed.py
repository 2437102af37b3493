"""python tools/ncu_iembed.py [precision] [C] [B]: launches the i_embed product (rows_gemm_kernel<EPI_TANH>) a few times --
the target of `ncu --set full -k regex:rows_gemm_kernel`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rau_vqa_b200 as R
from rau_vqa_b200 import core
from rau_vqa_b200._ffi import check, ffi
from rau_vqa_b200.core import fptr

prec = sys.argv[1] if len(sys.argv) > 1 else "mixed"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 512
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
cfg = R.RauConfig(V=16384, C=C, nHop=8, N=2000)
ctx = R.Context(0, seed=1, precision=dict(f32=0, bf16=1, bf16x3=2, mixed=3, f16img=4)[prec])
dev = torch.device("cuda", 0)
mult = torch.rand(cfg.group_size(2), device=dev) * 0.16 - 0.08
X = torch.relu(torch.randn(B, C, 196, device=dev))
ms = ffi.new("float*")
check(ctx.lib.rau_time_iembed(ctx.h, cfg.c(), B, fptr(mult), fptr(X), int(os.environ.get("ITERS", "3")), ms))
print(f"i_embed {prec} C={C} B={B}: {ms[0] * 1e3:.1f} us per launch")
ctx.close()
