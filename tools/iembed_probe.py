"""Times the dominant kernel (i_embed projection, F:240) alone, as launched inside the training step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, torch
import rau_vqa_b200 as R
from rau_vqa_b200 import core
from rau_vqa_b200._ffi import check, ffi
from rau_vqa_b200.core import fptr

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=256)
ap.add_argument("--C", type=int, default=512)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--precision", default="bf16x3")
a = ap.parse_args()
cfg = R.RauConfig(C=a.C)
ctx = R.Context(0, precision=dict(f32=0, bf16=1, bf16x3=2)[a.precision])
P = torch.rand(cfg.group_size(2), device="cuda") * 0.16 - 0.08
X = torch.relu(torch.randn(a.B, a.C, 196, device="cuda"))
ms = ffi.new("float*")
check(ctx.lib.rau_time_iembed(ctx.h, cfg.c(), a.B, fptr(P), fptr(X), a.iters, ms))
fl = 2.0 * cfg.M * a.C * 196 * a.B
print(f"B={a.B} C={a.C} {a.precision}: {ms[0]*1e3:.1f} us/launch, {fl/ms[0]/1e9:.1f} algorithmic TFLOP/s")
