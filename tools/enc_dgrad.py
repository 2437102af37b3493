"""python tools/enc_dgrad.py: the encoder backward's per-step split-K product [dH2 | du2] = dG2_t [Wh2 | Wi2]
(M = 256 batch rows, N = 1024, K = 2048, B operand MN-major, TMA reduce-add) through rau_rows_gemm -- the target of
`ncu --set full -k regex:rows_gemm_kernel` for the gate-GEMM figures of the backward recurrence."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rau_vqa_b200 as R
from rau_vqa_b200._ffi import check
from rau_vqa_b200.core import fptr

ctx = R.Context(0, precision=2)       # the recurrent chain runs bf16x3 in every mode
M, N, K = 256, 1024, 2048
a = torch.randn(M, K, device="cuda") * 1e-3
b = torch.randn(K, N, device="cuda") * 0.05
d = torch.zeros(M, N, device="cuda")
for _ in range(3):
    check(ctx.lib.rau_rows_gemm(ctx.h, M, N, K, fptr(a), K, 0, fptr(b), N, 1, fptr(d), N, 1))
ctx.sync()
ref = 3 * (a.double() @ b.double())
print("enc dgrad split-K product: rel err", float((d.double() - ref).abs().max() / ref.abs().max()))
ctx.close()
