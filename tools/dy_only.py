"""One dY-epilogue product (the hop backward's dZ Wa with the (1 - I^2) epilogue) for profiling: RAU_TIME_CAP=84 python tools/dy_only.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rau_vqa_b200 as R
from rau_vqa_b200 import core
from rau_vqa_b200._ffi import check, ffi
epi = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = R.Context(0, precision=core.PREC_BF16X3)
us = ffi.new("float*")
check(ctx.lib.rau_rows_gemm_time(ctx.h, 256 * 196, 512, 256, 0, 1, epi, 2, us))
print(us[0])
ctx.close()
