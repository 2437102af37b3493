"""Inference (rau_predict) samples/s on one GPU: eval-mode forward of all hops + uni/select merge."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rau_vqa_b200 as R
from rau_vqa_b200 import core

nHop, C, B = 8, 512, 256
cfg = R.RauConfig(V=16384, C=C, nHop=nHop, N=2000)
dev = torch.device("cuda", 0)
ctx = R.Context(0, seed=1)
gen = torch.Generator(device=dev).manual_seed(1)
P = [(torch.rand(cfg.group_size(g), device=dev, generator=gen) * 0.16 - 0.08) for g in range(3)]
rng = np.random.default_rng(0)
X = torch.from_numpy(np.maximum(rng.standard_normal((B, C, 196), dtype=np.float32), 0)).to(dev)
lens = rng.integers(8, 27, B)
tok = rng.integers(2, cfg.V + 1, (cfg.T, B))
for b in range(B):
    tok[lens[b]:, b] = 1
tok = torch.from_numpy(tok.astype(np.float32)).to(dev)
lens_t = torch.from_numpy(lens.astype(np.float32)).to(dev)
for _ in range(3):
    core.predict(ctx, cfg, P, X, tok, lens_t, max_len=26)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 10
for _ in range(n):
    core.predict(ctx, cfg, P, X, tok, lens_t, max_len=26)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"rau_predict Ours_Full B={B}: {ms:.3f} ms/batch, {B / ms * 1e3:.0f} samples/s (eager launches, one stream)")
ctx.close()
