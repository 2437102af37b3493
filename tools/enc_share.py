"""How much of the graph-replayed step is the question encoder?  Times the Ours_Full step with max_len = 26, 13 and 2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rau_vqa_b200 as R
from rau_vqa_b200 import core

nHop, C, B = 8, 512, 256
cfg = R.RauConfig(V=16384, C=C, nHop=nHop, N=2000)
dev = torch.device("cuda", 0)
for max_len in (26, 13, 2):
    ctx = R.Context(0, seed=123)
    gen = torch.Generator(device=dev).manual_seed(123)
    P = [(torch.rand(cfg.group_size(g), device=dev, generator=gen) * 0.16 - 0.08) for g in range(3)]
    G = [torch.zeros_like(p) for p in P]
    ST = [[torch.zeros_like(p), torch.zeros_like(p)] for p in P]
    out = R.StepBuffers(cfg, B, dev, want_scores=False)
    rng = np.random.default_rng(0)
    X = torch.from_numpy(np.maximum(rng.standard_normal((B, C, 196), dtype=np.float32), 0)).to(dev)
    lens = rng.integers(1, max_len + 1, B)
    tok = rng.integers(2, cfg.V + 1, (cfg.T, B))
    for b in range(B):
        tok[lens[b]:, b] = 1
    tok = torch.from_numpy(tok.astype(np.float32)).to(dev)
    lens_t = torch.from_numpy(lens.astype(np.float32)).to(dev)
    y = torch.from_numpy(rng.integers(1, cfg.N + 1, B).astype(np.float32)).to(dev)
    def step(i):
        core.train_step(ctx, cfg, P, G, ST, X, tok, lens_t, y, out, optim=core.OPT_ADAM, lrs=(3e-3, 3e-3, 3e-4),
                        hyper=(0.9, 0.999, 1e-8), eta=0.01, gamma=0.55, clip=0.1, step_t=i, max_len=max_len, B_global=B)
    for i in range(6):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        step(6 + i)
    e1.record()
    torch.cuda.synchronize()
    print(f"max_len {max_len}: {e0.elapsed_time(e1) / 10:.3f} ms/step", flush=True)
    ctx.close()
