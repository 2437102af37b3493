"""RAU_PHASES=1 python tools/phases.py [workload] [precision]: per-phase milliseconds of one eager training step."""
import os, sys
os.environ.setdefault("RAU_PHASES", "1")   # 2 = %globaltimer stamps that survive graph capture (both streams)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rau_vqa_b200 as R
from rau_vqa_b200 import core
from rau_vqa_b200._ffi import check

wl = sys.argv[1] if len(sys.argv) > 1 else "ours_full"
nHop, C, B = dict(ours_full=(8, 512, 256), ours_resnet=(8, 2048, 256), ours_ms=(3, 512, 64))[wl]
cfg = R.RauConfig(V=16384, C=C, nHop=nHop, N=2000)
ctx = R.Context(0, seed=123)
if len(sys.argv) > 2:
    ctx.set_precision(dict(f32=0, bf16=1, bf16x3=2, mixed=3, f16img=4)[sys.argv[2]])
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(123)
P = [(torch.rand(cfg.group_size(g), device=dev, generator=gen) * 0.16 - 0.08) for g in range(3)]
G = [torch.zeros_like(p) for p in P]
ST = [[torch.zeros_like(p), torch.zeros_like(p)] for p in P]
out = R.StepBuffers(cfg, B, dev, want_scores=False)
rng = np.random.default_rng(0)
X = torch.from_numpy(np.maximum(rng.standard_normal((B, C, 196), dtype=np.float32), 0)).to(dev)
lens = rng.integers(8, 27, B)
tok = rng.integers(2, cfg.V + 1, (cfg.T, B))
for b in range(B):
    tok[lens[b]:, b] = 1
tok = torch.from_numpy(tok.astype(np.float32)).to(dev)
lens_t = torch.from_numpy(lens.astype(np.float32)).to(dev)
y = torch.from_numpy(rng.integers(1, cfg.N + 1, B).astype(np.float32)).to(dev)
for it in range(8 if os.environ["RAU_PHASES"] == "2" else 4):
    core.train_step(ctx, cfg, P, G, ST, X, tok, lens_t, y, out, optim=core.OPT_ADAM, lrs=(3e-3, 3e-3, 3e-4),
                    hyper=(0.9, 0.999, 1e-8), eta=0.01, gamma=0.55, clip=0.1, step_t=it, max_len=26, B_global=B)
    ctx.sync()
    print(f"--- {wl} step {it} ({ctx.launches} launches so far)", file=sys.stderr)
    check(ctx.lib.rau_phase_report(ctx.h))
ctx.close()
