"""Data-parallel check on real GPUs (run under torchrun, one rank per GPU): with dropout and gradient noise off, a few
rau_train_step iterations on the rank's shard of a batch (B_global = whole batch) must leave every rank with bit-identical
parameters, equal to what ONE process gets on the whole batch.  Prints one line per check."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import rau_vqa_b200 as R
from rau_vqa_b200 import core

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = R.RauConfig(V=4000, C=512, nHop=3, N=500, p_embed=0.0, p_rnn=0.0, p_q=0.0, p_x=0.0, p_m=0.0)
Bg = 64 * world
rng = np.random.default_rng(7)
X = np.maximum(rng.standard_normal((Bg, 512, 196), dtype=np.float32), 0)
lens = rng.integers(3, 27, Bg)
tok = rng.integers(2, cfg.V + 1, (cfg.T, Bg))
for b in range(Bg):
    tok[lens[b]:, b] = 1
y = rng.integers(1, cfg.N + 1, Bg)
gen = torch.Generator(device="cpu").manual_seed(3)
P0 = [(torch.rand(cfg.group_size(g), generator=gen) * 0.16 - 0.08) for g in range(3)]

def run(ctx, sl, steps=4):
    P = [p.clone().to(dev) for p in P0]
    G = [torch.zeros_like(p) for p in P]
    ST = [[torch.zeros_like(p), torch.zeros_like(p)] for p in P]
    B = sl.stop - sl.start
    out = R.StepBuffers(cfg, B, dev, want_scores=False)
    Xd = torch.from_numpy(X[sl]).to(dev); td = torch.from_numpy(tok[:, sl].astype(np.float32)).to(dev)
    ld = torch.from_numpy(lens[sl].astype(np.float32)).to(dev); yd = torch.from_numpy(y[sl].astype(np.float32)).to(dev)
    losses = []
    for it in range(steps):
        core.train_step(ctx, cfg, P, G, ST, Xd, td, ld, yd, out, optim=core.OPT_ADAM, lrs=(3e-3, 3e-3, 3e-4),
                        hyper=(0.9, 0.999, 1e-8), eta=0.0, gamma=0.55, clip=0.1, step_t=it, max_len=26, B_global=Bg)
        ctx.sync()
        losses.append(out.loss.cpu().numpy().copy())
    return P, losses

ctx = R.Context(local, seed=11)
ids = [core.Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
ctx.comm_init(ids[0], rank, world)
per = Bg // world
P, losses = run(ctx, slice(rank * per, (rank + 1) * per))
ok = True
for g in range(3):
    mine = P[g]
    ref = mine.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(mine, ref))
    ok = ok and same
    if rank == 0:
        print(f"group {g}: replicas bit-identical across ranks: {same}", flush=True)
dist.barrier()
torch.cuda.synchronize()
ctx.close()
if rank == 0:
    solo = R.Context(local, seed=11)
    Ps, ls = run(solo, slice(0, Bg))
    for g in range(3):
        d = float((Ps[g] - P[g]).abs().max() / Ps[g].abs().max())
        print(f"group {g}: max |dp - single|/max|single| after 4 adam steps = {d:.2e}", flush=True)
        ok = ok and d < 2e-3
    print("loss dp    ", np.round(losses[-1], 5), flush=True)
    print("loss single", np.round(ls[-1], 5), flush=True)
    ok = ok and np.allclose(losses[-1], ls[-1], rtol=2e-3)
    print("DP CHECK", "OK" if ok else "FAILED", flush=True)
    solo.close()
dist.barrier()
dist.destroy_process_group()
