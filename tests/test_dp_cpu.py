"""Data-parallel host logic on CPU (gloo, world_size 2): sharding + sum-all-reduce of the flat gradients reproduces the
single-rank big-batch gradient.  The per-rank step is the float64 oracle (this is a test of the
sharding / scaling / reduction contract, not of the kernels)."""
import os
import socket

import numpy as np
import pytest

from conftest import small_cfg
from oracle import rau_oracle as O
from rau_vqa_b200 import parallel


def test_shard_rows_partition():
    for B in (1, 7, 8, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_rows(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_rows(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = small_cfg(nHop=2)
        params = O.init_params(cfg, seed=31)
        X, x, x_len, y = O.synth_batch(cfg, B, seed=32, min_len=1)
        masks = O.synth_masks(cfg, B, seed=33)
        lo, hi = parallel.shard_rows(B, rank, world)
        Xs, xs, ls, ys = parallel.shard_batch(X, x, x_len, y, rank, world)
        ms = dict(embed=masks["embed"][:, lo:hi], rnn=masks["rnn"][:, lo:hi],
                  hops=[{k: v[lo:hi] for k, v in h.items()} for h in masks["hops"]])
        res = O.feval(cfg, params, Xs, xs, ls, ys, masks=ms, clip=False)
        # the oracle averages over its local batch; rau_batch.B_global makes librau scale by 1/B_global instead
        scale = (hi - lo) / B
        grads = [torch.from_numpy(res.grads[g] * scale) for g in O.GROUPS]
        loss = torch.from_numpy(res.loss * scale)
        parallel.allreduce_flat(grads + [loss])
        if rank == 0:
            q.put(([g.numpy() for g in grads], loss.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_step_equals_big_batch():
    import torch.multiprocessing as mp
    B, world = 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    grads, loss = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = small_cfg(nHop=2)
    params = O.init_params(cfg, seed=31)
    X, x, x_len, y = O.synth_batch(cfg, B, seed=32, min_len=1)
    masks = O.synth_masks(cfg, B, seed=33)
    ref = O.feval(cfg, params, X, x, x_len, y, masks=masks, clip=False)
    for got, g in zip(grads, O.GROUPS):
        np.testing.assert_allclose(got, ref.grads[g], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(loss[:cfg.nHop], ref.loss[:cfg.nHop], rtol=1e-9)
