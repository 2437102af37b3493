#!/usr/bin/env python
"""bench.py -- RAU training-step throughput (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus 1 --steps K --warmup W                 # our arm (librau.so through the C ABI)
  python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the reference algorithm on host cores
  torchrun ... bench.py --gpus N ...                            # N > 1: one rank per GPU, NCCL all-reduce of the grads

A "step" is one whole training iteration of the experiment scripts (feval F:445-650 + the three optimizer calls
F:787-791) on one synthetic batch: encoder unroll -> nHop answering units -> joint loss -> BPTT -> [all-reduce] ->
noise -> per-group clip -> adam.  `value` has the batch resident in HBM; `e2e` copies the batch from pinned host
memory every step and reads the loss vector back.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nHop, C, per-GPU batch, description)
    "ours_full": (8, 512, 256, "Ours_Full training step, VGG16 pool5 14x14x512, nHop 8, batch 256/GPU, adam"),
    "ours_resnet": (8, 2048, 256, "Ours_ResNet training step, ResNet-101 14x14x2048, nHop 8, batch 256/GPU, adam"),
    "ours_ms": (3, 512, 64, "Ours_MS training step, VGG16 pool5 14x14x512, 3 answering units, batch 64, adam"),
    "ours_ss": (1, 512, 8, "Ours_SS training step, VGG16 pool5 14x14x512, 1 answering unit, batch 8, adam"),
}
METRIC = "RAU fwd+bwd+update samples/sec"
UNIT = "samples/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def oracle_cfg(nHop, C):
    from oracle import rau_oracle as O
    return O.RauConfig(V=16384, C=C, nHop=nHop, N=2000)


def cpu_reference_arm(workload, sample_B, steps, warmup):
    """The reference algorithm on the host cores: the float64 numpy restatement (oracle/rau_oracle.py, 'port'),
    one array op per reference nn module, BLAS-threaded; a bounded sample of the workload (sample_B rows)."""
    from oracle import rau_oracle as O
    nHop, C, _, _ = WORKLOADS[workload]
    cfg = oracle_cfg(nHop, C)
    params = O.init_params(cfg, seed=123)
    X, x, x_len, y = O.synth_batch(cfg, sample_B, seed=123)
    masks = O.synth_masks(cfg, sample_B, seed=7)
    opt = {}
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        rng = np.random.default_rng(it)
        noise = {g: rng.standard_normal(params[g].size) * O.noise_std(cfg, it) for g in O.GROUPS}
        O.train_step(cfg, params, opt, X, x, x_len, y, masks=masks, noise=noise, optim="adam")
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    return sample_B / sec, sec


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload
    nHop, C, B, desc = WORKLOADS[workload]
    sample_B = args.cpu_sample or 32
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    v, sec = cpu_reference_arm(workload, sample_B, steps, warmup)
    cores = os.cpu_count()
    sample = f"{steps} steps of the same step at batch {sample_B} (of {B}), float64 numpy, BLAS threads = host cores"
    line = dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warmup,
                ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", config=dict(workload=desc, nHop=nHop, C=C, batch_per_gpu=B, cpu_batch=sample_B),
                cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
                note="Torch7 cannot run in this image (no LuaJIT/torch rocks); this is the oracle port of the reference algorithm")
    emit(line)


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries loaded later (NCCL prints its version banner to stdout when the
    environment sets NCCL_DEBUG) must not add lines: everything written to fd 1 from here on goes to stderr, and emit()
    writes the result line to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def bind_to_gpu_numa(local_rank: int):
    """Multi-GPU runs: pin this rank to the CPUs next to its GPU before the pinned staging buffers are allocated, so the
    per-step host -> device uploads of the e2e leg (103 MB per rank and step) read NUMA-local memory.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local_rank])
                                              if os.environ.get("CUDA_VISIBLE_DEVICES") else local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception as e:   # no NVML, no permission, odd CUDA_VISIBLE_DEVICES: keep the default placement
        print(f"[bench] NUMA binding skipped: {e}", file=sys.stderr)
    return None


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ours_full", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--precision", default="default", choices=["default", "f32", "bf16", "bf16x3"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="rows of the CPU baseline sample (0 = 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_main(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    import rau_vqa_b200 as R
    from rau_vqa_b200 import core

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU path (use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nHop, C, B, desc = WORKLOADS[args.workload]
    B = args.batch or B
    cfg = R.RauConfig(V=16384, C=C, nHop=nHop, N=2000)
    ctx = R.Context(local, seed=123)
    if args.precision != "default":
        ctx.set_precision(dict(f32=core.PREC_F32, bf16=core.PREC_BF16, bf16x3=core.PREC_BF16X3)[args.precision])
    prec = {core.PREC_F32: "f32", core.PREC_BF16: "bf16", core.PREC_BF16X3: "bf16x3"}[int(ctx.lib.rau_get_precision(ctx.h))]
    if world > 1:
        ids = [core.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], rank, world)

    numa = bind_to_gpu_numa(local) if world > 1 else None

    dev = torch.device("cuda", local)
    gen = torch.Generator(device=dev).manual_seed(123)          # same parameters on every rank
    P = [(torch.rand(cfg.group_size(g), device=dev, generator=gen) * 0.16 - 0.08) for g in range(3)]   # F:352-354
    G = [torch.zeros_like(p) for p in P]
    ST = [[torch.zeros_like(p), torch.zeros_like(p)] for p in P]
    out = R.StepBuffers(cfg, B, dev, want_scores=False)
    # synthetic batches; NB rotating batches so that the inputs alone exceed the 126 MB L2
    NB = max(2, int(np.ceil(160e6 / (B * C * 196 * 4))))
    rng = np.random.default_rng(1000 + rank)
    host = []
    for _ in range(NB):
        X = np.maximum(rng.standard_normal((B, C, 196), dtype=np.float32), 0)
        lens = rng.integers(8, 27, B)
        tok = rng.integers(2, cfg.V + 1, (cfg.T, B))
        for b in range(B):
            tok[lens[b]:, b] = 1
        y = rng.integers(1, cfg.N + 1, B)
        host.append(tuple(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).pin_memory() for a in (X, tok, lens, y)))
    resident = [tuple(t.to(dev) for t in hb) for hb in host]
    stage = tuple(torch.empty_like(t) for t in resident[0])
    loss_host = torch.empty(cfg.nHop + 2, dtype=torch.float32).pin_memory()
    h2d = sum(t.numel() * 4 for t in host[0])
    d2h = loss_host.numel() * 4
    step_no = [0]

    def step(batch):
        core.train_step(ctx, cfg, P, G, ST, batch[0], batch[1], batch[2], batch[3], out, optim=core.OPT_ADAM,
                        lrs=(3e-3, 3e-3, 3e-4), hyper=(0.9, 0.999, 1e-8), eta=0.01, gamma=0.55, clip=0.1,
                        step_t=step_no[0], max_len=26, B_global=B * world)
        step_no[0] += 1

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
            ms = t.item()
        return ms

    def resident_step(i):
        step(resident[i % NB])

    # end to end: every step's batch comes from pinned host memory and its loss vector goes back to the host, inside the
    # timed region.  The upload is double buffered on a copy stream (the pinned async feed of SURVEY.md 8f rank 3): batch
    # i+1 crosses PCIe while step i computes; the step waits on its own batch's copy event.
    stages = [stage, tuple(torch.empty_like(t) for t in resident[0])]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[slot])
            for s_, t in zip(stages[slot], host[i % NB]):
                s_.copy_(t, non_blocking=True)
            ready[slot].record(copy_stream)

    e2e_state = {"next": 0}

    def e2e_step(i):
        cur = torch.cuda.current_stream(dev)
        if e2e_state["next"] <= i:          # first step of a timed run: nothing was prefetched yet
            prefetch(i)
            e2e_state["next"] = i + 1
        prefetch(i + 1)                     # overlaps with this step's compute
        e2e_state["next"] = i + 2
        slot = i % 2
        cur.wait_event(ready[slot])
        step(stages[slot])
        freed[slot].record(cur)
        loss_host.copy_(out.loss, non_blocking=True)

    # untimed priming: every rotating batch is seen often enough for its step to be captured as a CUDA graph
    for i in range(3 * NB + args.warmup):
        resident_step(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches
    ms = timed(resident_step, args.steps)
    launches = ctx.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    for ev in freed:
        ev.record(torch.cuda.current_stream(dev))
    for i in range(6):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_state["next"] = 0
    ms_e2e = timed(e2e_step, args.steps)
    torch.cuda.synchronize()
    assert torch.isfinite(out.loss).all().item(), "loss is not finite"

    value = B * world * args.steps / (ms * 1e-3)
    e2e_v = B * world * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        pk = peaks()
        # dominant kernel: the i_embed projection I = tanh(Wi drop(X) + bi) (F:240), 2*M*C*196 flop per image
        ms_k = ffi_time_iembed(ctx, cfg, B, P[2], resident[0][0])
        flops = 2.0 * cfg.M * cfg.C * 196 * B
        ach = flops / (ms_k * 1e-3) / 1e12
        # DRAM bytes of that kernel per launch from the committed `ncu --set full` capture (profiles/roofline_traffic.json,
        # keyed by workload and precision mode); null when no capture of this configuration is committed
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            traffic = tj.get(f"{args.workload}:{prec}:B{B}")
        except Exception:
            pass
        executed = 3 if prec == "bf16x3" else 1
        roof = dict(bound="tensor", kernel="rows_gemm_kernel<EPI_TANH>: i_embed product I = tanh(Wi drop(X) + bi) (F:238-242), engine " + prec,
                    achieved=ach, peak=pk["tensor_burst"], unit="TFLOP/s", frac=ach / pk["tensor_burst"], traffic=traffic,
                    peak_source=pk["src"] + " bf16 burst (kernel timed alone, 20 back-to-back launches, CUDA events)",
                    ms_per_launch=ms_k, flop_per_launch=flops, algorithmic="2*196*B*M*C flop per launch (SURVEY 8d)",
                    mma_passes=executed, executed_frac=executed * ach / pk["tensor_burst"])
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sb = args.cpu_sample or 16
            v, sec = cpu_reference_arm(args.workload, sb, 2, 1)
            cpu = dict(value=v, unit=UNIT, cores=os.cpu_count(), kind="port",
                       sample=f"2 steps of the same step at batch {sb} (of {B}), float64 numpy oracle, {sec:.2f} s/step")
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=prec,
                    data="synthetic",
                    config=dict(workload=desc, nHop=nHop, C=C, batch_per_gpu=B, global_batch=B * world, T=26, N=2000,
                                V=16384, parallelism=f"dp{world}",
                                l2=f"{NB} rotating batches ({NB * B * C * 196 * 4 / 1e6:.0f} MB of features) + >1 GB of saved "
                                   "activations per step exceed the 126 MB L2"),
                    e2e=dict(value=e2e_v, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=ms_e2e / args.steps),
                    gpu_launches=int(launches), launches_per_step=launches / args.steps, clocks=clocks, roofline=roof,
                    cpu_baseline=cpu)
        emit(line)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def ffi_time_iembed(ctx, cfg, B, mult, X):
    from rau_vqa_b200._ffi import check, ffi
    from rau_vqa_b200.core import fptr
    ms = ffi.new("float*")
    check(ctx.lib.rau_time_iembed(ctx.h, cfg.c(), B, fptr(mult), fptr(X), 20, ms))
    return float(ms[0])


if __name__ == "__main__":
    main()
