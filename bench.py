#!/usr/bin/env python
"""bench.py -- RAU training-step throughput (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus 1 --steps K --warmup W                 # our arm (librau.so through the C ABI)
  python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the reference algorithm on host cores
  torchrun ... bench.py --gpus N ...                            # N > 1: one rank per GPU, NCCL all-reduce of the grads
  python bench.py --sweep                                       # BASELINE configs[4]: the attention-kernel sweep only

A "step" is one whole training iteration of the experiment scripts (feval F:445-650 + the three optimizer calls
F:787-791) on one synthetic batch: encoder unroll -> nHop answering units -> joint loss -> BPTT -> [all-reduce] ->
noise -> per-group clip -> adam.  `value` has the batch resident in HBM; `e2e` feeds every step's batch from pinned host
memory through the library's double-buffered feed (rau_feed_*, float32 staging like the reference's upload F:452-456) and
reads the loss vector back.  One JSON line on stdout (rank 0); `extra` carries the other BASELINE configurations, the
attention-kernel sweep, the fp16-staged feed and (N > 1) a strong-scaling line.
"""
from __future__ import annotations

import os
import sys

# The CPU arm must use every host core whatever launcher started it: torchrun exports OMP_NUM_THREADS=1 to its workers,
# and BLAS reads that at import time -- fix the environment before numpy loads.
if "--impl" in sys.argv and "reference" in sys.argv:
    _n = str(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = _n

import argparse
import json
import subprocess
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
CPU_SAMPLE_B = 32          # rows of the bounded CPU sample: the SAME in cpu_baseline and in --impl reference
PREC_NAMES = {0: "f32", 1: "bf16", 2: "bf16x3", 3: "mixed", 4: "f16img"}
# dtype key of the JSON line: the arithmetic the path computes in
PREC_DTYPE = {"f32": "f32", "bf16": "bf16", "bf16x3": "bf16x3", "mixed": "f16/bf16x3 operands, f32 accumulate", "f16img": "f16"}


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
        self.nvml, self.stop_flag = None, False

    def _nvml_loop(self):
        """NVML in-process: one sample per ~4 ms (nvidia-smi's loop mode cannot go below 50 ms -- one sample per timed region)"""
        n, h = self.nvml
        bits = [("hw_slowdown", getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                ("hw_thermal_slowdown", getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                ("sw_thermal_slowdown", getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                ("sw_power_cap", getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4))]
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
                r = int(get_reasons(h))
                self.rows.append([str(sm), str(mx), "0"] + ["Active" if r & b else "Not Active" for _, b in bits])
            except Exception:
                break
            time.sleep(0.004)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(int(self.index)))
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=2)
        elif self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        else:
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


def make_config(workload, world, B=None, extra=None):
    """the `config` object of the JSON line: identical in both arms (the driver compares them)"""
    nHop, C, B0, desc = WORKLOADS[workload]
    B = B or B0
    cfg = dict(workload=desc, nHop=nHop, C=C, batch_per_gpu=B, global_batch=B * world, T=26, N=2000, V=16384,
               parallelism=f"dp{world}")
    if extra:
        cfg.update(extra)
    return cfg


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [int(i.get("num_threads", 0)) for i in threadpool_info() if i.get("user_api") == "blas"]
        return max(n) if n else None
    except Exception:
        return None


def cpu_reference_arm(workload, sample_B, steps, warmup):
    """The reference algorithm on the host cores: the float64 numpy restatement (oracle/rau_oracle.py, 'port'),
    one array op per reference nn module, BLAS-threaded; a bounded sample of the workload (sample_B rows)."""
    from oracle import rau_oracle as O
    nHop, C, _, _ = WORKLOADS[workload]
    cfg = O.RauConfig(V=16384, C=C, nHop=nHop, N=2000)
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


def cpu_baseline_dict(workload, steps, warmup, sample_B=0):
    nHop, C, B, _ = WORKLOADS[workload]
    sb = min(sample_B or CPU_SAMPLE_B, B)
    v, sec = cpu_reference_arm(workload, sb, steps, warmup)
    threads = blas_threads()
    cores = threads or len(os.sched_getaffinity(0))
    return dict(value=v, unit=UNIT, cores=cores, kind="port",
                sample=f"{steps} step(s) after {warmup} warm-up of the same step at batch {sb} (of {B}): float64 numpy restatement of "
                       f"the reference (oracle/rau_oracle.py), {threads} BLAS threads on {os.cpu_count()} host CPUs, {sec:.2f} s/step"), sec


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    steps, warmup = max(1, min(args.steps, 3)), 1
    cpu, sec = cpu_baseline_dict(args.workload, steps, warmup, args.cpu_sample)
    v = cpu["value"]
    line = dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warmup,
                ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", config=make_config(args.workload, world), cpu_baseline=cpu,
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
                note="Torch7 cannot run in this image (no LuaJIT / torch rocks, SURVEY.md 8c): this is the CPU restatement of the "
                     "reference's algorithm on a bounded sample of the workload; bench/ref_torch7_cpu.lua times the real thing "
                     "where Torch7 exists")
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
    per-step host -> device uploads of the e2e leg read NUMA-local memory.  Best effort."""
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


class Runner:
    """One configuration (workload, per-GPU batch) set up on this rank: parameters, rotating synthetic batches, the timed legs."""

    def __init__(self, ctx, workload, B, world, rank, dev, dist=None):
        import torch
        import rau_vqa_b200 as R
        self.torch, self.R, self.ctx, self.world, self.rank, self.dev, self.dist = torch, R, ctx, world, rank, dev, dist
        nHop, C, B0, _ = WORKLOADS[workload]
        self.B = B or B0
        self.cfg = cfg = R.RauConfig(V=16384, C=C, nHop=nHop, N=2000)
        gen = torch.Generator(device=dev).manual_seed(123)          # same parameters on every rank
        self.P = [(torch.rand(cfg.group_size(g), device=dev, generator=gen) * 0.16 - 0.08) for g in range(3)]   # F:352-354
        self.G = [torch.zeros_like(p) for p in self.P]
        self.ST = [[torch.zeros_like(p), torch.zeros_like(p)] for p in self.P]
        self.out = R.StepBuffers(cfg, self.B, dev, want_scores=False)
        # synthetic batches; NB rotating batches so that the inputs alone exceed the 126 MB L2
        self.NB = NB = max(2, int(np.ceil(160e6 / (self.B * C * 196 * 4))))
        rng = np.random.default_rng(1000 + rank)
        self.host = []
        for _ in range(min(NB, 8)):
            X = np.maximum(rng.standard_normal((self.B, C, 196), dtype=np.float32), 0)
            lens = rng.integers(8, 27, self.B)
            tok = rng.integers(2, cfg.V + 1, (cfg.T, self.B))
            for b in range(self.B):
                tok[lens[b]:, b] = 1
            y = rng.integers(1, cfg.N + 1, self.B)
            self.host.append(tuple(np.ascontiguousarray(a, dtype=np.float32) for a in (X, tok, lens, y)))
        self.NB = len(self.host)
        self.resident = [tuple(torch.from_numpy(a).to(dev) for a in hb) for hb in self.host]
        self.it = 0

    def step(self, batch):
        from rau_vqa_b200 import core
        self.it += 1        # `it` counts from 1 like the reference's main loop (F:783); adam's own counter equals it here
        core.train_step(self.ctx, self.cfg, self.P, self.G, self.ST, batch[0], batch[1], batch[2], batch[3], self.out,
                        optim=core.OPT_ADAM, lrs=(3e-3, 3e-3, 3e-4), hyper=(0.9, 0.999, 1e-8), eta=0.01, gamma=0.55, clip=0.1,
                        step_t=self.it, opt_t=self.it, max_len=26, B_global=self.B * self.world)

    def timed(self, fn, k):
        torch, dist = self.torch, self.dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
            ms = t.item()
        return ms

    def prime(self, extra=0):
        # untimed: every rotating batch is seen often enough for its step to be captured as a CUDA graph
        for i in range(3 * self.NB + extra):
            self.step(self.resident[i % self.NB])

    def resident_leg(self, steps):
        l0 = self.ctx.launches
        ms = self.timed(lambda i: self.step(self.resident[i % self.NB]), steps)
        return ms, self.ctx.launches - l0

    def e2e_leg(self, steps, fmt):
        """every step's batch comes from pinned host memory through the library's double-buffered feed (the upload of batch
        i+1 runs on the copy stream under step i) and the loss vector goes back to the host, all inside the timed region"""
        torch = self.torch
        from rau_vqa_b200 import feed as F
        fd = F.Feed(self.ctx, self.cfg, self.B, fmt=fmt, depth=2)
        for s in range(2):      # the loader's side of the contract, outside the timed region: the staging holds the batches
            fd.fill(s, *self.host[s % self.NB])
        loss_host = torch.empty(self.cfg.nHop + 2, dtype=torch.float32).pin_memory()
        state = {"submitted": -1}

        def step(i, last=False):
            if state["submitted"] < i:
                fd.submit(i % 2)
                state["submitted"] = i
            b = fd.acquire(i % 2, B_global=self.B * self.world)
            self.it += 1
            F.train_step_batch(self.ctx, self.cfg, self.P, self.G, self.ST, b, self.out, step_t=self.it, opt_t=self.it)
            fd.release(i % 2)
            if not last:                # the upload of the NEXT step's batch overlaps with this step's compute (a K-step run
                fd.submit((i + 1) % 2)  # uploads exactly K batches: none is prefetched behind the last step)
                state["submitted"] = i + 1
            loss_host.copy_(self.out.loss, non_blocking=True)

        for i in range(8):              # both slots seen often enough to be captured
            step(i, last=i == 7)
        torch.cuda.synchronize()
        state["submitted"] = -1
        ms = self.timed(lambda i: step(i, last=i == steps - 1), steps)
        torch.cuda.synchronize()
        h2d, d2h = fd.bytes_per_batch, loss_host.numel() * 4
        fd.close()
        return ms, h2d, d2h

    def cache_leg(self, steps):
        """the same step with the split's features RESIDENT in HBM as fp16 (rau_feat_cache): per step the host sends only the
        question tokens, lengths, labels and image indices (pinned, async); the feature tensor is a device-side gather"""
        torch = self.torch
        from rau_vqa_b200 import feed as F
        n_img = self.NB * self.B
        cache = F.FeatCache(self.ctx, n_img, self.cfg.C, self.cfg.S)
        for k, hb in enumerate(self.host):
            cache.put(k * self.B, hb[0])
        small_h = []
        for k, hb in enumerate(self.host):          # [tokens | lengths | labels | image index] of one batch, pinned
            idx = np.arange(k * self.B, (k + 1) * self.B, dtype=np.float32) + 1.0
            small_h.append(torch.from_numpy(np.concatenate([hb[1].ravel(), hb[2], hb[3], idx])).pin_memory())
        T, B = self.cfg.T, self.B
        small_d = [torch.empty_like(small_h[0], device=self.dev) for _ in range(2)]
        # (nHop > 1 in a tcgen05 mode: the gather stays fp16 and the feature pack reads it directly, rau_batch.feats_f16)
        f16 = self.cfg.nHop > 1 and PREC_NAMES[int(self.ctx.lib.rau_get_precision(self.ctx.h))] != "f32"
        feats_d = [torch.empty(B, self.cfg.C, self.cfg.S, device=self.dev, dtype=torch.float16 if f16 else torch.float32)
                   for _ in range(2)]
        loss_host = torch.empty(self.cfg.nHop + 2, dtype=torch.float32).pin_memory()

        def step(i):
            s = i % 2
            d = small_d[s]
            d.copy_(small_h[i % self.NB], non_blocking=True)
            (cache.gather_f16 if f16 else cache.gather)(d[T * B + 2 * B:], out=feats_d[s])
            self.step((feats_d[s], d[:T * B].view(T, B), d[T * B:T * B + B], d[T * B + B:T * B + 2 * B]))
            loss_host.copy_(self.out.loss, non_blocking=True)

        for i in range(8):
            step(i)
        torch.cuda.synchronize()
        ms = self.timed(step, steps)
        torch.cuda.synchronize()
        cache.close()
        return ms, small_h[0].numel() * 4

    def free(self):
        self.ctx.sync()
        del self.P, self.G, self.ST, self.resident, self.out
        self.torch.cuda.empty_cache()


def sweep(ctx, dev, pk, Cs=(512, 1024, 2048), Bs=(32, 128, 256, 1024), iters=3):
    """BASELINE.json configs[4]: the attention kernels of one answering unit at C x B, forward and backward, each launched
    alone with L2 evicted before every launch; HBM fraction (algorithmic bytes / time / copy peak) and tensor fraction
    (algorithmic flops / time / bf16 burst peak) per point.  Question length only affects the encoder, not these kernels."""
    import torch
    import rau_vqa_b200 as R
    from rau_vqa_b200._ffi import check, ffi
    from rau_vqa_b200.core import fptr
    prec = PREC_NAMES[int(ctx.lib.rau_get_precision(ctx.h))]
    xb = 2 if prec in ("mixed", "f16img", "bf16") else 4        # bytes per element of Xd
    ib = 2 if prec in ("f16img", "bf16") else 4                 # ... of I
    zb = 2 if prec in ("f16img", "bf16") else 4                 # ... of dZ
    yb = 2 if prec in ("mixed", "f16img", "bf16") else 4        # ... of dY
    names = ["pack", "i_embed", "Z", "score", "softmax_sum", "bwd_dp_dz", "dY", "gWa", "gWi"]
    M, A, S = 512, 256, 196
    rows = []
    for C in Cs:
        cfg = R.RauConfig(V=16384, C=C, nHop=1, N=2000)
        gen = torch.Generator(device=dev).manual_seed(5)
        mult = torch.rand(cfg.group_size(2), device=dev, generator=gen) * 0.16 - 0.08
        for B in Bs:
            X = torch.relu(torch.randn(B, C, S, device=dev, generator=gen))
            us = ffi.new("float[9]")
            check(ctx.lib.rau_sweep_attention(ctx.h, cfg.c(), B, fptr(mult), fptr(X), iters, 1, us))
            Rr = B * S
            # algorithmic bytes (what the kernel must move in this precision mode) and flops (2 M N K of the reference graph)
            byt = dict(pack=Rr * C * (4 + xb), i_embed=Rr * (C * xb + M * ib), Z=Rr * (M * ib + A * 4), score=Rr * A * 4,
                       softmax_sum=Rr * M * ib, bwd_dp_dz=Rr * (M * ib + A * 4 + A * zb), dY=Rr * (A * zb + M * ib + M * yb),
                       gWa=Rr * (A * zb + M * ib), gWi=Rr * (M * yb + C * xb))
            flo = dict(i_embed=2.0 * Rr * M * C, Z=2.0 * Rr * A * M, dY=2.0 * Rr * M * A, gWa=2.0 * Rr * A * M, gWi=2.0 * Rr * M * C)
            point = dict(C=C, B=B)
            for k, n in enumerate(names):
                t = us[k] * 1e-6
                e = dict(us=round(us[k], 2), hbm_frac=round(byt[n] / t / 1e9 / pk["hbm"], 3))
                if n in flo:
                    e["tensor_frac"] = round(flo[n] / t / 1e12 / pk["tensor_burst"], 3)
                point[n] = e
            rows.append(point)
            del X
        del mult
        torch.cuda.empty_cache()
    return dict(precision=prec, l2="evicted before every launch (256 MB read sweep: clean lines)", iters=iters,
                peaks=dict(hbm_gbs=pk["hbm"], bf16_tflops=pk["tensor_burst"], source=pk["src"]),
                note="hbm_frac = algorithmic bytes / time / copy peak; tensor_frac = 2MNK / time / bf16 burst peak; question length "
                     "8-26 only changes the encoder, not these kernels", points=rows)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ours_full", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--precision", default="default", choices=["default", "f32", "bf16", "bf16x3", "mixed", "f16img"])
    ap.add_argument("--cpu-sample", type=int, default=0, help=f"rows of the CPU sample (0 = {CPU_SAMPLE_B}, in both arms)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other configurations / sweep / fp16 feed legs")
    ap.add_argument("--sweep", action="store_true", help="run only the attention-kernel sweep and print it")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_main(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    import rau_vqa_b200 as R
    from rau_vqa_b200 import core
    from rau_vqa_b200 import feed as F

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU path (use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = R.Context(local, seed=123)
    if args.precision != "default":
        ctx.set_precision({v: k for k, v in PREC_NAMES.items()}[args.precision])
    prec = PREC_NAMES[int(ctx.lib.rau_get_precision(ctx.h))]
    if world > 1:
        ids = [core.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], rank, world)
    numa = bind_to_gpu_numa(local) if world > 1 else None
    dev = torch.device("cuda", local)
    pk = peaks()

    if args.sweep:
        res = sweep(ctx, dev, pk)
        if rank == 0:
            emit(dict(metric="attention-kernel sweep", **res))
        ctx.close()
        return

    run = Runner(ctx, args.workload, args.batch, world, rank, dev, dist)
    B, cfg, NB = run.B, run.cfg, run.NB
    run.prime(args.warmup)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = run.resident_leg(args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # The feed's staging format: in the mixed modes the features enter the tensor pipe as fp16 and the power-of-two dropout
    # scale commutes with the rounding, so fp16 staging (half the PCIe bytes) changes no bit of what the products read
    # (tests/test_gpu_feed.py) and is what `e2e` uses; the float32 staging of the reference's upload is in extra.
    # ... and the all-hops feature pack reads the uploaded fp16 buffer as it is (RAU_FEED_F16_DIRECT: no widening pass on
    # the copy stream; needs nHop > 1, the one-hop configuration keeps the widened form).  RAU_BENCH_FEED=f16 | f32 overrides.
    main_fmt = (F.FEED_F16_DIRECT if run.cfg.nHop > 1 else F.FEED_F16) if prec in ("mixed", "f16img") else F.FEED_F32
    main_fmt = dict(f16=F.FEED_F16, f32=F.FEED_F32, direct=F.FEED_F16_DIRECT).get(os.environ.get("RAU_BENCH_FEED", ""), main_fmt)
    ms_e2e, h2d, d2h = run.e2e_leg(args.steps, main_fmt)
    torch.cuda.synchronize()
    assert torch.isfinite(run.out.loss).all().item(), "loss is not finite"
    value = B * world * args.steps / (ms * 1e-3)
    e2e_v = B * world * args.steps / (ms_e2e * 1e-3)

    extra = {}
    if not args.no_extra and main_fmt != F.FEED_F32:
        # the same step fed through float32 staging, byte for byte what the reference uploads (F:452-456)
        ms32, h2d32, _ = run.e2e_leg(args.steps, F.FEED_F32)
        extra["e2e_f32_feed"] = dict(value=B * world * args.steps / (ms32 * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d32,
                                     ms_per_step=ms32 / args.steps)
    if not args.no_extra:
        msc, h2dc = run.cache_leg(args.steps)
        extra["e2e_feature_cache"] = dict(value=B * world * args.steps / (msc * 1e-3), unit=UNIT, h2d_bytes_per_step=h2dc,
                                          ms_per_step=msc / args.steps,
                                          note="features resident in HBM as fp16 (rau_feat_cache_*), gathered by image index")
    if not args.no_extra and world > 1 and args.workload == "ours_full" and not args.batch and B % world == 0:
        # strong scaling: the SAME global batch of 256 split over the ranks (SURVEY.md 8d/8e asked for it "for honesty")
        run.free()
        rs = Runner(ctx, args.workload, B // world, world, rank, dev, dist)
        rs.prime(args.warmup)
        ms_s, _ = rs.resident_leg(args.steps)
        extra["strong_scaling"] = dict(value=B * args.steps / (ms_s * 1e-3), unit=UNIT, global_batch=B, batch_per_gpu=B // world,
                                       ms_per_step=ms_s / args.steps, scaling="strong")
        rs.free()
        run = None
    if rank == 0:
        # dominant kernel: the i_embed projection I = tanh(Wi drop(X) + bi) (F:240), 2*M*C*196 flop per image
        if run is None:
            run = Runner(ctx, args.workload, args.batch, 1, rank, dev, None)
        ms_k = ffi_time_iembed(ctx, cfg, B, run.P[2], run.resident[0][0])
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
        roof = dict(bound="tensor", kernel="rows_gemm_kernel<EPI_TANH>: i_embed product I = tanh(Wi drop(X) + bi) (F:238-242), "
                                           "precision mode " + prec + (" (one fp16 pass)" if prec in ("mixed", "f16img") else ""),
                    achieved=ach, peak=pk["tensor_burst"], unit="TFLOP/s", frac=ach / pk["tensor_burst"], traffic=traffic,
                    peak_source=pk["src"] + " bf16 burst (kernel timed alone, 20 back-to-back launches, CUDA events)",
                    ms_per_launch=ms_k, flop_per_launch=flops, algorithmic="2*196*B*M*C flop per launch (SURVEY 8d)",
                    mma_passes=executed, executed_frac=executed * ach / pk["tensor_burst"])
        # the same launch against the kernel's OTHER roof: packed features in (one fp16 plane in the mixed modes, bf16 hi / hi+lo
        # otherwise) + I out (fp16 / bf16 hi / hi+lo) per launch over the measured copy bandwidth.  In the default mode the
        # arithmetic intensity (171 flop/B) is below the ridge of the two measured peaks (253 flop/B): HBM caps `frac` at ~0.67.
        xb_ = 2 if prec in ("mixed", "f16img", "bf16") else 4
        ib_ = 2 if prec in ("f16img", "bf16") else 4
        hbm_bytes = 196.0 * B * (cfg.C * xb_ + cfg.M * ib_)
        roof["hbm"] = dict(bytes_per_launch=hbm_bytes, achieved_gbs=hbm_bytes / (ms_k * 1e-3) / 1e9, peak_gbs=pk["hbm"],
                           frac=hbm_bytes / (ms_k * 1e-3) / 1e9 / pk["hbm"])
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu, _ = cpu_baseline_dict(args.workload, 2, 1, args.cpu_sample)
        if world == 1 and not args.no_extra and args.workload == "ours_full" and not args.batch:
            # the other BASELINE configurations, short runs (same legs, same clocks): configs[1], configs[3], configs[0]
            run.free()
            for w in ("ours_ms", "ours_resnet", "ours_ss"):
                r2 = Runner(ctx, w, 0, 1, 0, dev, None)
                r2.prime(3)
                cs2 = ClockSampler(local)
                cs2.start()
                m2, l2 = r2.resident_leg(args.steps)
                extra[w] = dict(value=r2.B * args.steps / (m2 * 1e-3), unit=UNIT, ms_per_step=m2 / args.steps,
                                launches_per_step=l2 / args.steps, clocks=cs2.stop(), config=make_config(w, 1))
                r2.free()
            extra["sweep"] = sweep(ctx, dev, pk)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=PREC_DTYPE[prec],
                    precision_mode=prec, data="synthetic",
                    config=make_config(args.workload, world, B),
                    l2=f"{NB} rotating batches ({NB * B * cfg.C * 196 * 4 / 1e6:.0f} MB of features) + >1 GB of saved activations per "
                       "step exceed the 126 MB L2",
                    e2e=dict(value=e2e_v, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=ms_e2e / args.steps,
                             feed="rau_feed_* (pinned " + {F.FEED_F32: "float32", F.FEED_F16: "float16",
                                                            F.FEED_F16_DIRECT: "float16, read directly by the feature pack:"}[main_fmt] +
                                  " staging, depth 2, copy stream)"),
                    gpu_launches=int(launches), launches_per_step=launches / args.steps, clocks=clocks, roofline=roof,
                    cpu_baseline=cpu, extra=extra)
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
