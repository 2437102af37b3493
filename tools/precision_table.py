"""python tools/precision_table.py: per-tensor error of every precision mode against the float64 oracle on two BASELINE
configurations (Ours_Full B 256 nHop 8; Ours_SS B 8 nHop 1), through the graph-replayed training step with drawn masks
(tests/test_gpu_baseline_configs.py::_graph_step_vs_oracle).  Writes gpurun_out/precision_table.json; modes whose error
exceeds the 1e-3 bar are recorded, not hidden."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import rau_oracle as O
import test_gpu_baseline_configs as T

cases = [("ours_full_b256", O.RauConfig(V=16384, C=512, nHop=8, N=2000), 256, 2301),
         ("ours_ss_b8", O.RauConfig(V=16384, C=512, nHop=1, N=2000), 8, 2341)]
modes = [("bf16x3", 2), ("mixed", 3), ("f16img", 4), ("bf16", 1)]
table = []
for name, cfg, B, seed in cases:
    for mname, mode in modes:
        try:
            T._graph_step_vs_oracle(name, cfg, B, seed=seed, precision=mode)
            ok = True
        except AssertionError:
            ok = False
        d = json.load(open(os.path.join(ROOT, "gpurun_out", f"parity_{name}_{mname}.json")))
        pt = {k: v for k, v in d["per_tensor"].items() if isinstance(v, float)}
        worst = max(pt.items(), key=lambda kv: kv[1])
        row = dict(case=name, precision=mname, within_1e3=ok, worst_forward=d["worst_forward"], worst_gradient=worst[1],
                   worst_gradient_tensor=worst[0], per_tensor=pt)
        table.append(row)
        print(name, mname, "ok" if ok else "EXCEEDS", f"fwd {d['worst_forward']:.1e}", f"grad {worst[1]:.1e} ({worst[0]})", flush=True)
json.dump(table, open(os.path.join(ROOT, "gpurun_out", "precision_table.json"), "w"), indent=1)
