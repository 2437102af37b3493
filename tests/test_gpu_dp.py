"""Data parallelism on real GPUs: two ranks, real NCCL, through rau_train_step.  Needs two devices -- the
single-GPU test box skips it; `gpurun --gpus 2` runs it, and its log is kept under profiles/."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_match_one_process_on_the_whole_batch():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under `gpurun --gpus 2`)")
    env = dict(os.environ)
    env.pop("RAU_TC_MIN_WORK", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "tools", "dp_check.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "dp_check_2gpu.log"), "w") as f:
        f.write(r.stdout + "\n---- stderr ----\n" + r.stderr[-4000:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert "DP CHECK OK" in r.stdout, r.stdout
    assert r.stdout.count("replicas bit-identical across ranks: True") == 3, r.stdout
