"""Row-slab sharded CG on real GPUs: launches tests/multi_gpu_cg_check.py under torchrun with 2
ranks when the box has at least 2 GPUs (gpurun --gpus 2), else skipped."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_sharded_cg_two_ranks(built):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29531",
           os.path.join(ROOT, "tests", "multi_gpu_cg_check.py"), "--grid", "512"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=840, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["ok"] and out["ranks_identical"], out
