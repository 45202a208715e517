"""Row-slab sharded CG on real GPUs: launches tests/multi_gpu_cg_check.py under torchrun with 2
ranks when the box has at least 2 GPUs (gpurun --gpus 2), else skipped."""
import json
import os
import signal
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


@pytest.mark.gpu
@pytest.mark.timeout(600)
@pytest.mark.parametrize("kind", ["picflip", "sl"])
@pytest.mark.parametrize("backend", ["gloo", "nccl"])
def test_particle_slabs_over_torch_distributed(built, backend, kind):
    """DistSlabs (fluid_simulation_b200/sharding.py): the slab-partitioned PIC/FLIP step over
    torch.distributed.  gloo: two ranks share GPU 0 (checks the distributed call sequence on a
    one-GPU box); nccl: one rank per GPU, needs 2 GPUs."""
    import torch
    if backend == "nccl":
        if torch.cuda.device_count() < 2:
            pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533" if backend == "gloo" else "29534",
           os.path.join(ROOT, "tests", "multi_gpu_slab_check.py"), "--backend", backend, "--kind", kind]
    # own process group, killed as a whole on a time-out: a hung rank must not keep the GPU
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT,
                         start_new_session=True)
    try:
        so, se = p.communicate(timeout=180)
    except subprocess.TimeoutExpired:
        os.killpg(p.pid, signal.SIGKILL)
        so, se = p.communicate()
        pytest.fail("slab check timed out:\n" + so[-2000:] + se[-2000:])
    assert p.returncode == 0, so[-3000:] + se[-3000:]
    out = json.loads([l for l in so.splitlines() if l.startswith("{")][-1])
    assert out["ok"] and out["all_particles"] > out["own_particles"] > 0, out
