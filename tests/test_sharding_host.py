"""Host-side logic of the multi-GPU path on CPU: the slab partition and the handle exchange over
torch.distributed with the gloo backend, world_size 2 (no GPU, no compute calls)."""
import os
import socket

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_rows_partition(capi):
    from fluid_simulation_b200.sharding import slab_rows
    for ny in (64, 1000, 4096, 8192, 16384, 37):
        for world in (1, 2, 3, 4, 8):
            rows = [slab_rows(ny, world, r) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == ny
            for a, b in zip(rows, rows[1:]):
                assert a[1] == b[0]  # contiguous, no overlap
            sizes = [hi - lo for lo, hi in rows]
            assert max(sizes) - min(sizes) <= 1  # balanced


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from fluid_simulation_b200 import sharding
    from fluid_simulation_b200.capi import SHARD_BLOB_BYTES
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    blob = np.full(SHARD_BLOB_BYTES, rank + 1, dtype=np.uint8)
    blob[:4] = [0x31, 0x42, 0x53, 0x46]
    got = sharding.gather_blobs(blob, dist)
    ok = got.shape == (world, SHARD_BLOB_BYTES) and all((got[r, 4:] == r + 1).all() for r in range(world))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_handle_exchange_gloo_world2(capi):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=90) for _ in procs)
    for p in procs:
        p.join(timeout=30)
    assert res == [(0, True), (1, True)]


def test_shard_calls_fail_without_gpu(capi):
    """No device: contexts cannot be created, so there is nothing to shard (and no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        capi.Sim(64, 64).shard_export()
