"""Host-side logic of the particle slabs on CPU (no GPU, no libfsb compute calls): the exchange
protocol of fluid_simulation_b200/sharding.py -- ghost rows, row all-gather, migration, the global
gather -- driven by a stand-in for the device context that implements the fsb_slab_* contract in
numpy (ownership = row of the position, a trivial "physics": every particle drifts by its velocity).
LocalSlabs in one process, DistSlabs over torch.distributed / gloo with world_size 2 and 3.
The device phases themselves are checked on the GPU (tests/test_gpu_parity.py,
tests/test_multi_gpu.py)."""
import os
import socket

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NY, NX = 25, 8  # 25 rows: the slabs of 2 / 3 / 4 ranks differ in height (padded row all-gather)


class FakeSim:
    """The slab contract of include/fsb.h with numpy arrays instead of HBM."""

    def __init__(self, parts, ids):
        self.nx, self.ny = NX, NY
        self.dy = 1.0 / NY
        self.parts, self.ids = parts.copy(), ids.copy()
        self.labels = np.ones((NY, NX), dtype=np.uint8)
        self.grids = {0: np.zeros((NY, NX), dtype=np.float32), 1: np.zeros((NY, NX), dtype=np.float32)}
        self.groups = None
        self.log = []

    def _row(self, parts):
        return np.clip((parts[:, 1] / np.float32(self.dy)).astype(np.int64), 0, NY - 1)

    def _owner(self, rows):
        return np.searchsorted(np.array([NY * (q + 1) // self.world for q in range(self.world)]), rows, side="right")

    def slab_configure(self, rank, world):
        self.rank, self.world = rank, world
        self.lo, self.hi = NY * rank // world, NY * (rank + 1) // world
        return self.lo, self.hi

    def num_particles(self):
        return self.ids.shape[0]

    def slab_add(self, parts, ids):
        self.parts = np.concatenate([self.parts, np.asarray(parts, dtype=np.float32).reshape(-1, 4)])
        self.ids = np.concatenate([self.ids, np.asarray(ids, dtype=np.int32)])

    def slab_sort_out(self, world):
        live = self.ids >= 0
        self.parts, self.ids = self.parts[live], self.ids[live]
        own = self._owner(self._row(self.parts))
        self.groups = [np.nonzero(own == q)[0] for q in range(world)]
        return [int(g.size) for g in self.groups]

    def slab_take(self, dest, n):
        g = self.groups[dest]
        assert g.size == n
        return self.parts[g].copy(), self.ids[g].copy()

    def slab_keep_own(self):
        g = self.groups[self.rank]
        self.parts, self.ids, self.groups = self.parts[g], self.ids[g], None

    def slab_boundary(self, side):
        rows = self._row(self.parts)
        sel = (rows == (self.lo if side == 0 else self.hi - 1)) & (self.ids >= 0)
        return self.parts[sel].copy(), self.ids[sel].copy()

    def slab_get(self):
        return self.parts.copy(), self.ids.copy()

    # the raw-pointer forms DistSlabs uses (host memory here; device memory with NCCL on a GPU)
    @staticmethod
    def _view(ptr, shape, dtype):
        import ctypes
        n = int(np.prod(shape))
        if n == 0:
            return np.empty(shape, dtype=dtype)
        buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def slab_add_ptr(self, pp, ip, n):
        if n:
            self.slab_add(self._view(pp, (n, 4), np.float32).copy(), self._view(ip, (n,), np.int32).copy())

    def slab_take_ptr(self, dest, pp, ip):
        parts, ids = self.slab_take(dest, self.groups[dest].size)
        self._view(pp, parts.shape, np.float32)[:] = parts
        self._view(ip, ids.shape, np.int32)[:] = ids

    def slab_boundary_count(self, side):
        self._sel = self.slab_boundary(side)
        return self._sel[1].shape[0]

    def slab_boundary_take_ptr(self, pp, ip):
        parts, ids = self._sel
        self._view(pp, parts.shape, np.float32)[:] = parts
        self._view(ip, ids.shape, np.int32)[:] = ids

    def slab_get_ptr(self, pp, ip):
        self._view(pp, self.parts.shape, np.float32)[:] = self.parts
        self._view(ip, self.ids.shape, np.int32)[:] = self.ids

    def get_rows_ptr(self, which, lo, hi, ptr):
        a = self.get_rows(which, lo, hi)
        self._view(ptr, a.shape, a.dtype)[:] = a

    def set_rows_ptr(self, which, lo, hi, ptr):
        dtype = np.uint8 if which == 8 else np.float32
        self.set_rows(which, lo, hi, self._view(ptr, (hi - lo, self.nx), dtype).copy())

    def get_rows(self, which, lo, hi):
        return (self.labels if which == 8 else self.grids[which])[lo:hi].copy()

    def set_rows(self, which, lo, hi, a):
        (self.labels if which == 8 else self.grids[which])[lo:hi] = a

    def slab_step_a(self, kind):
        # "classification" and "P2G" of the own rows: particle count per cell row / sum of ids
        rows = self._row(self.parts)
        self.ghost_rows_seen = sorted(set(int(r) for r in rows if not (self.lo <= r < self.hi)))
        for j in range(self.lo, self.hi):
            m = rows == j
            self.labels[j, :] = 0 if m.any() else 1
            if kind != 0:  # a semi-Lagrangian step (kind 0) has no P2G: labels only
                self.grids[0][j, :] = np.float32(m.sum())
                self.grids[1][j, :] = np.float32(self.ids[m].sum() % 1000)
        self.log.append("a")

    def slab_step_b(self, kind, dt):
        self.log.append("b")

    def slab_step_c(self, kind, dt):
        rows = self._row(self.parts)
        ghost = (rows < self.lo) | (rows >= self.hi)
        self.ids[ghost] = -1
        live = ~ghost
        self.parts[live, 0] += self.parts[live, 2] * np.float32(dt)
        self.parts[live, 1] += self.parts[live, 3] * np.float32(dt)
        self.log.append("c")


def make_set(seed=7, n=300):
    rng = np.random.default_rng(seed)
    p = np.empty((n, 4), dtype=np.float32)
    p[:, 0] = rng.uniform(0.05, 0.95, n)
    p[:, 1] = rng.uniform(0.05, 0.95, n)
    p[:, 2] = rng.uniform(-1, 1, n)
    p[:, 3] = rng.uniform(-3, 3, n)  # up to ~0.7 cell rows per step of dt = 0.01: slabs are crossed
    return p, np.arange(n, dtype=np.int32)


def reference_run(steps, dt):
    p, ids = make_set()
    for _ in range(steps):
        p[:, 0] += p[:, 2] * np.float32(dt)
        p[:, 1] += p[:, 3] * np.float32(dt)
    return p


def expected_rows(parts):
    rows = np.clip((parts[:, 1] / np.float32(1.0 / NY)).astype(np.int64), 0, NY - 1)
    return np.bincount(rows, minlength=NY).astype(np.float32)


@pytest.mark.parametrize("kind", [3, 0])
@pytest.mark.parametrize("world", [2, 3, 4])
def test_local_slabs_protocol(world, kind):
    import sys
    sys.path.insert(0, ROOT)
    from fluid_simulation_b200 import sharding
    p0, ids = make_set()
    sims = [FakeSim(p0, ids) for _ in range(world)]
    slabs = sharding.LocalSlabs(sims)
    slabs.distribute()
    assert sum(s.num_particles() for s in sims) == ids.size
    cur = p0.copy()
    moved = 0
    for step in range(8):
        before = expected_rows(cur)
        moved += slabs.step(kind, 0.01)
        cur[:, 0] += cur[:, 2] * np.float32(0.01)
        cur[:, 1] += cur[:, 3] * np.float32(0.01)
        for q, s in enumerate(sims):
            assert s.log[-3:] == ["a", "b", "c"]
            # after the row all-gather every rank holds the labels of the whole grid
            assert np.array_equal(s.labels[:, 0], (before == 0).astype(np.uint8)), (step, q)
            if kind == 0:
                # semi-Lagrangian: no ghost rows travel, and the velocity rows are not exchanged
                assert s.ghost_rows_seen == [] and not s.grids[0].any()
                continue
            # every rank saw exactly its neighbours' boundary rows as ghosts ...
            assert set(s.ghost_rows_seen) <= {s.lo - 1, s.hi}
            # ... and holds the whole "grid": per-row particle counts
            assert np.array_equal(s.grids[0][:, 0], before), (step, q)
            # only own, live particles remain
        for s in sims:
            rows = s._row(s.parts)
            assert ((rows >= s.lo) & (rows < s.hi)).all() and (s.ids >= 0).all()
    assert moved > 0
    assert np.array_equal(slabs.particles(), reference_run(8, 0.01))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, kind=3):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from fluid_simulation_b200 import sharding
    import test_slabs_host as me
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    p0, ids = me.make_set()
    sim = me.FakeSim(p0, ids)
    slabs = sharding.DistSlabs(sim, dist, None)
    slabs.distribute()
    cur, ok, moved = p0.copy(), True, 0
    for step in range(8):
        before = me.expected_rows(cur)
        moved += slabs.step(kind, 0.01)
        cur[:, 0] += cur[:, 2] * np.float32(0.01)
        cur[:, 1] += cur[:, 3] * np.float32(0.01)
        ok = ok and np.array_equal(sim.labels[:, 0], (before == 0).astype(np.uint8))
        if kind == 0:
            ok = ok and sim.ghost_rows_seen == [] and not sim.grids[0].any()
        else:
            ok = ok and np.array_equal(sim.grids[0][:, 0], before)
        rows = sim._row(sim.parts)
        ok = ok and bool(((rows >= sim.lo) & (rows < sim.hi)).all()) and bool((sim.ids >= 0).all())
    ok = ok and np.array_equal(slabs.particles(), me.reference_run(8, 0.01))
    q.put((rank, bool(ok), int(moved)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world,kind", [(2, 3), (3, 3), (2, 0)])
def test_dist_slabs_protocol_gloo(world, kind):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, kind)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in procs)
    for p in procs:
        p.join(timeout=30)
    assert [r[:2] for r in res] == [(r, True) for r in range(world)]
    assert sum(r[2] for r in res) > 0  # particles did migrate
