"""Host-side plumbing of the row-slab sharded pressure solve (one process per GPU).

The data path needs no collective: slab halos and dot products move through peer memory inside
the CG kernels (include/fsb.h, "multi-GPU").  What the host has to do once per context is hand
every rank the CUDA IPC handles of all ranks; this module does that with torch.distributed
(NCCL on GPUs, gloo in the CPU tests).  torch is plumbing only.
"""
import numpy as np

from .capi import SHARD_BLOB_BYTES


def slab_rows(ny, world, rank):
    """Rows [lo, hi) rank `rank` iterates on: the even split fsb_shard_connect uses."""
    return ny * rank // world, ny * (rank + 1) // world


def gather_blobs(blob, dist=None, device=None):
    """All-gather one fixed-size uint8 blob per rank; returns a (world, len) numpy array in rank
    order.  `device`: where the staging tensors live (cuda for NCCL, cpu for gloo)."""
    import torch
    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(blob, dtype=np.uint8))
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()


def connect(sim, dist=None, device=None):
    """Export this rank's handles, exchange them, connect the slabs.  Collective: every rank of
    the default process group must call it.  Returns this rank's (row_lo, row_hi)."""
    if dist is None:
        import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return 0, sim.ny
    blobs = gather_blobs(sim.shard_export(), dist, device)
    assert blobs.shape == (world, SHARD_BLOB_BYTES)
    sim.shard_connect(rank, world, blobs)
    dist.barrier()  # nobody starts a solve before every rank has opened its peers
    lo, hi = sim.shard_rows()
    assert (lo, hi) == slab_rows(sim.ny, world, rank)
    return lo, hi


# ----------------------------------------------------------------------------------------------
# Particle slabs (include/fsb.h "particle slabs"): every rank keeps the full grids but only the
# particles of its row slab.  The device phases are fsb_slab_step_a / _b / _c; this module moves
# the data between them.  Two transports with the same call sequence:
#   * LocalSlabs  -- all ranks are Sim objects in ONE process (tests: the whole protocol on one GPU)
#   * DistSlabs   -- one Sim per process, torch.distributed (gloo / NCCL) moves numpy buffers
# Because the cell sort's in-cell order goes by global particle id, a slab-partitioned run gives
# the same bits as the single-GPU run (tests/test_gpu_parity.py::test_particle_slabs_*).
from .capi import ROWS_LABELS, STEP_SL, U_FRONT, V_FRONT  # noqa: E402


class LocalSlabs:
    """`world` ranks in one process; exchanges are plain numpy hand-overs."""

    def __init__(self, sims):
        self.sims = list(sims)
        self.world = len(self.sims)
        self.rows = [s.slab_configure(q, self.world) for q, s in enumerate(self.sims)]

    def distribute(self):
        """Every rank was given the SAME full particle set (fsb_set_particles / fsb_emit_source):
        keep the own slab only."""
        for s in self.sims:
            s.slab_sort_out(self.world)
            s.slab_keep_own()

    def _ghosts(self):
        out = [[s.slab_boundary(0), s.slab_boundary(1)] for s in self.sims]
        for q, s in enumerate(self.sims):
            if q > 0:
                s.slab_add(*out[q - 1][1])  # the lower neighbour's last row
            if q + 1 < self.world:
                s.slab_add(*out[q + 1][0])  # the upper neighbour's first row

    def _rows(self, labels_only=False):
        for which in ((ROWS_LABELS,) if labels_only else (ROWS_LABELS, U_FRONT, V_FRONT)):
            slabs = [s.get_rows(which, lo, hi) for s, (lo, hi) in zip(self.sims, self.rows)]
            for s in self.sims:
                for (lo, hi), a in zip(self.rows, slabs):
                    s.set_rows(which, lo, hi, a)

    def _migrate(self):
        counts = [s.slab_sort_out(self.world) for s in self.sims]
        moving = [[self.sims[q].slab_take(d, counts[q][d]) if d != q and counts[q][d] else None
                   for d in range(self.world)] for q in range(self.world)]
        for s in self.sims:
            s.slab_keep_own()
        for q in range(self.world):
            for d in range(self.world):
                if moving[q][d] is not None:
                    self.sims[d].slab_add(*moving[q][d])
        return sum(counts[q][d] for q in range(self.world) for d in range(self.world) if d != q)

    def step(self, kind, dt):
        sl = kind == STEP_SL  # markers only: no ghost rows, label rows only (include/fsb.h)
        if not sl:
            self._ghosts()
        for s in self.sims:
            s.slab_step_a(kind)
        self._rows(labels_only=sl)
        for s in self.sims:
            s.slab_step_b(kind, dt)
        for s in self.sims:
            s.slab_step_c(kind, dt)
        return self._migrate()

    def particles(self):
        """The whole set in the caller's (global id) order."""
        parts, ids = zip(*[s.slab_get() for s in self.sims])
        parts, ids = np.concatenate(parts), np.concatenate(ids)
        out = np.empty_like(parts)
        out[ids] = parts
        assert np.unique(ids).size == ids.size == out.shape[0]
        return out


class DistSlabs:
    """One rank of a torch.distributed job.  Same call sequence as LocalSlabs.  All buffers are torch
    tensors on `device` -- CUDA tensors with the NCCL backend, CPU tensors with gloo -- and the library
    reads / writes them through their raw pointers (fsb_slab_add / _take / fsb_get_rows / fsb_set_rows
    take host or device memory): with NCCL nothing is staged through the host.  Point-to-point
    transfers are issued as ONE batch_isend_irecv group (an ungrouped send / recv pair in opposite
    directions between two ranks deadlocks under NCCL); the library's copies run on the context's own
    stream and have completed when its calls return, the collectives are waited for before the library
    touches a buffer."""

    def __init__(self, sim, dist=None, device=None):
        import torch
        if dist is None:
            import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.sim = sim
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.lo, self.hi = sim.slab_configure(self.rank, self.world)
        self.rows = [slab_rows(sim.ny, self.world, q) for q in range(self.world)]
        self.max_rows = max(hi - lo for lo, hi in self.rows)

    def _sync(self):
        if self.device.type == "cuda":
            self.torch.cuda.synchronize(self.device)

    def _pair(self, n):
        torch = self.torch
        return (torch.empty((max(n, 1), 4), dtype=torch.float32, device=self.device),
                torch.empty(max(n, 1), dtype=torch.int32, device=self.device))

    def _exchange(self, outgoing):
        """outgoing[d] = (parts tensor, ids tensor, n) or None; returns [(parts, ids, n)] received."""
        torch, dist = self.torch, self.dist
        n_out = torch.tensor([0 if o is None else o[2] for o in outgoing], dtype=torch.int64, device=self.device)
        table = [torch.empty_like(n_out) for _ in range(self.world)]
        dist.all_gather(table, n_out)  # table[q][d] = what q sends to d
        n_in = [int(v) for v in torch.stack(table)[:, self.rank].cpu().tolist()]
        ops, recv = [], []
        for q in range(self.world):
            if q != self.rank and outgoing[q] is not None and outgoing[q][2]:
                p, i, n = outgoing[q]
                ops.append(dist.P2POp(dist.isend, p[:n], q))
                ops.append(dist.P2POp(dist.isend, i[:n], q))
        for q in range(self.world):
            if q != self.rank and n_in[q]:
                p, i = self._pair(n_in[q])
                ops.append(dist.P2POp(dist.irecv, p[:n_in[q]], q))
                ops.append(dist.P2POp(dist.irecv, i[:n_in[q]], q))
                recv.append((p, i, n_in[q]))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
            self._sync()
        return recv

    def distribute(self):
        self.sim.slab_sort_out(self.world)
        self.sim.slab_keep_own()

    def _boundary(self, side):
        n = self.sim.slab_boundary_count(side)
        p, i = self._pair(n)
        if n:
            self.sim.slab_boundary_take_ptr(p.data_ptr(), i.data_ptr())
        return p, i, n

    def _gather_rows(self, labels_only=False):
        """Every rank's label / u / v rows to every rank.  Slabs differ by a row when ny is not a multiple
        of the world size: the buffers are padded to the tallest slab (all_gather wants equal shapes)."""
        torch, dist, s = self.torch, self.dist, self.sim
        for which in ((ROWS_LABELS,) if labels_only else (ROWS_LABELS, U_FRONT, V_FRONT)):
            dtype = torch.uint8 if which == ROWS_LABELS else torch.float32
            mine = torch.zeros((self.max_rows, s.nx), dtype=dtype, device=self.device)
            s.get_rows_ptr(which, self.lo, self.hi, mine.data_ptr())
            slabs = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(slabs, mine)
            self._sync()
            for q, (lo, hi) in enumerate(self.rows):
                if q != self.rank:
                    s.set_rows_ptr(which, lo, hi, slabs[q].data_ptr())

    def step(self, kind, dt):
        s = self.sim
        sl = kind == STEP_SL  # markers only: no ghost rows, label rows only (include/fsb.h)
        if not sl:
            out = [None] * self.world
            if self.rank > 0:
                out[self.rank - 1] = self._boundary(0)
            if self.rank + 1 < self.world:
                out[self.rank + 1] = self._boundary(1)
            for p, i, n in self._exchange(out):
                s.slab_add_ptr(p.data_ptr(), i.data_ptr(), n)
        s.slab_step_a(kind)
        self._gather_rows(labels_only=sl)
        s.slab_step_b(kind, dt)
        s.slab_step_c(kind, dt)
        counts = s.slab_sort_out(self.world)
        out = [None] * self.world
        for d in range(self.world):
            if d != self.rank and counts[d]:
                p, i = self._pair(counts[d])
                s.slab_take_ptr(d, p.data_ptr(), i.data_ptr())
                out[d] = (p, i, counts[d])
        s.slab_keep_own()
        for p, i, n in self._exchange(out):
            s.slab_add_ptr(p.data_ptr(), i.data_ptr(), n)
        return sum(c for d, c in enumerate(counts) if d != self.rank)

    def particles(self):
        """The whole set in global-id order, gathered on every rank (tests; small scenes)."""
        torch, dist = self.torch, self.dist
        n_own = self.sim.num_particles()
        n = torch.tensor([n_own], dtype=torch.int64, device=self.device)
        ns = [torch.empty_like(n) for _ in range(self.world)]
        dist.all_gather(ns, n)
        ns = [int(v.item()) for v in ns]
        cap = max(ns) if ns else 0
        pp, ii = self._pair(cap)
        pp.zero_(); ii.zero_()
        if n_own:
            self.sim.slab_get_ptr(pp.data_ptr(), ii.data_ptr())
        gp = [torch.empty_like(pp) for _ in range(self.world)]
        gi = [torch.empty_like(ii) for _ in range(self.world)]
        dist.all_gather(gp, pp)
        dist.all_gather(gi, ii)
        self._sync()
        out = np.empty((sum(ns), 4), dtype=np.float32)
        for q in range(self.world):
            out[gi[q][:ns[q]].cpu().numpy()] = gp[q][:ns[q]].cpu().numpy()
        return out
