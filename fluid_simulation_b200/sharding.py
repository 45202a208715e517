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
from .capi import ROWS_LABELS, U_FRONT, V_FRONT  # noqa: E402


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

    def _rows(self):
        for which in (ROWS_LABELS, U_FRONT, V_FRONT):
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
        self._ghosts()
        for s in self.sims:
            s.slab_step_a(kind)
        self._rows()
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
    """One rank of a torch.distributed job (gloo on CPU tensors, or NCCL with device="cuda").
    Same call sequence as LocalSlabs; buffers travel as torch tensors staged through the host.
    Round-1 status: the gloo transport is verified on a B200 (two ranks sharing one GPU,
    tests/test_multi_gpu.py); the NCCL transport's first version deadlocked in ungrouped
    isend / recv pairs, was rewritten with batch_isend_irecv and has NOT been re-run (the round's GPU
    budget was spent): its test is opt-in (FSB_TEST_NCCL_SLABS=1)."""

    def __init__(self, sim, dist=None, device=None):
        import torch
        if dist is None:
            import torch.distributed as dist
        self.torch, self.dist, self.device = torch, dist, device
        self.sim = sim
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.lo, self.hi = sim.slab_configure(self.rank, self.world)
        self.rows = [slab_rows(sim.ny, self.world, q) for q in range(self.world)]

    def _t(self, a):
        t = self.torch.from_numpy(np.ascontiguousarray(a))
        return t.to(self.device) if self.device is not None else t

    def _exchange(self, outgoing):
        """outgoing[d] = (parts, ids) or None; returns the list of (parts, ids) received."""
        torch, dist = self.torch, self.dist
        n_out = torch.tensor([0 if o is None else o[1].shape[0] for o in outgoing], dtype=torch.int64)
        n_out = n_out.to(self.device) if self.device is not None else n_out
        table = [torch.empty_like(n_out) for _ in range(self.world)]
        dist.all_gather(table, n_out)  # table[q][d] = what q sends to d (gloo has no all-to-all)
        n_in = [int(table[q][self.rank].item()) for q in range(self.world)]
        recv = []
        if dist.get_backend() == "nccl":
            # NCCL point-to-point calls must be issued as ONE group: an ungrouped send / recv pair in
            # opposite directions between two ranks deadlocks (each send kernel waits for the peer's
            # recv, which is queued behind the peer's own send)
            ops, bufs = [], []
            for q in range(self.world):
                if q != self.rank and outgoing[q] is not None and outgoing[q][1].shape[0]:
                    ops.append(dist.P2POp(dist.isend, self._t(outgoing[q][0]), q))
                    ops.append(dist.P2POp(dist.isend, self._t(outgoing[q][1]), q))
            for q in range(self.world):
                if q != self.rank and n_in[q]:
                    p = torch.empty((n_in[q], 4), dtype=torch.float32, device=self.device)
                    i = torch.empty(n_in[q], dtype=torch.int32, device=self.device)
                    ops.append(dist.P2POp(dist.irecv, p, q))
                    ops.append(dist.P2POp(dist.irecv, i, q))
                    bufs.append((p, i))
            if ops:
                for r in dist.batch_isend_irecv(ops):
                    r.wait()
                torch.cuda.synchronize()
            return [(p.cpu().numpy(), i.cpu().numpy()) for p, i in bufs]
        reqs = []
        for q in range(self.world):
            if q != self.rank and outgoing[q] is not None and outgoing[q][1].shape[0]:
                reqs.append(dist.isend(self._t(outgoing[q][0]), q))
                reqs.append(dist.isend(self._t(outgoing[q][1]), q))
        for q in range(self.world):
            if q != self.rank and n_in[q]:
                p = torch.empty((n_in[q], 4), dtype=torch.float32, device=self.device)
                i = torch.empty(n_in[q], dtype=torch.int32, device=self.device)
                dist.recv(p, q)
                dist.recv(i, q)
                recv.append((p.cpu().numpy(), i.cpu().numpy()))
        for r in reqs:
            r.wait()
        return recv

    def distribute(self):
        self.sim.slab_sort_out(self.world)
        self.sim.slab_keep_own()

    def step(self, kind, dt):
        s, torch, dist = self.sim, self.torch, self.dist
        out = [None] * self.world
        if self.rank > 0:
            out[self.rank - 1] = s.slab_boundary(0)
        if self.rank + 1 < self.world:
            out[self.rank + 1] = s.slab_boundary(1)
        for parts, ids in self._exchange(out):
            s.slab_add(parts, ids)
        s.slab_step_a(kind)
        for which in (ROWS_LABELS, U_FRONT, V_FRONT):
            mine = self._t(s.get_rows(which, self.lo, self.hi))
            slabs = [torch.empty((hi - lo, s.nx), dtype=mine.dtype, device=mine.device)
                     for lo, hi in self.rows]
            dist.all_gather(slabs, mine)
            for q, (lo, hi) in enumerate(self.rows):
                if q != self.rank:
                    s.set_rows(which, lo, hi, slabs[q].cpu().numpy())
        s.slab_step_b(kind, dt)
        s.slab_step_c(kind, dt)
        counts = s.slab_sort_out(self.world)
        out = [s.slab_take(d, counts[d]) if d != self.rank and counts[d] else None
               for d in range(self.world)]
        s.slab_keep_own()
        for parts, ids in self._exchange(out):
            s.slab_add(parts, ids)
        return sum(c for d, c in enumerate(counts) if d != self.rank)

    def particles(self):
        """The whole set in global-id order, gathered on every rank."""
        torch, dist = self.torch, self.dist
        parts, ids = self.sim.slab_get()
        n = torch.tensor([ids.shape[0]], dtype=torch.int64)
        n = n.to(self.device) if self.device is not None else n
        ns = [torch.empty_like(n) for _ in range(self.world)]
        dist.all_gather(ns, n)
        ns = [int(v.item()) for v in ns]
        cap = max(ns) if ns else 0
        pp = np.zeros((cap, 4), dtype=np.float32); pp[:ids.shape[0]] = parts
        ii = np.zeros(cap, dtype=np.int32); ii[:ids.shape[0]] = ids
        gp = [torch.empty((cap, 4), dtype=torch.float32, device=self.device) for _ in range(self.world)]
        gi = [torch.empty(cap, dtype=torch.int32, device=self.device) for _ in range(self.world)]
        dist.all_gather(gp, self._t(pp))
        dist.all_gather(gi, self._t(ii))
        out = np.empty((sum(ns), 4), dtype=np.float32)
        for q in range(self.world):
            out[gi[q][:ns[q]].cpu().numpy()] = gp[q][:ns[q]].cpu().numpy()
        return out
