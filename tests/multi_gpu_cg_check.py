"""Run under torchrun (one rank per GPU): the row-slab sharded pressure solve against the
single-GPU solve of the same problem on every rank, then a sharded full PIC/FLIP step against an
unsharded one.  Prints one JSON line on rank 0; exit code 0 = all checks passed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tests/multi_gpu_cg_check.py [--grid 1024]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", dest="n", type=int, default=1024)
    ap.add_argument("--tol", type=float, default=1e-6)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import scenes
    from fluid_simulation_b200 import capi, sharding

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n
    dt = float(np.float32(0.01 * 64.0 / n))
    parts = scenes.tank_particles(n, np.random.default_rng(1234), 2)
    grav = float(np.float32(-9.82))

    def prepared():
        s = capi.Sim(n, n, 1.0, 1.0, dt, 0.02, device=local)
        s.set_cg(200000, args.tol)
        s.set_particles(parts)
        s.classify_cells(); s.p2g_spread(); s.save_previous()
        s.add_acceleration(0.0, grav, dt); s.enforce_dirichlet(); s.extend_velocity(2)
        return s

    out = {"world": world, "n": n}
    # ---- 1. pressure solve: single GPU vs sharded
    a = prepared()
    a.timer_start(); a.pressure_solve(dt, dt); t_single = a.timer_stop()
    it_a, err_a = a.cg_info()
    xa, ua = a.get_pressure(), a.get_grid(capi.U_FRONT)

    b = prepared()
    lo, hi = sharding.connect(b, dist, torch.device("cuda", local))
    dist.barrier()
    b.timer_start(); b.pressure_solve(dt, dt); t_shard = b.timer_stop()
    it_b, err_b = b.cg_info()
    xb, ub = b.get_pressure(), b.get_grid(capi.U_FRONT)
    rel = float(np.linalg.norm(xb.astype(np.float64) - xa) / np.linalg.norm(xa.astype(np.float64)))
    relu = scenes.field_rel_err(ub, ua)
    ok = abs(it_b - it_a) <= max(2, 0.05 * it_a) and err_b < args.tol and rel < 2e-3 and relu < 1e-3
    # every rank must hold the SAME pressure field and iteration count
    h = torch.tensor([float(np.float64(xb.astype(np.float64).sum())), float(it_b)], device="cuda", dtype=torch.float64)
    hs = [torch.empty_like(h) for _ in range(world)]
    dist.all_gather(hs, h)
    same = all(bool((x == hs[0]).all()) for x in hs)
    out.update(rows=[lo, hi], iters_single=it_a, iters_sharded=it_b, relres=err_b, rel_pressure=rel,
               rel_velocity=relu, ranks_identical=same, ms_single=t_single, ms_sharded=t_shard,
               cg_speedup=(t_single / it_a) / (t_shard / it_b) if it_a and it_b else None)
    ok = ok and same

    # ---- 2. three full PIC/FLIP steps: unsharded vs sharded contexts
    c = capi.Sim(n, n, 1.0, 1.0, dt, 0.02, device=local); c.set_cg(200000, args.tol); c.set_particles(parts)
    d = capi.Sim(n, n, 1.0, 1.0, dt, 0.02, device=local); d.set_cg(200000, args.tol); d.set_particles(parts)
    sharding.connect(d, dist, torch.device("cuda", local))
    for _ in range(3):
        c.step(capi.STEP_PICFLIP, dt)
        d.step(capi.STEP_PICFLIP, dt)
    pc, pd = c.get_particles(), d.get_particles()
    perr = float(np.abs(pc[:, :2] - pd[:, :2]).max())
    lab_same = bool(np.array_equal(c.get_cell_types(), d.get_cell_types()))
    out.update(step_particle_err=perr, step_labels_equal=lab_same)
    ok = ok and perr < 1e-4 and lab_same

    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["ok"] = bool(flag.item())
    if rank == 0:
        print(json.dumps(out), flush=True)
    d.shard_disconnect(); b.shard_disconnect()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if out["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
