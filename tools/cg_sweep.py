#!/usr/bin/env python
"""Sweep of the CG launch modes / tuning knobs on the CG-only workloads, one process, one GPU
(or one rank per GPU under torchrun).  Every configuration runs the SAME solve (tank labels, swirl
+ gravity field, bench.py's cgN workloads) with the iteration cap `--cap`, and must reproduce the
iteration count and the pressure field of the first configuration; prints us / iteration.

    python tools/cg_sweep.py --grids 1024,4096 --cap 2000
    python -m torch.distributed.run --nproc-per-node 2 ... tools/cg_sweep.py --grids 4096,8192
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = [
    ("two-kernel  serp=0", dict(FSB_CG_MODE="graph", FSB_CG_SERP="0")),
    ("two-kernel  serp=1", dict(FSB_CG_MODE="graph", FSB_CG_SERP="1")),
    ("fused serp=0 pre=0", dict(FSB_CG_MODE="fused", FSB_CG_SERP="0", FSB_CG_PREFETCH="0")),
    ("fused serp=0 pre=1", dict(FSB_CG_MODE="fused", FSB_CG_SERP="0", FSB_CG_PREFETCH="1")),
    ("fused serp=1 pre=1", dict(FSB_CG_MODE="fused", FSB_CG_SERP="1", FSB_CG_PREFETCH="1")),
    ("fused serp=1 pre=1 xhint", dict(FSB_CG_MODE="fused", FSB_CG_SERP="1", FSB_CG_PREFETCH="1",
                                      FSB_CG_XHINT="1")),
    ("fused serp=1 pre=1 1cta/sm", dict(FSB_CG_MODE="fused", FSB_CG_SERP="1", FSB_CG_PREFETCH="1",
                                        FSB_CG_CTAS_PER_SM="1")),
    # L2 residency hints (r + stencil codes evict-last on keep/4 of the accesses, x and the dead
    # old direction evict-first)
    ("fused keep=4", dict(FSB_CG_MODE="fused", FSB_CG_KEEP="4")),                              # 7
    ("fused keep=4 xhint", dict(FSB_CG_MODE="fused", FSB_CG_KEEP="4", FSB_CG_XHINT="1")),      # 8
    ("fused keep=4 xhint phint", dict(FSB_CG_MODE="fused", FSB_CG_KEEP="4", FSB_CG_XHINT="1",
                                      FSB_CG_PHINT="1")),                                      # 9
    ("fused keep=3 xhint phint", dict(FSB_CG_MODE="fused", FSB_CG_KEEP="3", FSB_CG_XHINT="1",
                                      FSB_CG_PHINT="1")),                                      # 10
    ("fused keep=2 xhint phint", dict(FSB_CG_MODE="fused", FSB_CG_KEEP="2", FSB_CG_XHINT="1",
                                      FSB_CG_PHINT="1")),                                      # 11
    ("fused keep=1 xhint phint", dict(FSB_CG_MODE="fused", FSB_CG_KEEP="1", FSB_CG_XHINT="1",
                                      FSB_CG_PHINT="1")),                                      # 12
    ("fused xhint phint", dict(FSB_CG_MODE="fused", FSB_CG_XHINT="1", FSB_CG_PHINT="1")),      # 13
    ("fused keep=4 xhint phint rows=8", dict(FSB_CG_MODE="fused", FSB_CG_KEEP="4", FSB_CG_XHINT="1",
                                             FSB_CG_PHINT="1", FSB_CG_TILE_ROWS="8")),         # 14
    ("fused keep=4 xhint phint rows=32", dict(FSB_CG_MODE="fused", FSB_CG_KEEP="4", FSB_CG_XHINT="1",
                                              FSB_CG_PHINT="1", FSB_CG_TILE_ROWS="32")),       # 15
    ("fused rows=8", dict(FSB_CG_MODE="fused", FSB_CG_TILE_ROWS="8")),                         # 16
    ("fused rows=32", dict(FSB_CG_MODE="fused", FSB_CG_TILE_ROWS="32")),                       # 17
    ("fused stages=2", dict(FSB_CG_MODE="fused", FSB_CG_STAGES="2")),                          # 18
    ("fused stages=3", dict(FSB_CG_MODE="fused", FSB_CG_STAGES="3")),                          # 19
    # L2 set-aside + access-policy window over r
    ("fused persist=32MB", dict(FSB_CG_MODE="fused", FSB_CG_PERSIST_MB="32")),                 # 20
    ("fused persist=64MB", dict(FSB_CG_MODE="fused", FSB_CG_PERSIST_MB="64")),                 # 21
    ("fused persist=96MB", dict(FSB_CG_MODE="fused", FSB_CG_PERSIST_MB="96")),                 # 22
    ("fused persist=64MB xhint phint", dict(FSB_CG_MODE="fused", FSB_CG_PERSIST_MB="64",
                                            FSB_CG_XHINT="1", FSB_CG_PHINT="1")),              # 23
    ("fused persist=48MB", dict(FSB_CG_MODE="fused", FSB_CG_PERSIST_MB="48")),                 # 24
    ("fused persist=40MB", dict(FSB_CG_MODE="fused", FSB_CG_PERSIST_MB="40")),                 # 25
    ("fused persist=56MB", dict(FSB_CG_MODE="fused", FSB_CG_PERSIST_MB="56")),                 # 26
    ("fused persist=32MB miss=normal", dict(FSB_CG_MODE="fused", FSB_CG_PERSIST_MB="32",
                                            FSB_CG_PERSIST_MISS_NORMAL="1")),                  # 27
    ("fused persist=48MB miss=normal", dict(FSB_CG_MODE="fused", FSB_CG_PERSIST_MB="48",
                                            FSB_CG_PERSIST_MISS_NORMAL="1")),                  # 28
    ("fused persist=64MB miss=normal", dict(FSB_CG_MODE="fused", FSB_CG_PERSIST_MB="64",
                                            FSB_CG_PERSIST_MISS_NORMAL="1")),                  # 29
    ("fused persist=16MB miss=normal", dict(FSB_CG_MODE="fused", FSB_CG_PERSIST_MB="16",
                                            FSB_CG_PERSIST_MISS_NORMAL="1")),                  # 30
    ("fused xdefer=0", dict(FSB_CG_MODE="fused", FSB_CG_XDEFER="0")),                          # 31
    ("fused xdefer=0 persist=48MB", dict(FSB_CG_MODE="fused", FSB_CG_XDEFER="0",
                                         FSB_CG_PERSIST_MB="48")),                             # 32
    ("fused keep=1 xhint phint", dict(FSB_CG_MODE="fused", FSB_CG_KEEP="1", FSB_CG_XHINT="1",
                                      FSB_CG_PHINT="1")),                                      # 33
    ("fused phint", dict(FSB_CG_MODE="fused", FSB_CG_PHINT="1")),                              # 34
    # sharded solves: slab boundary tiles first in every sweep (default) vs natural order
    ("two-kernel edge-first=0", dict(FSB_CG_MODE="graph", FSB_CG_EDGE_FIRST="0")),             # 35
    ("fused edge-first=0", dict(FSB_CG_MODE="fused", FSB_CG_EDGE_FIRST="0")),                  # 36
]
KNOBS = ["FSB_CG_MODE", "FSB_CG_SERP", "FSB_CG_PREFETCH", "FSB_CG_XHINT", "FSB_CG_CTAS_PER_SM",
         "FSB_CG_TILE_ROWS", "FSB_CG_STAGES", "FSB_CG_KEEP", "FSB_CG_PHINT", "FSB_CG_PERSIST_MB", "FSB_CG_PERSIST_MISS_NORMAL", "FSB_CG_XDEFER", "FSB_CG_EDGE_FIRST", "FSB_CG_SKIP_TILES"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grids", default="1024,4096")
    ap.add_argument("--cap", type=int, default=2000)
    ap.add_argument("--only", default=None, help="comma-separated indices into CONFIGS")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    from bench import tank_fields
    from fluid_simulation_b200 import capi, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    picks = range(len(CONFIGS)) if args.only is None else [int(k) for k in args.only.split(",")]
    rows = []
    for n in [int(g) for g in args.grids.split(",")]:
        dt = float(np.float32(0.01 * 64.0 / n))
        lab, u0, v0 = tank_fields(n)
        ref = None
        for k in picks:
            name, env = CONFIGS[k]
            for kn in KNOBS:
                os.environ.pop(kn, None)
            os.environ.update(env)
            sim = capi.Sim(n, n, 1.0, 1.0, dt, 0.02, device=local)
            sim.set_cg(args.cap, 1e-6)
            sim.set_cell_types(lab)
            if world > 1:
                sharding.connect(sim, dist, torch.device("cuda", local))
            best = None
            for rep in range(2):  # first solve: configuration + graph capture + warm-up
                sim.set_grid(capi.U_FRONT, u0)
                sim.set_grid(capi.V_FRONT, v0)
                sim.synchronize()
                if world > 1:
                    dist.barrier()
                sim.profile_enable(True)
                sim.profile_read()
                sim.pressure_solve(dt, dt)
                sim.synchronize()
                prof = sim.profile_read()
                sim.profile_enable(False)
                iters, relres = sim.cg_info()
                ms = prof["cg"][0]
                if world > 1:
                    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t.item())
                us = 1e3 * ms / max(iters, 1)
                best = us if best is None else min(best, us)
            x = sim.get_pressure().astype(np.float64)
            if ref is None:
                ref = (iters, x)
            rel = float(np.linalg.norm(x - ref[1]) / max(np.linalg.norm(ref[1]), 1e-300))
            ok = iters == ref[0] and rel < 1e-4
            row = dict(n=n, world=world, config=name, us_per_iter=round(best, 2), iters=iters,
                       relres=relres, rel_vs_first=rel, ok=bool(ok))
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
            if world > 1:
                sim.shard_disconnect()
                dist.barrier()
            sim.close()
    if rank == 0 and args.out:
        json.dump(rows, open(args.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()
    return 0 if all(r["ok"] for r in rows) else 1


if __name__ == "__main__":
    sys.exit(main())
