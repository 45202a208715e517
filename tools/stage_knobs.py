#!/usr/bin/env python
"""Stage-kernel knobs on the default workload (4096^2 PIC/FLIP tank step), one process, one GPU.

Every configuration gets its own context (the knobs are read when a context is created), the same
particle set, one untimed and `--steps` timed steps with the CG capped (the stages around the solve do not
depend on how long it iterates), and must leave the SAME labels, velocity grids and particles as the first
configuration -- the knobs change launch geometry and fusion, never arithmetic.  Prints ms per step of
every stage (CUDA events of fsb_profile_read).

    python tools/stage_knobs.py [--grid 4096] [--steps 3] [--cap 20]
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
    ("default", {}),
    ("sort: 1 particle per thread in the counting and gather passes (rounds 1 - 2)", {"FSB_SORT_PER": "1"}),
    ("sort: 4 particles per thread", {"FSB_SORT_PER": "4"}),
    ("sort: 1 cell per thread in the in-cell ordering pass (rounds 1 - 2)", {"FSB_CANON_PER": "1"}),
    ("p2g: record loaded when needed (rounds 1 - 2)", {"FSB_P2G_PIPE": "0"}),
    ("g2p: 1 particle per thread (rounds 1 - 2)", {"FSB_G2P_PER": "1"}),
    ("g2p: 4 particles per thread", {"FSB_G2P_PER": "4"}),
    ("rhs: 8 CTAs/SM (rounds 1 - 2)", {"FSB_BUILD_BLOCKS_PER_SM": "8"}),
]
KNOBS = sorted({k for _, env in CONFIGS for k in env})
if os.environ.get("FSB_KNOBS_ONLY_DEFAULT"):
    CONFIGS = CONFIGS[:1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cap", type=int, default=20)
    args = ap.parse_args()
    import torch
    import bench
    from fluid_simulation_b200 import capi
    n = args.grid
    dt = float(np.float32(0.01 * 64.0 / n))
    parts = bench.tank_particles(n, 2)
    host = torch.empty(parts.shape, dtype=torch.float32, pin_memory=True)
    host.numpy()[:] = parts
    n_part = parts.shape[0]
    del parts
    first, rows = None, []
    for name, env in CONFIGS:
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        sim = capi.Sim(n, n, 1.0, 1.0, dt, 0.02)
        sim.set_cg(args.cap, 1e-6)
        sim.set_particles_ptr(host.data_ptr(), n_part)
        sim.step(capi.STEP_PICFLIP, dt)
        sim.synchronize()
        sim.profile_enable(True)
        for _ in range(args.steps):
            sim.step(capi.STEP_PICFLIP, dt)
        sim.synchronize()
        prof = sim.profile_read()
        ms = {k: round(v[0] / args.steps, 4) for k, v in prof.items() if v[1]}
        state = [sim.get_cell_types(), sim.get_grid(capi.U_FRONT), sim.get_grid(capi.V_FRONT),
                 sim.get_grid(capi.U_BACK), sim.get_grid(capi.V_BACK), sim.get_pressure(), sim.get_particles()]
        tiles = sim.cg_info()
        if first is None:
            first = state
            same = True
        else:
            same = all(np.array_equal(a, b) for a, b in zip(first, state))
        rows.append({"config": name, "env": env, "identical_to_default": bool(same), "cg": list(tiles), "ms": ms})
        print(json.dumps(rows[-1]), flush=True)
        sim.close()
        del sim, state
    bad = [r["config"] for r in rows if not r["identical_to_default"]]
    print("ALL IDENTICAL" if not bad else f"DIFFERENT RESULTS: {bad}", flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
