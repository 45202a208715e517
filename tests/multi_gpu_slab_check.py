"""Run under torchrun: particle slabs over torch.distributed (DistSlabs) against an unpartitioned
run on every rank.  --backend nccl: one rank per GPU; --backend gloo: CPU tensors as transport, all
ranks may share one GPU (two contexts on device 0), which checks the distributed call sequence
without a second GPU.  Prints one JSON line on rank 0; exit code 0 = bit-identical.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multi_gpu_slab_check.py --backend gloo
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
    ap.add_argument("--backend", default="nccl", choices=["nccl", "gloo"])
    ap.add_argument("--grid", dest="n", type=int, default=128)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--kind", default="picflip", choices=["picflip", "flip", "pic", "sl"],
                    help="step kind (sl: src/FluidSolver.cpp:99-134 -- label rows only, no ghost rows)")
    ap.add_argument("--shard-cg", action="store_true",
                    help="additionally shard the pressure CG over the ranks (peer memory; nccl only)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import scenes
    from fluid_simulation_b200 import capi, sharding

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank)) % max(1, torch.cuda.device_count())
    torch.cuda.set_device(local)
    if args.backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        device = torch.device("cuda", local)
    else:
        dist.init_process_group("gloo")
        device = None
    n = args.n
    src = scenes.dam_break_args(n)
    ref = capi.Sim(n, n, 1.0, 1.0, 0.01, 0.05, device=local)
    own = capi.Sim(n, n, 1.0, 1.0, 0.01, 0.05, device=local)
    for s in (ref, own):
        s.set_cg(2000, 1e-6)
        s.emit_source(*src)
    slabs = sharding.DistSlabs(own, dist, device)
    slabs.distribute()
    if args.shard_cg:
        sharding.connect(own, dist, device)
    ok, moved = True, 0
    kind = {"picflip": capi.STEP_PICFLIP, "flip": capi.STEP_FLIP, "pic": capi.STEP_PIC, "sl": capi.STEP_SL}[args.kind]
    for step in range(args.steps):
        ref.step(kind, 0.01)
        moved += slabs.step(kind, 0.01)
        ok = ok and own.cg_info() == ref.cg_info()
        ok = ok and np.array_equal(own.get_cell_types(), ref.get_cell_types())
        for w in (capi.U_FRONT, capi.V_FRONT, capi.U_BACK, capi.V_BACK):
            ok = ok and np.array_equal(own.get_grid(w), ref.get_grid(w))
    allp = slabs.particles()
    ok = ok and np.array_equal(allp, ref.get_particles())
    flag = torch.tensor([1 if ok else 0])
    flag = flag.to(device) if device is not None else flag
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out = {"world": world, "backend": args.backend, "n": n, "steps": args.steps, "kind": args.kind, "cg_sharded": bool(args.shard_cg),
           "own_particles": int(own.num_particles()), "all_particles": int(allp.shape[0]),
           "migrated_by_this_rank": int(moved), "ok": bool(flag.item())}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if args.shard_cg:
        own.shard_disconnect()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if out["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
