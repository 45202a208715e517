#!/bin/bash
# round 2, call O (2 GPUs): time stamps inside the reduction
cd "$GRAFT_REPO_ROOT" || exit 1
rm -rf gpurun_out/dbg; mkdir -p gpurun_out/dbg
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
FSB_CG_DEBUG_NOTILES=1 FSB_CG_DEBUG_TIMES=1 FSB_CG_DEBUG_SUMS=gpurun_out/dbg/t1 timeout 200 python bench.py --workload cg1024 --cg-cap 400 --warmup 0 --steps 1 --no-cpu-baseline --no-optin --no-e2e --no-scale > /dev/null 2>&1
FSB_CG_DEBUG_TIMES=1 FSB_CG_DEBUG_SUMS=gpurun_out/dbg/r1 timeout 200 python bench.py --workload cg4096 --cg-cap 400 --warmup 0 --steps 1 --no-cpu-baseline --no-optin --no-e2e --no-scale > /dev/null 2>&1
export FSB_CG_DEBUG_TIMES=1
FSB_CG_DEBUG_NOTILES=1 FSB_CG_DEBUG_SUMS=gpurun_out/dbg/t2 timeout 200 $TR --master-port 29542 bench.py --gpus 2 --workload cg1024 --cg-cap 400 --warmup 0 --steps 1 --no-cpu-baseline --no-optin --no-e2e --no-scale > /dev/null 2>&1
FSB_CG_DEBUG_SUMS=gpurun_out/dbg/r2 timeout 200 $TR --master-port 29542 bench.py --gpus 2 --workload cg4096 --cg-cap 400 --warmup 0 --steps 1 --no-cpu-baseline --no-optin --no-e2e --no-scale > /dev/null 2>&1
ls -la gpurun_out/dbg
