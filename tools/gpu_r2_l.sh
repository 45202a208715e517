#!/bin/bash
# round 2, call L (2 GPUs): cost of the cross-GPU reduction of the one-sweep solve
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 $TR --master-port 29542 bench.py --gpus 2 "$@" --warmup 1 --steps 2 --no-cpu-baseline --no-optin --no-e2e --no-scale > gpurun_out/l_$name.json 2> gpurun_out/l_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/l_$name.json"))
    print("$name: iters", d["cg_iters_per_step"], "us/iter %.2f" % d["roofline"]["avg_iteration_us"], "relres", d["cg_relres"], "ms/step %.1f" % d["ms_per_step"])
except Exception as e:
    print("$name: no line:", e); print(open("gpurun_out/l_$name.err").read()[-600:])
PY
}
run notiles FSB_CG_DEBUG_NOTILES=1 -- --workload cg1024 --cg-cap 3000
run notiles_pollsys FSB_CG_DEBUG_NOTILES=1 FSB_CG_POLL_FENCE_SYS=1 -- --workload cg1024 --cg-cap 3000
run cg1024 X=1 -- --workload cg1024
run cg4096 X=1 -- --workload cg4096
run cg8192 X=1 -- --workload cg8192
run picflip2048 X=1 -- --workload picflip2048
timeout 300 $TR --master-port 29541 tests/multi_gpu_cg_check.py --grid 1030 > gpurun_out/l_check.log 2>&1; echo "check rc=$? $(grep '^{' gpurun_out/l_check.log | tail -1 | cut -c1-260)"
