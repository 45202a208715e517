#!/bin/bash
# round 2, call J (2 GPUs): sharded one-sweep CG -- correctness check against the single-GPU solve, then timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for grid in 512 1030 2048; do
  timeout 300 $TR --master-port 29541 tests/multi_gpu_cg_check.py --grid $grid > gpurun_out/j_check_$grid.log 2>&1
  echo "check $grid rc=$? $(grep '^{' gpurun_out/j_check_$grid.log | tail -1)"
done
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 400 $TR --master-port 29542 bench.py --gpus 2 "$@" --steps 2 --warmup 1 --no-cpu-baseline --no-optin --no-e2e > gpurun_out/j_$name.json 2> gpurun_out/j_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/j_$name.json"))
    print("$name: iters", d["cg_iters_per_step"], "us/iter %.2f" % d["roofline"]["avg_iteration_us"], "relres", d["cg_relres"], "ms/step %.1f" % d["ms_per_step"])
except Exception as e:
    print("$name: no line:", e); print(open("gpurun_out/j_$name.err").read()[-1500:])
PY
}
run cg8192_one X=1 -- --workload cg8192
run cg8192_fused FSB_CG_MODE=fused -- --workload cg8192
run cg4096_one X=1 -- --workload cg4096
run cg4096_graph FSB_CG_MODE=graph -- --workload cg4096
run cg1024_one X=1 -- --workload cg1024
run picflip4096_one X=1 -- --workload picflip4096
