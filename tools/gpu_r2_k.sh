#!/bin/bash
# round 2, call K (2 GPUs): sharded full steps, repeated (an intermittent mailbox time-out)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 $TR --master-port 29542 bench.py --gpus 2 "$@" --warmup 1 --no-cpu-baseline --no-optin --no-e2e --verbose > gpurun_out/k_$name.json 2> gpurun_out/k_$name.err
  echo "== $name rc=$?"; grep "libfsb error" gpurun_out/k_$name.err | head -4
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/k_$name.json"))
    print("$name: iters", d["cg_iters_per_step"], "us/iter %.2f" % d["roofline"]["avg_iteration_us"], "relres", d["cg_relres"], "ms/step %.1f" % d["ms_per_step"])
except Exception as e:
    print("$name: no line:", e)
PY
}
for k in 1 2 3; do run picflip1024_$k X=1 -- --workload picflip1024 --steps 6; done
for k in 1 2 3; do run picflip2048_$k X=1 -- --workload picflip2048 --steps 3; done
