#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
cp fluid_simulation_b200/lib/libfsb.so /tmp/head.so
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py "$@" --steps 2 --warmup 1 --no-cpu-baseline --no-optin --no-e2e --no-scale > gpurun_out/v_$name.json 2> gpurun_out/v_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/v_$name.json"))
    print("$name: us/iter %.2f" % d["roofline"]["avg_iteration_us"], "iters", d["cg_iters_per_step"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name: no line:", e)
PY
}
for v in prev head prev head; do
  if [ $v = head ]; then cp /tmp/head.so fluid_simulation_b200/lib/libfsb.so; else cp build_variants/$v/libfsb.so fluid_simulation_b200/lib/libfsb.so; fi
  run ${v}_notiles FSB_CG_DEBUG_NOTILES=1 -- --workload cg1024 --cg-cap 3000
  run ${v}_cg1024 X=1 -- --workload cg1024
  run ${v}_cg4096 X=1 -- --workload cg4096
done
cp /tmp/head.so fluid_simulation_b200/lib/libfsb.so
TAG=determinism REPS=10 timeout 200 python tools/determinism_check.py 2>&1 | tail -1
