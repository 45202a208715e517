#!/bin/bash
# round 2, call S (1 GPU): ring depth
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py "$@" --steps 2 --warmup 1 --no-cpu-baseline --no-optin --no-e2e --no-scale > gpurun_out/s_$name.json 2> gpurun_out/s_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s_$name.json"))
    print("$name: iters", d["cg_iters_per_step"], "us/iter %.2f" % d["roofline"]["avg_iteration_us"], "frac %.3f" % d["roofline"]["frac"], d["clocks"])
except Exception as e:
    print("$name: no line:", e); print(open("gpurun_out/s_$name.err").read()[-800:])
PY
}
for st in 5 3 4 2; do run cg4096_st$st FSB_CG_STAGES=$st -- --workload cg4096; done
for st in 5 3 2; do run slab8_st$st FSB_CG_STAGES=$st -- --workload cg8192slab8; done
for st in 5 3; do run cg8192_st$st FSB_CG_STAGES=$st -- --workload cg8192; done
run slab8_st3_both FSB_CG_STAGES=3 FSB_CG_PHINT=1 FSB_CG_XHINT=1 -- --workload cg8192slab8
run slab4_st3 FSB_CG_STAGES=3 -- --workload cg8192slab4
run cg4096_st5_again FSB_CG_STAGES=5 -- --workload cg4096
