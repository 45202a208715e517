#!/bin/bash
# round 2, call P (1 GPU): CG kernel variants: tests, timings, determinism
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pressure or deferred or one_sweep or active_tile or multigrid_falls or full_steps" > gpurun_out/p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p_pytest.log; tail -3 gpurun_out/p_pytest.log
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py "$@" --steps 2 --warmup 1 --no-cpu-baseline --no-optin --no-e2e --no-scale > gpurun_out/p_$name.json 2> gpurun_out/p_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/p_$name.json"))
    print("$name: iters", d["cg_iters_per_step"], "us/iter %.2f" % d["roofline"]["avg_iteration_us"], "relres", d["cg_relres"], "frac %.3f" % d["roofline"]["frac"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name: no line:", e); print(open("gpurun_out/p_$name.err").read()[-800:])
PY
}
run cg4096 X=1 -- --workload cg4096
run cg4096_st4 FSB_CG_STAGES=4 -- --workload cg4096
run cg4096_st5 FSB_CG_STAGES=5 -- --workload cg4096
run slab8 X=1 -- --workload cg8192slab8
run slab8_st4 FSB_CG_STAGES=4 -- --workload cg8192slab8
run cg1024 X=1 -- --workload cg1024
run cg8192 X=1 -- --workload cg8192
TAG=determinism REPS=12 timeout 200 python tools/gpu_r2_e.py 2>&1 | tail -1
TAG=determinism_st5 REPS=12 FSB_CG_STAGES=5 timeout 200 python tools/gpu_r2_e.py 2>&1 | tail -1
