#!/bin/bash
# round 2, call C: one-sweep CG with early loads; bare barrier cost (sweeps without tiles)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pressure or deferred or one_sweep or active_tile or multigrid_falls" > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -4 gpurun_out/c_pytest.log
run() { # name, env..., -- bench args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py "$@" --steps 2 --warmup 1 --no-cpu-baseline --no-optin --no-e2e > gpurun_out/c_$name.json 2> gpurun_out/c_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c_$name.json"))
    print("$name: iters", d["cg_iters_per_step"], "us/iter %.2f" % d["roofline"]["avg_iteration_us"], "relres", d["cg_relres"])
except Exception as e:
    print("$name: no line:", e); print(open("gpurun_out/c_$name.err").read()[-800:])
PY
}
run cg4096_early X=1 -- --workload cg4096
run cg4096_noearly FSB_CG_PREFETCH=0 -- --workload cg4096
run cg1024_early X=1 -- --workload cg1024
run cg1024_noearly FSB_CG_PREFETCH=0 -- --workload cg1024
run cg1024_1cta FSB_CG_CTAS_PER_SM=1 -- --workload cg1024
run cg8192_early X=1 -- --workload cg8192
run barrier_296 FSB_CG_DEBUG_NOTILES=1 -- --workload cg1024 --cg-cap 3000
run barrier_148 FSB_CG_DEBUG_NOTILES=1 FSB_CG_CTAS_PER_SM=1 -- --workload cg1024 --cg-cap 3000
run barrier_296_noearly FSB_CG_DEBUG_NOTILES=1 FSB_CG_PREFETCH=0 -- --workload cg1024 --cg-cap 3000
run cg2048 X=1 -- --workload cg1024
