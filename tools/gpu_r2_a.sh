#!/bin/bash
# round 2, call A: first run of the one-sweep CG: its parity tests, then per-iteration times
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pressure or deferred or one_sweep or active_tile or multigrid_falls" > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -15 gpurun_out/a_pytest.log
for wl in cg4096 cg1024 cg8192; do
  for mode in one fused; do
    FSB_CG_MODE=$mode timeout 300 python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-optin --no-e2e > gpurun_out/a_${wl}_${mode}.json 2> gpurun_out/a_${wl}_${mode}.err
    echo "$wl $mode rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/a_${wl}_${mode}.json"))
    print("  iters", d["cg_iters_per_step"], "us/iter", d["roofline"]["avg_iteration_us"], "relres", d["cg_relres"], "ms", d["ms_per_step"])
except Exception as e:
    print("  no line:", e); print(open("gpurun_out/a_${wl}_${mode}.err").read()[-1500:])
PY
  done
done
