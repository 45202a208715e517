#!/bin/bash
# round 2, call I: full GPU test suite, then the default bench line and the CG-only timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/i_pytest.log
tail -5 gpurun_out/i_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/i_bench_default.json 2> gpurun_out/i_bench_default.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/i_bench_default.json"))
print("default: ms/step %.1f iters %.0f us/iter %.2f stages" % (d["ms_per_step"], d["cg_iters_per_step"], d["roofline"]["avg_iteration_us"]), d["stage_ms_per_step"])
PY
for wl in cg4096 cg1024 cg8192; do
  timeout 300 python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-optin --no-e2e > gpurun_out/i_$wl.json 2> gpurun_out/i_$wl.err
  python - <<PY
import json
d=json.load(open("gpurun_out/i_$wl.json"))
print("$wl: iters", d["cg_iters_per_step"], "us/iter %.2f" % d["roofline"]["avg_iteration_us"], "relres", d["cg_relres"])
PY
done
