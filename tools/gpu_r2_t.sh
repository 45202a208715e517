#!/bin/bash
# round 2, call T (1 GPU): full GPU suite (SL gather, averaging extension, state files, slab guards, big goldens), SL benches
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/t_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t_pytest.log
tail -12 gpurun_out/t_pytest.log
for wl in sl1024 sl4096; do
  timeout 600 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline --no-optin --no-scale > gpurun_out/t_$wl.json 2> gpurun_out/t_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/t_$wl.json"))
    print("$wl: ms/step %.2f iters %.0f us/iter %.2f" % (d["ms_per_step"], d["cg_iters_per_step"], d["roofline"]["avg_iteration_us"]), d["stage_ms_per_step"], d["stage_roofline"].get("advect_sl"), d["stage_roofline"].get("advect_part"))
except Exception as e:
    print("$wl: no line:", e); print(open("gpurun_out/t_$wl.err").read()[-800:])
PY
done
