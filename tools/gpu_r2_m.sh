#!/bin/bash
# round 2, call M (1 GPU): full GPU test suite + the default bench line with the new roofline / scale_cg8192 objects
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/m_pytest.log
tail -5 gpurun_out/m_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_default.json 2> gpurun_out/m_bench_default.err
echo "bench rc=$?"; tail -3 gpurun_out/m_bench_default.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/m_bench_default.json"))
print("default: ms/step %.1f iters %.0f" % (d["ms_per_step"], d["cg_iters_per_step"]))
print(json.dumps(d["roofline"], indent=1)); print(json.dumps(d["scale_cg8192"])); print(d["e2e"]); print(d["optin_multigrid"])
PY
