# round 1, call z (1 GPU): full GPU suite, smoke, default bench (both arms) with the opt-in multigrid side measurement
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -10 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
timeout 300 python bench.py --impl reference > gpurun_out/bench_z_reference.json 2> /dev/null; echo "bench reference rc=$?"
timeout 600 python bench.py > gpurun_out/bench_z_default.json 2> gpurun_out/bench_z_default.err; echo "bench default rc=$?"
wc -l gpurun_out/bench_z_default.json
python -c "import json;d=json.load(open('gpurun_out/bench_z_default.json'));print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['avg_iteration_us'], d['optin_multigrid'], d['cpu_baseline']['value'])"
