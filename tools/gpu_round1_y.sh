# round 1, call y (1 GPU): opt-in multigrid-preconditioned CG -- tests, then the 4096^2 step with it
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multigrid" > gpurun_out/pytest_mg.log 2>&1; echo "pytest mg rc=$?"
tail -30 gpurun_out/pytest_mg.log
timeout 300 python bench.py --workload cg4096 --precond mg --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_mg_cg4096.json 2> gpurun_out/bench_mg_cg4096.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_mg_cg4096.json'));print(d['ms_per_step'], d['cg_iters_per_step'], d['cg_relres'], d['stage_ms_per_step'], d['gpu_launches'], d['roofline']['kernel'][:30])"
timeout 300 python bench.py --precond mg --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_mg_picflip4096.json 2> gpurun_out/bench_mg_picflip4096.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_mg_picflip4096.json'));print(d['ms_per_step'], d['value'], d['cg_iters_per_step'], d['cg_relres'], d['stage_ms_per_step'], d['e2e'])"
