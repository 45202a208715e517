# round 1, call aa (1 GPU): active-tile list in the CG sweeps -- full GPU suite, mode sweep, default bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python tools/cg_sweep.py --grids 4096,1024 --cap 2000 --only 1,4 --out gpurun_out/cg_sweep_aa.json > gpurun_out/cg_sweep_aa.log 2>&1; echo "sweep rc=$?"
grep "^{" gpurun_out/cg_sweep_aa.log | cut -c1-110
FSB_CG_SKIP_TILES=0 timeout 300 python tools/cg_sweep.py --grids 4096 --cap 2000 --only 4 > gpurun_out/cg_sweep_aa0.log 2>&1; grep "^{" gpurun_out/cg_sweep_aa0.log | cut -c1-110
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_aa.json 2> gpurun_out/bench_aa.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_aa.json'));print(d['ms_per_step'], d['value'], d['roofline']['avg_iteration_us'], d['cg_iters_per_step'], d['optin_multigrid']['ms_per_step'])"
