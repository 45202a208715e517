# round 1, call r (1 GPU): deferred x update in the persistent CG kernel; L2 persistence sweep (miss property, sizes)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=3 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 400 python tools/cg_sweep.py --grids 4096 --cap 2000 --only 1,31,4,34,33,24,25,26,27,28,29,30,32,4 --out gpurun_out/cg_sweep_r_4096.json > gpurun_out/cg_sweep_r.log 2>&1; echo "sweep rc=$?"
grep "^{" gpurun_out/cg_sweep_r.log | cut -c1-110
timeout 300 python tools/cg_sweep.py --grids 8192 --cap 600 --only 31,4,34 --out gpurun_out/cg_sweep_r_8192.json > gpurun_out/cg_sweep_r8.log 2>&1; echo "sweep8 rc=$?"
grep "^{" gpurun_out/cg_sweep_r8.log | cut -c1-110
timeout 300 python tools/cg_sweep.py --grids 1024,2048 --cap 2000 --only 31,4,28 --out gpurun_out/cg_sweep_r_small.json > gpurun_out/cg_sweep_rs.log 2>&1; echo "sweeps rc=$?"
grep "^{" gpurun_out/cg_sweep_rs.log | cut -c1-110
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_r.json'));print(d['ms_per_step'], d['stage_ms_per_step']); print(d['stage_roofline']); print(d['roofline']['avg_iteration_us'])"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_cg_solve" -s 1 -c 1 -o gpurun_out/prof_cg_solve_4096_r python bench.py --workload cg4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --cg-cap 64 > gpurun_out/ncu_cg_r.log 2>&1; echo "ncu cg rc=$?"
