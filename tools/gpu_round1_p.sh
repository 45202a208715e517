# round 1, call p (1 GPU): vectorised stage kernels -> parity tests, default bench, L2-hint sweep of the CG, ncu of the new stage kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log
timeout 400 python tools/cg_sweep.py --grids 4096 --cap 2000 --only 4,5,7,8,9,10,11,12,13,14,15,16,17,18,19 --out gpurun_out/cg_sweep_hints_4096.json > gpurun_out/cg_sweep_hints.log 2>&1; echo "sweep rc=$?"
grep "^{" gpurun_out/cg_sweep_hints.log | cut -c1-140
timeout 300 python tools/cg_sweep.py --grids 8192 --cap 600 --only 4,9,11,12,13 --out gpurun_out/cg_sweep_hints_8192.json > gpurun_out/cg_sweep_hints8.log 2>&1; echo "sweep8 rc=$?"
grep "^{" gpurun_out/cg_sweep_hints8.log | cut -c1-140
timeout 300 python tools/cg_sweep.py --grids 1024,2048 --cap 2000 --only 4,9,13 --out gpurun_out/cg_sweep_hints_small.json > gpurun_out/cg_sweep_hints_s.log 2>&1; echo "sweeps rc=$?"
grep "^{" gpurun_out/cg_sweep_hints_s.log | cut -c1-140
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_p.json'));print(d['ms_per_step'], d['stage_ms_per_step']); print(d['stage_roofline']); print(d['roofline']['avg_iteration_us'])"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_g2p|k_p2g|k_sort|k_scan|k_mark|k_fill|k_extend|k_cg_build|k_pressure_patch|k_prev|k_enforce" -s 18 -c 22 -o gpurun_out/prof_stages_4096_p python bench.py --workload picflip4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --cg-cap 32 > gpurun_out/ncu_stages_p.log 2>&1; echo "ncu stages rc=$?"
ls -la gpurun_out | tail -8
