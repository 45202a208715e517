# round 1, call m: parity tests, CG mode sweep (new persistent kernel), ncu of the non-CG stage kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/cg_sweep.py --grids 1024,4096 --cap 2000 --out gpurun_out/cg_sweep_1gpu.json > gpurun_out/cg_sweep.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/cg_sweep.log | tail -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_g2p|k_p2g|k_sort|k_scan|k_mark|k_fill|k_extend|k_cg_build|k_pressure_patch|k_prev|k_enforce" -s 18 -c 18 -o gpurun_out/prof_stages_4096 python bench.py --workload picflip4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --cg-cap 32 > gpurun_out/ncu_stages.log 2>&1; echo "ncu stages rc=$?"
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"
cat gpurun_out/bench_default.json
