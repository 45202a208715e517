# round 1, call o (1 GPU): parity tests, default bench (both arms), mode sweep, ncu launch list + full captures
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/pytest_gpu.log
timeout 500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"
cat gpurun_out/bench_default.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?"
cat gpurun_out/bench_reference.json
timeout 300 python tools/cg_sweep.py --grids 1024,4096 --cap 2000 --only 1,4 --out gpurun_out/cg_sweep_1gpu.json > gpurun_out/cg_sweep.log 2>&1; echo "sweep rc=$?"
grep "^{" gpurun_out/cg_sweep.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_picflip4096.csv python bench.py --workload picflip4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --cg-cap 64 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_cg_solve" -s 1 -c 1 -o gpurun_out/prof_cg_solve_4096 python bench.py --workload cg4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --cg-cap 64 > gpurun_out/ncu_cg.log 2>&1; echo "ncu cg rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_g2p|k_p2g|k_sort|k_scan|k_mark|k_fill|k_extend|k_cg_build|k_pressure_patch|k_prev|k_enforce|k_place|k_count|k_advect|k_unperm|k_update_diff|k_add_acc" -c 40 -o gpurun_out/prof_stages_4096 python bench.py --workload picflip4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --cg-cap 32 > gpurun_out/ncu_stages.log 2>&1; echo "ncu stages rc=$?"
ls -la gpurun_out
