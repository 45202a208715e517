# round 1, call q (1 GPU): specialised particle kernels (compile-time delta fast path, lean fused G2P, register sort network) + L2 persistence window sweep
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=3 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
FSB_CG_VERBOSE=1 timeout 400 python tools/cg_sweep.py --grids 4096 --cap 2000 --only 4,20,21,22,23,24,12 --out gpurun_out/cg_sweep_persist_4096.json > gpurun_out/cg_sweep_persist.log 2>&1; echo "sweep rc=$?"
grep "^{\|fsb\]" gpurun_out/cg_sweep_persist.log | cut -c1-150 | sort | uniq | head -30
FSB_CG_VERBOSE=1 timeout 300 python tools/cg_sweep.py --grids 8192 --cap 600 --only 4,21,22 --out gpurun_out/cg_sweep_persist_8192.json > gpurun_out/cg_sweep_persist8.log 2>&1; echo "sweep8 rc=$?"
grep "^{" gpurun_out/cg_sweep_persist8.log | cut -c1-150
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_q.json'));print(d['ms_per_step'], d['stage_ms_per_step']); print(d['stage_roofline']); print(d['roofline']['avg_iteration_us'])"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_g2p|k_p2g|k_sort|k_scan|k_mark|k_fill|k_extend|k_cg_build|k_pressure_patch|k_prev|k_enforce" -s 18 -c 22 -o gpurun_out/prof_stages_4096_q python bench.py --workload picflip4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --cg-cap 32 > gpurun_out/ncu_stages_q.log 2>&1; echo "ncu stages rc=$?"
