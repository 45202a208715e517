set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "pressure_solve or config0" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer.log | tail -3
for th in 8 16 32; do
  FSB_CG_TILE_ROWS=$th timeout 600 python bench.py --workload picflip1024 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_1024_th$th.json 2>gpurun_out/bench_1024_th$th.err; echo "rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_1024_th$th.json'));print('1024 th$th', d['roofline']['avg_iteration_us'], d['cg_iters_per_step'], d['ms_per_step'])"
done
for th in 16 32; do
  FSB_CG_TILE_ROWS=$th timeout 600 python bench.py --workload picflip4096 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_4096_th$th.json 2>gpurun_out/bench_4096_th$th.err; echo "rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_4096_th$th.json'));print('4096 th$th', d['roofline']['avg_iteration_us'], d['roofline']['frac'], d['cg_iters_per_step'], d['ms_per_step'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cg_ -s 40 -c 4 -o gpurun_out/prof_cg_4096_b python bench.py --workload picflip4096 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --cg-cap 64 > gpurun_out/ncu_full_b.log 2>&1; echo "ncu full rc=$?"
