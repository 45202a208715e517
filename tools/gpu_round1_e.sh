set -x
timeout 180 python -m pytest tests/test_gpu_parity.py -x -q -k "pressure" > gpurun_out/pytest_pressure.log 2>&1; rc=$?; echo "pressure rc=$rc"
tail -15 gpurun_out/pytest_pressure.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
for cfg in "16 2" "16 1" "32 1" "8 4"; do
  set -- $cfg
  FSB_CG_TILE_ROWS=$1 FSB_CG_CTAS_PER_SM=$2 timeout 600 python bench.py --workload picflip4096 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_4096_th$1_c$2.json 2>gpurun_out/bench_4096_th$1_c$2.err; echo "rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_4096_th$1_c$2.json'));print('4096 th$1 ctas$2', d['roofline']['avg_iteration_us'], d['roofline']['frac'], d['cg_iters_per_step'], d['ms_per_step'], d['stage_ms_per_step'])"
done
for cfg in "16 2" "8 4" "8 2"; do
  set -- $cfg
  FSB_CG_TILE_ROWS=$1 FSB_CG_CTAS_PER_SM=$2 timeout 600 python bench.py --workload picflip1024 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_1024_th$1_c$2.json 2>gpurun_out/bench_1024_th$1_c$2.err; echo "rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_1024_th$1_c$2.json'));print('1024 th$1 ctas$2', d['roofline']['avg_iteration_us'], d['cg_iters_per_step'], d['ms_per_step'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_cg_|k_p2g" -s 10 -c 5 -o gpurun_out/prof_cg_4096_e python bench.py --workload picflip4096 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --cg-cap 64 > gpurun_out/ncu_full_e.log 2>&1; echo "ncu full rc=$?"
