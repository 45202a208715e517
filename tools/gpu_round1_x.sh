# round 1, call x (1 GPU): frame rasteriser + bench stdout hygiene
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload picflip1024 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_x_1024.out 2> gpurun_out/bench_x_1024.err; echo "bench rc=$?"
wc -l gpurun_out/bench_x_1024.out; cut -c1-200 gpurun_out/bench_x_1024.out
