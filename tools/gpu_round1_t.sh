# round 1, call t (1 GPU): new ABI entries (uncalled routines, state files), final r01h bench lines (both arms), SL workload, 16384^2 / 8192^2 device-emitted workloads on one GPU
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=3 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_t_default.json 2> gpurun_out/bench_t_default.err; echo "bench default rc=$?"
cat gpurun_out/bench_t_default.json | cut -c1-400
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_t_reference.json 2> /dev/null; echo "bench reference rc=$?"
timeout 300 python bench.py --workload sl1024 --no-cpu-baseline > gpurun_out/bench_t_sl1024.json 2> gpurun_out/bench_t_sl1024.err; echo "bench sl rc=$?"
cat gpurun_out/bench_t_sl1024.json | cut -c1-900
timeout 300 python bench.py --workload picflip8192e --steps 1 --warmup 1 --no-cpu-baseline --cg-cap 300 --verbose > gpurun_out/bench_t_8192e.json 2> gpurun_out/bench_t_8192e.err; echo "bench 8192e rc=$?"
tail -3 gpurun_out/bench_t_8192e.err; cat gpurun_out/bench_t_8192e.json | cut -c1-700
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
timeout 400 python bench.py --workload picflip16384 --steps 1 --warmup 1 --no-cpu-baseline --cg-cap 300 --verbose > gpurun_out/bench_t_16384.json 2> gpurun_out/bench_t_16384.err; echo "bench 16384 rc=$?"
tail -4 gpurun_out/bench_t_16384.err; cat gpurun_out/bench_t_16384.json | cut -c1-900
