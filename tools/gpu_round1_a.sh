set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload picflip1024 --verbose > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err; echo "bench1024 rc=$?"
cat gpurun_out/bench_1024.json
timeout 900 python bench.py --verbose > gpurun_out/bench_4096.json 2> gpurun_out/bench_4096.err; echo "bench4096 rc=$?"
cat gpurun_out/bench_4096.json
tail -3 gpurun_out/bench_4096.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_1024.csv python bench.py --workload picflip1024 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --cg-cap 60 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cg_ -s 40 -c 4 -o gpurun_out/prof_cg_4096 python bench.py --workload picflip4096 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --cg-cap 40 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
