# round 1, call s (2 GPUs): the sharded CG with the deferred x update -- parity tests, 2-rank check, 2-rank sweep, 2-GPU bench lines
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=3 > gpurun_out/pytest_gpu_2gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_cg_check.py --grid 1024 > gpurun_out/mgpu_check_1024.log 2>&1; echo "check rc=$?"
grep "^{" gpurun_out/mgpu_check_1024.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/cg_sweep.py --grids 4096,8192 --cap 1000 --only 1,31,4 --out gpurun_out/cg_sweep_2gpu.json > gpurun_out/cg_sweep2.log 2>&1; echo "sweep2 rc=$?"
grep "^{" gpurun_out/cg_sweep2.log | cut -c1-120
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_2gpu_picflip4096.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_2gpu_picflip4096.json'));print(d['ms_per_step'], d['cg_iters_per_step'], d['roofline']['avg_iteration_us'], d['e2e'])"
timeout 300 python bench.py --workload cg8192 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_1gpu_cg8192.json 2>/dev/null; echo "rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --workload cg8192 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_2gpu_cg8192.json 2>/dev/null; echo "rc=$?"
python -c "
import json
for f in ['bench_1gpu_cg8192','bench_2gpu_cg8192']:
    d=json.load(open('gpurun_out/%s.json'%f));print(f, d['ms_per_step'], d['cg_iters_per_step'], d['roofline']['avg_iteration_us'])"
