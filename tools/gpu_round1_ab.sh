# round 1, call ab (2 GPUs): sharded CG with the active-tile list; bench stdout hygiene under torchrun
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -q -x -k "sharded or active_tile or deferred or launch_modes" > gpurun_out/pytest_gpu_2gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_cg_check.py --grid 1024 > gpurun_out/mgpu_check_1024.log 2>&1; echo "check rc=$?"
grep "^{" gpurun_out/mgpu_check_1024.log
FSB_CG_MODE=fused timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 tests/multi_gpu_cg_check.py --grid 1024 > gpurun_out/mgpu_check_1024_fused.log 2>&1; echo "check fused rc=$?"
grep "^{" gpurun_out/mgpu_check_1024_fused.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_2gpu_ab.json 2> gpurun_out/bench_2gpu_ab.err; echo "bench2 rc=$?"
wc -l gpurun_out/bench_2gpu_ab.json
python -c "import json;d=json.load(open('gpurun_out/bench_2gpu_ab.json'));print(d['ms_per_step'], d['cg_iters_per_step'], d['roofline']['avg_iteration_us'], d['roofline']['kernel'][:50], d['gpu_launches'])"
