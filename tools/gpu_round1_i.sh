set -x
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_host_api.py -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_cg_check.py --grid 512 > gpurun_out/mgpu_512.log 2>&1; echo "mgpu512 rc=$?"
grep "^{" gpurun_out/mgpu_512.log
for wl in cg1024 cg4096; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_${wl}_n2_2k.json 2> gpurun_out/bench_${wl}_n2_2k.err; echo "rc=$?"
grep "^{" gpurun_out/bench_${wl}_n2_2k.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$wl n2 2k', d['roofline']['avg_iteration_us'], d['roofline']['frac'], d['cg_iters_per_step'], d['ms_per_step'])"
done
