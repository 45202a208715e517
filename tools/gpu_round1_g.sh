set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_cg_check.py --grid 512 > gpurun_out/mgpu_512.log 2>&1; echo "mgpu512 rc=$?"
grep "^{" gpurun_out/mgpu_512.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 tests/multi_gpu_cg_check.py --grid 4096 > gpurun_out/mgpu_4096.log 2>&1; echo "mgpu4096 rc=$?"
grep "^{" gpurun_out/mgpu_4096.log
timeout 600 python bench.py --workload cg4096 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cg4096_n1.json 2> gpurun_out/bench_cg4096_n1.err; echo "rc=$?"
cat gpurun_out/bench_cg4096_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --workload cg4096 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cg4096_n2.json 2> gpurun_out/bench_cg4096_n2.err; echo "rc=$?"
grep "^{" gpurun_out/bench_cg4096_n2.json
tail -3 gpurun_out/bench_cg4096_n2.err
