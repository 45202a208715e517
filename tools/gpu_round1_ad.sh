# round 1, call ad (2 GPUs): slab boundary tiles first -- bit-identity check, then the 2-rank sweep with and without
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_cg_check.py --grid 1024 > gpurun_out/mgpu_check_1024.log 2>&1; echo "check rc=$?"
grep "^{" gpurun_out/mgpu_check_1024.log | cut -c1-330
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/cg_sweep.py --grids 4096,8192 --cap 1000 --only 1,35,4,36 --out gpurun_out/cg_sweep_2gpu_edge.json > gpurun_out/cg_sweep2e.log 2>&1; echo "sweep2 rc=$?"
grep "^{" gpurun_out/cg_sweep2e.log | cut -c1-120
