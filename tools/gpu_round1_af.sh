# round 1, call af (2 GPUs): particle slabs over NCCL, alone and together with the sharded CG
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/pytest_gpu_2gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_slab_check.py --backend nccl --grid 512 --steps 4 --shard-cg > gpurun_out/slab_check_nccl_cg.log 2>&1; echo "slab+cg rc=$?"
grep "^{" gpurun_out/slab_check_nccl_cg.log; tail -3 gpurun_out/slab_check_nccl_cg.log | cut -c1-300
