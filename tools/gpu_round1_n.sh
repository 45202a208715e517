# round 1, call n (2 GPUs): parity tests incl. the 2-rank check, CG mode sweep on 2 row slabs
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/cg_sweep.py --grids 4096,8192 --cap 1000 --only 0,1,2,4,6 --out gpurun_out/cg_sweep_2gpu.json > gpurun_out/cg_sweep2.log 2>&1; echo "sweep2 rc=$?"
grep "^{" gpurun_out/cg_sweep2.log
tail -5 gpurun_out/cg_sweep2.log
timeout 300 python tools/cg_sweep.py --grids 8192 --cap 1000 --only 1,4 --out gpurun_out/cg_sweep_8192_1gpu.json > gpurun_out/cg_sweep1.log 2>&1; echo "sweep1 rc=$?"
grep "^{" gpurun_out/cg_sweep1.log
