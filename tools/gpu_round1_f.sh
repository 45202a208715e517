set -x
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "pressure or config0" > gpurun_out/pytest_pressure.log 2>&1; rc=$?; echo "pressure rc=$rc"
tail -5 gpurun_out/pytest_pressure.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_cg_check.py --grid 512 > gpurun_out/mgpu_512.log 2>&1; echo "mgpu512 rc=$?"
tail -5 gpurun_out/mgpu_512.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 tests/multi_gpu_cg_check.py --grid 4096 > gpurun_out/mgpu_4096.log 2>&1; echo "mgpu4096 rc=$?"
tail -3 gpurun_out/mgpu_4096.log
