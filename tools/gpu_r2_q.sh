#!/bin/bash
# round 2, call Q (2 GPUs): particle slabs over NCCL (device buffers), with and without the sharded CG
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/q_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/q_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29551 tests/multi_gpu_slab_check.py --backend nccl --grid 131 --steps 6 --shard-cg > gpurun_out/q_slab_shard.log 2>&1
echo "slabs + sharded CG (131^2, uneven slabs) rc=$? $(grep '^{' gpurun_out/q_slab_shard.log | tail -1)"
timeout 300 $TR --master-port 29552 tests/multi_gpu_slab_check.py --backend nccl --grid 512 --steps 4 --shard-cg > gpurun_out/q_slab_shard512.log 2>&1
echo "slabs + sharded CG 512 rc=$? $(grep '^{' gpurun_out/q_slab_shard512.log | tail -1)"
timeout 600 $TR --master-port 29553 bench.py --gpus 2 --workload picflip1024 --steps 1 --warmup 1 --no-cpu-baseline --no-optin --no-e2e --config4 4096 --verbose > gpurun_out/q_config4.json 2> gpurun_out/q_config4.err
echo "bench config4@4096 rc=$?"; grep "config4\|error" gpurun_out/q_config4.err | tail -5
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/q_config4.json")); print(json.dumps(d["config4_picflip16384"])[:1500])
except Exception as e: print("no line", e)
PY
