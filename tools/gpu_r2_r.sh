#!/bin/bash
# round 2, call R (8 GPUs): slabs + sharded CG on 8 ranks against the single-GPU run, then the default bench line
# (4096^2 step, scale_cg8192, config 4 = 16384^2 with particle slabs)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29551 tests/multi_gpu_slab_check.py --backend nccl --grid 515 --steps 4 --shard-cg > gpurun_out/r_slab8.log 2>&1
echo "slabs + sharded CG, 8 ranks, 515^2 rc=$? $(grep '^{' gpurun_out/r_slab8.log | tail -1)"
timeout 900 $TR --master-port 29553 bench.py --gpus 8 --steps 3 --warmup 2 --no-cpu-baseline --no-optin --verbose > gpurun_out/r_default8.json 2> gpurun_out/r_default8.err
echo "bench rc=$?"; grep "config4:\|scale_cg8192\|rror" gpurun_out/r_default8.err | cut -c1-400 | tail -6
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r_default8.json"))
    print("default N=8: ms/step %.1f us/iter %.2f iters %.0f e2e %s" % (d["ms_per_step"], d["roofline"]["avg_iteration_us"], d["cg_iters_per_step"], d["e2e"]))
    print("scale:", json.dumps(d["scale_cg8192"])); print("config4:", json.dumps(d["config4_picflip16384"]))
except Exception as e: print("no line", e)
PY
