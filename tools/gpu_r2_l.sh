#!/bin/bash
# round 2, call L (2 GPUs): cost of the cross-GPU reduction of the one-sweep solve
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 $TR --master-port 29542 bench.py --gpus 2 "$@" --warmup 1 --steps 2 --no-cpu-baseline --no-optin --no-e2e > gpurun_out/l_$name.json 2> gpurun_out/l_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/l_$name.json"))
    print("$name: iters", d["cg_iters_per_step"], "us/iter %.2f" % d["roofline"]["avg_iteration_us"], "relres", d["cg_relres"], "ms/step %.1f" % d["ms_per_step"])
except Exception as e:
    print("$name: no line:", e); print(open("gpurun_out/l_$name.err").read()[-600:])
PY
}
run notiles_sys_sys FSB_CG_DEBUG_NOTILES=1 -- --workload cg1024 --cg-cap 3000
run notiles_pollgpu FSB_CG_DEBUG_NOTILES=1 FSB_CG_POLL_FENCE_GPU=1 -- --workload cg1024 --cg-cap 3000
run notiles_bothgpu FSB_CG_DEBUG_NOTILES=1 FSB_CG_POLL_FENCE_GPU=1 FSB_CG_POST_FENCE_GPU=1 -- --workload cg1024 --cg-cap 3000
run notiles_bothgpu_1cta FSB_CG_DEBUG_NOTILES=1 FSB_CG_POLL_FENCE_GPU=1 FSB_CG_POST_FENCE_GPU=1 FSB_CG_CTAS_PER_SM=1 -- --workload cg1024 --cg-cap 3000
run cg1024_sys X=1 -- --workload cg1024
run cg1024_bothgpu FSB_CG_POLL_FENCE_GPU=1 FSB_CG_POST_FENCE_GPU=1 -- --workload cg1024
run cg4096_bothgpu FSB_CG_POLL_FENCE_GPU=1 FSB_CG_POST_FENCE_GPU=1 -- --workload cg4096
run cg8192_bothgpu FSB_CG_POLL_FENCE_GPU=1 FSB_CG_POST_FENCE_GPU=1 -- --workload cg8192
timeout 200 python bench.py --workload cg1024 --cg-cap 3000 --warmup 1 --steps 2 --no-cpu-baseline --no-optin --no-e2e > gpurun_out/l_1gpu_notiles.json 2>/dev/null; FSB_CG_DEBUG_NOTILES=1 timeout 200 python bench.py --workload cg1024 --cg-cap 3000 --warmup 1 --steps 2 --no-cpu-baseline --no-optin --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1gpu notiles us/iter %.2f' % d['roofline']['avg_iteration_us'])"
timeout 200 python bench.py --workload cg1024 --warmup 1 --steps 2 --no-cpu-baseline --no-optin --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1gpu cg1024 us/iter %.2f' % d['roofline']['avg_iteration_us'])"
