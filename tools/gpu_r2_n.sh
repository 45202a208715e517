#!/bin/bash
# round 2, call N (8 GPUs): the sharded one-sweep CG on 8 ranks -- correctness against the single-GPU solve, the
# default bench line (4096^2 full step + the 8192^2 side measurement), the CG-only workloads
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${NG:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tests/multi_gpu_cg_check.py --grid 2048 > gpurun_out/n_check_2048.log 2>&1
echo "check 2048 rc=$? $(grep '^{' gpurun_out/n_check_2048.log | tail -1)"
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 400 $TR --master-port 29542 bench.py --gpus $N "$@" --no-cpu-baseline --no-optin > gpurun_out/n_$name.json 2> gpurun_out/n_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/n_$name.json"))
    print("$name: iters", d["cg_iters_per_step"], "us/iter %.2f" % d["roofline"]["avg_iteration_us"], "relres", d["cg_relres"], "ms/step %.1f" % d["ms_per_step"], "scale", d.get("scale_cg8192"))
except Exception as e:
    print("$name: no line:", e); print(open("gpurun_out/n_$name.err").read()[-1200:])
PY
}
run default X=1 -- --steps 3 --warmup 2
run cg4096 X=1 -- --workload cg4096 --steps 2 --warmup 1 --no-e2e --no-scale
run cg1024_notiles FSB_CG_DEBUG_NOTILES=1 -- --workload cg1024 --cg-cap 3000 --steps 2 --warmup 1 --no-e2e --no-scale
