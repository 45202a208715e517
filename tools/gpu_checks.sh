#!/bin/bash
# The GPU checks of a change, one gpurun call (every command under its own timeout):
#   gpurun --timeout 2400 -- 'bash tools/gpu_checks.sh'            1 GPU: tests, default bench line, CG timings
#   gpurun --gpus 2 --timeout 1800 -- 'bash tools/gpu_checks.sh 2'  2 GPUs: sharded CG + particle slabs
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_checks.sh 8'  8 GPUs: the same + the 8-GPU bench line
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
N=${1:-1}
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d["roofline"]
    print(f"  {sys.argv[1]}: {d['ms_per_step']:.1f} ms/step, {d['cg_iters_per_step']:.0f} CG iterations at "
          f"{r['avg_iteration_us']:.2f} us, frac {r['frac']:.3f}; scale_cg8192 {d.get('scale_cg8192')}")
except Exception as e:
    print("  no line in", sys.argv[1], e)
PY
}
if [ "$N" = 1 ]; then
  timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/checks_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/checks_pytest.log
  timeout 900 python bench.py --no-cpu-baseline > gpurun_out/checks_bench_n1.json 2> gpurun_out/checks_bench_n1.err; line gpurun_out/checks_bench_n1.json
  timeout 400 python tools/stage_knobs.py > gpurun_out/checks_stage_knobs.log 2>&1; echo "stage knobs rc=$?"; tail -1 gpurun_out/checks_stage_knobs.log
  for wl in cg4096 cg1024 cg8192; do
    timeout 300 python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-optin --no-e2e --no-scale > gpurun_out/checks_$wl.json 2> gpurun_out/checks_$wl.err; line gpurun_out/checks_$wl.json
  done
  TAG=determinism REPS=12 timeout 200 python tools/determinism_check.py 2>&1 | tail -1
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 300 $TR --master-port 29541 tests/multi_gpu_cg_check.py --grid 1030 > gpurun_out/checks_cg_$N.log 2>&1
  echo "sharded CG vs single GPU rc=$? $(grep '^{' gpurun_out/checks_cg_$N.log | tail -1 | cut -c1-300)"
  timeout 300 $TR --master-port 29551 tests/multi_gpu_slab_check.py --backend nccl --grid 515 --steps 4 --shard-cg > gpurun_out/checks_slab_$N.log 2>&1
  echo "particle slabs + sharded CG vs single GPU rc=$? $(grep '^{' gpurun_out/checks_slab_$N.log | tail -1)"
  timeout 200 $TR --master-port 29552 tests/multi_gpu_slab_check.py --backend nccl --grid 515 --steps 4 --kind sl --shard-cg > gpurun_out/checks_slab_sl_$N.log 2>&1
  echo "semi-Lagrangian slabs + sharded CG vs single GPU rc=$? $(grep '^{' gpurun_out/checks_slab_sl_$N.log | tail -1)"
  timeout 900 $TR --master-port 29553 bench.py --gpus $N --no-cpu-baseline --no-optin > gpurun_out/checks_bench_n$N.json 2> gpurun_out/checks_bench_n$N.err; line gpurun_out/checks_bench_n$N.json
fi
