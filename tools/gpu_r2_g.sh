#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export CAP=4 REPS=2 TAG=san
for tool in memcheck initcheck racecheck synccheck; do
  echo "== $tool"
  timeout 600 compute-sanitizer --tool $tool --kernel-regex kns=k_cg_solve1 --print-limit 20 python tools/gpu_r2_e.py > gpurun_out/g_$tool.log 2>&1
  grep -v "^=========$" gpurun_out/g_$tool.log | tail -25
done
