#!/bin/bash
# round 2, call B: ncu --set full of the one-sweep solve kernel (4096^2, 64 iterations in one launch; 1024^2, 300)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cg_solve1 -s 1 -c 1 -o gpurun_out/b_one4096 \
  python bench.py --workload cg4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-optin --cg-cap 64 > gpurun_out/b_one4096.log 2>&1
echo "ncu one4096 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cg_solve1 -s 1 -c 1 -o gpurun_out/b_one1024 \
  python bench.py --workload cg1024 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-optin --cg-cap 300 > gpurun_out/b_one1024.log 2>&1
echo "ncu one1024 rc=$?"
ls -la gpurun_out/*.ncu-rep
