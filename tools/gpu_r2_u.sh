#!/bin/bash
# round 2, call U (1 GPU): ncu captures for profiles/ -- the one-sweep solve kernel (full set, source), the launch
# list of one default step, the stage kernels (full set)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cg_solve1 -s 1 -c 1 -o gpurun_out/u_one4096 \
  python bench.py --workload cg4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-optin --no-scale --cg-cap 64 > gpurun_out/u_one4096.log 2>&1
echo "ncu one4096 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/u_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-optin --no-scale --cg-cap 300 > gpurun_out/u_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'^(?!.*k_cg_solve1).*$' -s 30 -c 40 -o gpurun_out/u_stages \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-optin --no-scale --cg-cap 300 > gpurun_out/u_stages.log 2>&1
echo "ncu stages rc=$?"
ls -la gpurun_out/u_*
