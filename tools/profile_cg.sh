#!/bin/bash
# ncu captures behind profiles/rNN_*: the solve kernel (full set, with source), the launch list of one default
# step, the stage kernels.  One GPU:  gpurun --timeout 2400 -- 'bash tools/profile_cg.sh'
# Read the reports here with tools/ncu_summary.py and tools/ncu_sass_mix.py.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-optin --no-scale"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cg_solve1 -s 1 -c 1 -o gpurun_out/prof_cg4096 $B --workload cg4096 --cg-cap 64 > gpurun_out/prof_cg4096.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof_launches.csv $B --cg-cap 300 > gpurun_out/prof_launches.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'^(?!.*k_cg_solve1).*$' -s 12 -c 60 -o gpurun_out/prof_stages $B --cg-cap 300 > gpurun_out/prof_stages.log 2>&1; echo "rc=$?"
