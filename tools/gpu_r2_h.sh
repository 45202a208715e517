#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
rm -rf gpurun_out/dbg; mkdir -p gpurun_out/dbg
TAG=check CAP=40 REPS=8 FSB_CG_PREFETCH=0 FSB_CG_DEBUG_CHECK=1 FSB_CG_DEBUG_SUMS=gpurun_out/dbg/sums timeout 300 python tools/gpu_r2_e.py 2>&1 | tail -2
TAG=nocheck CAP=40 REPS=8 FSB_CG_PREFETCH=0 timeout 300 python tools/gpu_r2_e.py 2>&1 | tail -1
