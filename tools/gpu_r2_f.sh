#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
cp fluid_simulation_b200/lib/libfsb.so /tmp/head.so
t() { tag=$1; shift; env TAG="$tag" "$@" timeout 200 python tools/gpu_r2_e.py 2>&1 | tail -1; }
for v in F0 F1; do
  cp build_variants/$v/libfsb.so fluid_simulation_b200/lib/libfsb.so
  t ${v}_noearly FSB_CG_PREFETCH=0
  t ${v}_default X=1
done
cp /tmp/head.so fluid_simulation_b200/lib/libfsb.so
t head_default REPS=20
t head_noearly FSB_CG_PREFETCH=0 REPS=20
t head_nofast FSB_CG_DEBUG_NOFAST=1
t head_noserp FSB_CG_SERP=0
t head_n8192 N=8192 REPS=6
