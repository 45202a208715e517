#!/bin/bash
# round 2, call D: is the one-sweep solve deterministic at 4096^2?  per-step iteration counts under several knobs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() {
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py "$@" --steps 4 --warmup 1 --no-cpu-baseline --no-optin --no-e2e --verbose > gpurun_out/d_$name.json 2> gpurun_out/d_$name.err
  echo "$name: $(grep -o 'step done, cg ([0-9]*, [0-9.e-]*)' gpurun_out/d_$name.err | tr '\n' ' ')"
}
run default X=1 -- --workload cg4096
run noearly FSB_CG_PREFETCH=0 -- --workload cg4096
run noearly_noserp FSB_CG_PREFETCH=0 FSB_CG_SERP=0 -- --workload cg4096
run noearly_xevery FSB_CG_PREFETCH=0 FSB_CG_XDEFER=0 -- --workload cg4096
run noearly_stages2 FSB_CG_PREFETCH=0 FSB_CG_STAGES=2 -- --workload cg4096
run noearly_1cta FSB_CG_PREFETCH=0 FSB_CG_CTAS_PER_SM=1 -- --workload cg4096
run noearly_noskip FSB_CG_PREFETCH=0 FSB_CG_SKIP_TILES=0 -- --workload cg4096
run fused FSB_CG_MODE=fused -- --workload cg4096
