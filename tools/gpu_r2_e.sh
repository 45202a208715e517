#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
t() { tag=$1; shift; env TAG="$tag" "$@" timeout 200 python tools/gpu_r2_e.py 2>&1 | tail -1; }
t default X=1
t noearly FSB_CG_PREFETCH=0
t noearly_nofast FSB_CG_PREFETCH=0 FSB_CG_DEBUG_NOFAST=1
t noearly_xevery FSB_CG_PREFETCH=0 FSB_CG_XDEFER=0
t noearly_stages2 FSB_CG_PREFETCH=0 FSB_CG_STAGES=2
t noearly_stages3 FSB_CG_PREFETCH=0 FSB_CG_STAGES=3
t noearly_stages4 FSB_CG_PREFETCH=0 FSB_CG_STAGES=4
t noearly_1cta FSB_CG_PREFETCH=0 FSB_CG_CTAS_PER_SM=1
t noearly_rows8 FSB_CG_PREFETCH=0 FSB_CG_TILE_ROWS=8
t noearly_rows32 FSB_CG_PREFETCH=0 FSB_CG_TILE_ROWS=32
t noearly_noskip FSB_CG_PREFETCH=0 FSB_CG_SKIP_TILES=0
t fused FSB_CG_MODE=fused
t n2048_default N=2048
t n1024_default N=1024
