set -x
mkdir -p gpurun_out
FSB_CG_VERBOSE=1 timeout 200 python tools/cg_sweep.py --grids 4096 --cap 500 --only 4 2>&1 | grep "fsb\]\|^{" | cut -c1-160 | head -6
FSB_CG_VERBOSE=1 timeout 300 python tools/dam_break_tile_skip.py 2>&1 | grep "fsb\]\|dam-break" | cut -c1-160 | head -12
