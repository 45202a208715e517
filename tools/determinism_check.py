"""Debug: do repeated runs of the same capped solve give the same bits?  (tank scene)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from fluid_simulation_b200 import capi
n = int(os.environ.get("N", "4096"))
cap = int(os.environ.get("CAP", "40"))
reps = int(os.environ.get("REPS", "10"))
lab, u, v = bench.tank_fields(n)
dt = float(np.float32(0.01 * 64.0 / n))
g = capi.Sim(n, n, 1.0, 1.0, dt, 0.02)
g.set_cell_types(lab)
ref, res, bad = None, [], 0
for rep in range(reps):
    g.set_grid(capi.U_FRONT, u); g.set_grid(capi.V_FRONT, v)
    g.set_cg(cap, 1e-6)
    g.pressure_solve(dt, dt)
    x = g.get_pressure()
    res.append(g.cg_info()[1])
    if ref is None:
        ref = x
    else:
        bad += int((x != ref).any())
print(f"{os.environ.get('TAG','')}: {bad} of {reps - 1} repeats differ from the first; relres values {sorted(set(res))}")
