import sys, os, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from fluid_simulation_b200 import capi
import scenes
n=4096
dt=float(np.float32(0.01*64.0/n))
for skip in ("1","0"):
    os.environ["FSB_CG_SKIP_TILES"]=skip
    g=capi.Sim(n,n,1.0,1.0,dt,0.05)
    g.set_cg(2000,1e-6)
    d=1.0/n
    g.emit_source(2*d,0.35,2*d,1-2*d,1.25*d,1.25*d,0.0,0.0)
    g.step(capi.STEP_PICFLIP,dt)
    g.profile_enable(True); g.profile_read()
    g.step(capi.STEP_PICFLIP,dt)
    g.synchronize(); p=g.profile_read()
    it=g.cg_info()[0]
    print("dam-break 4096^2 skip",skip,"iters",it,"cg ms",p["cg"][0],"us/iter",1e3*p["cg"][0]/it, flush=True)
    g.close()
