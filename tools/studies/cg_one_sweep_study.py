"""One-sweep Jacobi-PCG study (numpy fp32): the standard iteration (Eigen's statement order) against
the one-sweep form whose beta comes from the exact identity
    r'.z' = r.z - 2 alpha (z.q) + alpha^2 (q.D^-1 q),   r' = r - alpha q, z = D^-1 r, q = A p
evaluated in double from dot products of the vectors of the PREVIOUS iteration (all available in the
sweep that formed them), while alpha and the stopping test keep using the exact r.z and |r|^2 of the
actual vectors.  One reduction point per iteration, no extra vectors."""
import numpy as np, sys, time, warnings
warnings.filterwarnings("ignore")
from mgpcg_prototype import *
from cg_single_reduction_study import jacobi_pcg
f64=np.float64
def dot(a,b): return float((a.astype(f64)*b.astype(f64)).sum())
def cg_one_sweep(L,b,h2,dinv,tol=1e-6,maxit=200000):
    x=np.zeros_like(b); r=b.copy()
    rhs2=dot(b,b); thr=tol*tol*rhs2
    z=(dinv*r).astype(f32); p=z.copy(); q=applyA(L,p,h2)
    rz=dot(r,z); pq=dot(p,q); zq=dot(z,q); qmq=dot(q,(dinv*q).astype(f32))  # reduction of the init sweep
    it=0
    while it<maxit:
        alpha=f32(rz/pq)
        a=float(alpha)
        rz_pred=rz-2*a*zq+a*a*qmq
        beta=f32(rz_pred/rz)
        # ---- the sweep: everything below uses only alpha, beta and the old vectors
        x=(x+alpha*p).astype(f32); r=(r-alpha*q).astype(f32)
        z=(dinv*r).astype(f32); p=(z+beta*p).astype(f32); q=applyA(L,p,h2)
        r2=dot(r,r); rz_new=dot(r,z); pq=dot(p,q); zq=dot(z,q); qmq=dot(q,(dinv*q).astype(f32))
        # ---- reduction point
        if r2<thr: break
        rz=rz_new; it+=1
    return x,it+1,np.sqrt(r2/rhs2)
if __name__=="__main__":
    for n in [int(a) for a in sys.argv[1:]] or (256,512,1024):
        lab,u,v=tank(n); dx=f32(1)/f32(n); b=rhs_from(lab,u,v,dx); L=make_level(lab); h2=f32(1)/(dx*dx)
        dinv=np.where(L['cnt']>0,f32(-1)/(np.maximum(L['cnt'],1)*h2),f32(0)).astype(f32)
        xa,ia,ea=jacobi_pcg(L,b,h2,dinv)
        xb,ib,eb=cg_one_sweep(L,b,h2,dinv)
        ra=np.linalg.norm((b-applyA(L,xa,h2)).astype(f64))/np.linalg.norm(b.astype(f64))
        rb=np.linalg.norm((b-applyA(L,xb,h2)).astype(f64))/np.linalg.norm(b.astype(f64))
        print(n,"standard iters",ia,"relres",float(ea),"true",ra,"| one-sweep iters",ib,"relres",float(eb),"true",rb,
              "| rel diff",np.linalg.norm(xa.astype(f64)-xb)/np.linalg.norm(xa.astype(f64)),flush=True)
