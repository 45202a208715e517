import numpy as np, sys, time, warnings
warnings.filterwarnings("ignore")
from mgpcg_prototype import *
def jacobi_pcg(L,b,h2,dinv,tol=1e-6,maxit=200000):
    x=np.zeros_like(b); r=b.copy()
    rhs2=float((b.astype(np.float64)**2).sum()); thr=tol*tol*rhs2
    z=(dinv*r).astype(f32); p=z.copy(); absNew=float((r.astype(np.float64)*z).sum()); it=0
    while it<maxit:
        q=applyA(L,p,h2)
        alpha=f32(absNew/float((p.astype(np.float64)*q).sum()))
        x=(x+alpha*p).astype(f32); r=(r-alpha*q).astype(f32)
        r2=float((r.astype(np.float64)**2).sum())
        if r2<thr: break
        z=(dinv*r).astype(f32); absOld=absNew; absNew=float((r.astype(np.float64)*z).sum())
        p=(z+f32(absNew/absOld)*p).astype(f32); it+=1
    return x,it+1,np.sqrt(r2/rhs2)
def cg_single_reduction(L,b,h2,dinv,tol=1e-6,maxit=200000):
    """Chronopoulos-Gear preconditioned CG: one reduction point per iteration (gamma=r.z, delta=z.Az, |r|^2),
    s = A p kept by recurrence s = w + beta s with w = A z."""
    x=np.zeros_like(b); r=b.copy()
    rhs2=float((b.astype(np.float64)**2).sum()); thr=tol*tol*rhs2
    z=(dinv*r).astype(f32); w=applyA(L,z,h2)
    gamma=float((r.astype(np.float64)*z).sum()); delta=float((z.astype(np.float64)*w).sum())
    alpha=f32(gamma/delta); beta=f32(0)
    p=np.zeros_like(b); s=np.zeros_like(b); it=0
    while it<maxit:
        p=(z+beta*p).astype(f32); s=(w+beta*s).astype(f32)
        x=(x+alpha*p).astype(f32); r=(r-alpha*s).astype(f32)
        z=(dinv*r).astype(f32); w=applyA(L,z,h2)
        # ---- single reduction point
        gnew=float((r.astype(np.float64)*z).sum()); delta=float((z.astype(np.float64)*w).sum())
        r2=float((r.astype(np.float64)**2).sum())
        it+=1
        if r2<thr: break
        beta=f32(gnew/gamma); gamma=gnew
        alpha=f32(gamma/(delta-float(beta)*gamma/float(alpha)))
    return x,it,np.sqrt(r2/rhs2)
if __name__=="__main__":
  for n in (256,512,1024):
      lab,u,v=tank(n); dx=f32(1)/f32(n); b=rhs_from(lab,u,v,dx); L=make_level(lab); h2=f32(1)/(dx*dx)
      dinv=np.where(L['cnt']>0,f32(-1)/(np.maximum(L['cnt'],1)*h2),f32(0)).astype(f32)
      t=time.time(); xa,ia,ea=jacobi_pcg(L,b,h2,dinv); ta=time.time()-t
      t=time.time(); xb,ib,eb=cg_single_reduction(L,b,h2,dinv); tb=time.time()-t
      # true residuals
      ra=np.linalg.norm((b-applyA(L,xa,h2)).astype(np.float64))/np.linalg.norm(b.astype(np.float64))
      rb=np.linalg.norm((b-applyA(L,xb,h2)).astype(np.float64))/np.linalg.norm(b.astype(np.float64))
      print(n,"standard iters",ia,"relres",float(ea),"true",ra,"| single-reduction iters",ib,"relres",float(eb),"true",rb,
            "| rel diff",np.linalg.norm(xa.astype(np.float64)-xb)/np.linalg.norm(xa.astype(np.float64)),flush=True)
