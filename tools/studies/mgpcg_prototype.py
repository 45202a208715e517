import numpy as np, sys, time
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import scenes
LIQ,AIR,SOL=0,1,2
f32=np.float32

def make_level(lab):
    ny,nx=lab.shape
    liq=(lab==LIQ)
    nons=(lab!=SOL)
    pad=np.pad(nons,1,mode='constant',constant_values=False)  # outside = solid
    cnt=(pad[1:-1,:-2].astype(np.int32)+pad[1:-1,2:]+pad[:-2,1:-1]+pad[2:,1:-1])
    cnt=np.where(liq,cnt,0)
    return dict(lab=lab,liq=liq,cnt=cnt.astype(f32),ny=ny,nx=nx)

def coarsen(lab):
    ny,nx=lab.shape
    NY,NX=(ny+1)//2,(nx+1)//2
    p=np.full((NY*2,NX*2),SOL,dtype=np.uint8); p[:ny,:nx]=lab
    ch=np.stack([p[0::2,0::2],p[0::2,1::2],p[1::2,0::2],p[1::2,1::2]])
    anyair=(ch==AIR).any(0); anyliq=(ch==LIQ).any(0)
    out=np.full((NY,NX),SOL,dtype=np.uint8)
    out[anyliq]=LIQ
    out[anyair]=AIR
    return out

def applyA(L,x,inv_h2):
    # x is zero outside liquid
    xp=np.pad(x,1)
    s=xp[1:-1,:-2]+xp[1:-1,2:]+xp[:-2,1:-1]+xp[2:,1:-1]
    return np.where(L['liq'],(s-L['cnt']*x)*inv_h2,f32(0)).astype(f32)

def smooth(L,x,b,inv_h2,n,omega=f32(2/3)):
    # damped jacobi: x += omega * Dinv (b - A x), D = -cnt*inv_h2
    dinv=np.where(L['cnt']>0,f32(-1)/(np.maximum(L['cnt'],1)*inv_h2),f32(0)).astype(f32)
    for _ in range(n):
        r=b-applyA(L,x,inv_h2)
        x=(x+omega*dinv*r).astype(f32)
        x=np.where(L['liq'],x,f32(0))
    return x

W=np.array([1,3,3,1],dtype=f32)/f32(8)
def restrict(Lf,Lc,r):
    ny,nx=r.shape; NY,NX=Lc['ny'],Lc['nx']
    # coarse (I,J) <- fine rows 2J-1..2J+2, cols 2I-1..2I+2 with weights W x W
    rp=np.zeros((2*NY+2,2*NX+2),dtype=f32); rp[1:ny+1,1:nx+1]=r
    out=np.zeros((NY,NX),dtype=f32)
    for a in range(4):
        for c in range(4):
            out+=W[a]*W[c]*rp[a:a+2*NY:2,c:c+2*NX:2]
    return np.where(Lc['liq'],out,f32(0)).astype(f32)

def prolong(Lf,Lc,e):
    ny,nx=Lf['ny'],Lf['nx']; NY,NX=Lc['ny'],Lc['nx']
    out=np.zeros((2*NY+2,2*NX+2),dtype=f32)
    for a in range(4):
        for c in range(4):
            out[a:a+2*NY:2,c:c+2*NX:2]+=f32(4)*W[a]*W[c]*e
    out=out[1:ny+1,1:nx+1]
    return np.where(Lf['liq'],out,f32(0)).astype(f32)

def parent_norm(Lf,Lc):
    """Sum of a fine cell's bilinear weights over its non-SOLID coarse parents (outside the grid = SOLID):
    1 away from walls, < 1 next to them.  Dividing by it makes the prolongation a constant extension across
    walls and its transpose, the restriction, conservative there (tools/studies/mg_transfer_study.py)."""
    nons=(Lc['lab']!=SOL).astype(f32)
    ny,nx=Lf['ny'],Lf['nx']; NY,NX=Lc['ny'],Lc['nx']
    out=np.zeros((2*NY+2,2*NX+2),dtype=f32)
    for a in range(4):
        for c in range(4):
            out[a:a+2*NY:2,c:c+2*NX:2]+=f32(4)*W[a]*W[c]*nons
    w=out[1:ny+1,1:nx+1]
    return np.where(w>0,w,f32(1)).astype(f32)

class MG:
    def __init__(self,lab,dx,nmin=8,pre=2,post=2,coarse_sweeps=40,scale=1.0,renorm=False):
        self.levels=[make_level(lab)]; self.h2=[f32(1)/(f32(dx)*f32(dx))]
        while max(self.levels[-1]['lab'].shape)>nmin:
            cl=coarsen(self.levels[-1]['lab'])
            self.levels.append(make_level(cl)); self.h2.append(self.h2[-1]/f32(4))
        self.pre,self.post,self.cs=pre,post,coarse_sweeps
        self.scale=f32(scale)
        self.norm=[parent_norm(self.levels[l],self.levels[l+1]) if renorm else None for l in range(len(self.levels)-1)]
    def vcycle(self,b,l=0):
        L=self.levels[l]; h2=self.h2[l]
        x=np.zeros_like(b)
        if l==len(self.levels)-1:
            return smooth(L,x,b,h2,self.cs)
        x=smooth(L,x,b,h2,self.pre)
        r=b-applyA(L,x,h2); r=np.where(L['liq'],r,f32(0))
        if self.norm[l] is not None: r=(r/self.norm[l]).astype(f32)
        rc=restrict(L,self.levels[l+1],r)
        ec=self.vcycle(rc,l+1)
        e=prolong(L,self.levels[l+1],ec)
        if self.norm[l] is not None: e=(e/self.norm[l]).astype(f32)
        x=(x+self.scale*e).astype(f32)
        x=smooth(L,x,b,h2,self.post)
        return x

def pcg(L,b,inv_h2,prec,tol=1e-6,maxit=200000):
    x=np.zeros_like(b); r=b.copy()
    rhs2=float((b.astype(np.float64)**2).sum()); thr=tol*tol*rhs2
    z=prec(r); p=z.copy(); absNew=float((r.astype(np.float64)*z).sum())
    it=0
    while it<maxit:
        q=applyA(L,p,inv_h2)
        alpha=f32(absNew/float((p.astype(np.float64)*q).sum()))
        x=(x+alpha*p).astype(f32); r=(r-alpha*q).astype(f32)
        r2=float((r.astype(np.float64)**2).sum())
        if r2<thr: break
        z=prec(r); absOld=absNew; absNew=float((r.astype(np.float64)*z).sum())
        beta=f32(absNew/absOld); p=(z+beta*p).astype(f32); it+=1
    return x,it+1,np.sqrt(r2/rhs2)

def tank(n):
    from bench import tank_fields
    lab,u,v=tank_fields(n)
    return lab,u,v

def rhs_from(lab,u,v,dx):
    liq=lab==LIQ
    ue=np.concatenate([u[:,1:],u[:,-1:]],1); vn=np.concatenate([v[1:,:],v[-1:,:]],0)
    return np.where(liq,(ue-u)/f32(dx)+(vn-v)/f32(dx),f32(0)).astype(f32)

if __name__=="__main__":
    which=sys.argv[1]; n=int(sys.argv[2])
    if which=="tank":
        lab,u,v=tank(n)
    else:
        rng=np.random.default_rng(3)
        lab=scenes.random_labels(n,n,rng,p_solid=float(sys.argv[3]) if len(sys.argv)>3 else 0.03)
        u=scenes.random_field(n,n,rng); v=scenes.random_field(n,n,rng)
    dx=f32(1)/f32(n)
    b=rhs_from(lab,u,v,dx)
    L=make_level(lab); h2=f32(1)/(dx*dx)
    dinv=np.where(L['cnt']>0,f32(-1)/(np.maximum(L['cnt'],1)*h2),f32(0)).astype(f32)
    t=time.time(); xj,itj,ej=pcg(L,b,h2,lambda r:(dinv*r).astype(f32)); tj=time.time()-t
    print("jacobi iters",itj,"relres",ej,"t",round(tj,1))
    for kw in [dict(),dict(pre=1,post=1),dict(pre=3,post=3),dict(scale=0.5)]:
        mg=MG(lab,dx,**kw)
        t=time.time(); xm,itm,em=pcg(L,b,h2,lambda r:mg.vcycle(r)); tm=time.time()-t
        rel=np.linalg.norm(xm.astype(np.float64)-xj)/np.linalg.norm(xj.astype(np.float64))
        print("mg",kw,"levels",len(mg.levels),"iters",itm,"relres",em,"rel diff vs jacobi",rel,"t",round(tm,1))
