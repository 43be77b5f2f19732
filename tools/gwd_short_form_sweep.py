"""fp64-vs-fp32 error sweep of the two forms of U = tr(Sp St) + 2 sqrt(det det) in the GWD value
(all-positive long form vs U = V^2 - (A-B)(C-D) sin^2, csrc/gd_math.cuh gd::pw::gwd_value): random
boxes with in-plane aspect ratios up to `am`, ordinary and near-coincident pairs.  Prints, per
aspect bound, the 50 / 99 / 99.99 / 99.9999 % quantiles of the relative error of the normalised
distance for the long form and then for the short form (ordinary pairs, then near pairs).
Numpy only; the output is recorded in profiles/r03_pairwise.md."""
import numpy as np
rng=np.random.default_rng(0)
def run(aspect_max, n=2_000_000, near=False):
    f=np.float32
    # boxes: half extents
    base=np.exp(rng.uniform(np.log(0.05),np.log(20),n))
    asp_p=np.exp(rng.uniform(0,np.log(aspect_max),n)); asp_t=np.exp(rng.uniform(0,np.log(aspect_max),n))
    swap_p=rng.random(n)<0.5; swap_t=rng.random(n)<0.5
    ap=base*np.where(swap_p,asp_p,1); bp=base*np.where(swap_p,1,asp_p)
    sc=np.exp(rng.normal(0,0.3,n)) if not near else np.exp(rng.normal(0,1e-3,n))
    at=base*sc*np.where(swap_t,asp_t,1); bt=base*sc*np.where(swap_t,1,asp_t)
    ep=base*np.exp(rng.normal(0,0.3,n)); et=ep*np.exp(rng.normal(0,0.3 if not near else 1e-3,n))
    dl=rng.uniform(-np.pi,np.pi,n)
    dx=rng.normal(0,1,n)*base*(1 if not near else 1e-3); dy=rng.normal(0,1,n)*base*(1 if not near else 1e-3); dz=rng.normal(0,.3,n)*base*(1 if not near else 1e-3)
    def dist(ap,bp,at,bt,ep,et,sd,cd,dx,dy,dz,short,T):
        ap,bp,at,bt,ep,et,sd,cd,dx,dy,dz=[x.astype(T) for x in (ap,bp,at,bt,ep,et,sd,cd,dx,dy,dz)]
        A=ap*ap;B=bp*bp;C=at*at;D=bt*bt
        s2=sd*sd;c2=cd*cd
        K=(ap*bp)*(at*bt)
        V=ap*at+bp*bt
        amb=(ap-bp)*(ap+bp); cmd=(at-bt)*(at+bt)
        eps=amb*cmd*s2
        if short: U=V*V-eps
        else: U=(A*C+B*D)*c2+(A*D+B*C)*s2+T(2)*K
        rU=np.sqrt(np.maximum(U,0))
        eta=eps/(V+rU)
        da=ap-at;db=bp-bt;de=ep-et
        W=da*da+db*db+T(2)*eta+de*de
        d2=dx*dx+dy*dy+dz*dz+W
        d=np.sqrt(np.maximum(d2,0))
        n_=T(2)*((ap*bp*ep)*(at*bt*et))**(T(1)/T(6))
        return d/n_
    sd=np.sin(dl);cd=np.cos(dl)
    ref=dist(ap,bp,at,bt,ep,et,sd,cd,dx,dy,dz,False,np.float64)
    # float32 inputs rounded first, reference on rounded inputs
    args=[x.astype(f) for x in (ap,bp,at,bt,ep,et,sd,cd,dx,dy,dz)]
    ref=dist(*[x.astype(np.float64) for x in args],False,np.float64)
    lo=dist(*args,False,f); sh=dist(*args,True,f)
    den=np.maximum(np.abs(ref),1e-30)
    return [float('%.2g'%np.quantile(np.abs(x-ref)/den,q)) for x in (lo,sh) for q in (0.5,0.99,0.9999,0.999999)]
for am in (3,30,1000):
    print(am, run(am), run(am,near=True))
