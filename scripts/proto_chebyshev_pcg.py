import numpy as np
rng=np.random.default_rng(0)
n=48
def lap_sum(x):
    out=np.zeros_like(x)
    out[1:]+=x[:-1]; out[:-1]+=x[1:]
    out[:,1:]+=x[:,:-1]; out[:,:-1]+=x[:,1:]
    out[:,:,1:]+=x[:,:,:-1]; out[:,:,:-1]+=x[:,:,1:]
    return out
def run(c, ks=(0,2,3,4,5,6)):
    # M = (1+6c) I - c * lapsum ; eigenvalues in (1, 1+12c)
    M=lambda v:(1+6*c)*v - c*lap_sum(v)
    lmin,lmax=1.0,1+12*c
    b=rng.standard_normal((n,n,n)); x0=rng.standard_normal((n,n,n))
    def prec(r,k):
        if k==0: return r.copy()
        th=(lmax+lmin)/2; de=(lmax-lmin)/2; sg=th/de
        rho=1/sg; d=r/th; z=d.copy()
        for j in range(1,k):
            rho_n=1/(2*sg-rho)
            d=rho_n*rho*d+(2*rho_n/de)*(r-M(z))
            z=z+d; rho=rho_n
        return z
    res={}
    for k in ks:
        x=x0.copy(); r=b-M(x); bb=np.sqrt((b*b).sum()); z=prec(r,k); p=z.copy(); rz=(r*z).sum(); it=0
        while np.sqrt((r*r).sum())>=1e-12*bb and it<1000:
            q=M(p); a=rz/(p*q).sum(); x+=a*p; r-=a*q; z=prec(r,k); rzn=(r*z).sum(); p=z+(rzn/rz)*p; rz=rzn; it+=1
        true=np.sqrt(((b-M(x))**2).sum())/bb
        res[k]=(it, true)
    return res
for c in (0.144, 0.67, 1.16, 1.64, 6.55):
    r=run(c)
    print(f"c={c} kappa={1+12*c:.1f}: ", {k:(v[0], f"{v[1]:.1e}", (f"streams {v[0]*(8 if k==0 else 8+2.55+ (0))}")) for k,v in r.items()})
