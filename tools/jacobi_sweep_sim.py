"""tools/jacobi_sweep_sim.py — CPU simulation (numpy) of the cyclic one-sided Jacobi ordering on a benchmark-like R-hat: sweeps needed
with and without cheap preconditioners (transpose, column sorting, one or two QR steps of the transpose, column-pivoted QR).
Result (round 2): 8-9 sweeps in every variant - the dense logspace spectrum, not the starting point, sets the sweep count."""
import numpy as np, time
def rr_pairs(N):
    # classic round-robin: N even, N-1 steps of N/2 disjoint pairs
    idx=list(range(N)); steps=[]
    for s in range(N-1):
        steps.append([(idx[i], idx[N-1-i]) for i in range(N//2)])
        idx=[idx[0]]+[idx[-1]]+idx[1:-1]
    return steps
def jacobi_sweeps(G, maxsw=40, tol=None):
    G=G.copy(); n=G.shape[1]; N=n+(n&1)
    if N!=n: G=np.hstack([G,np.zeros((G.shape[0],1))])
    tol=tol or 2.2e-16*np.sqrt(n)
    steps=rr_pairs(N)
    hist=[]
    for sw in range(maxsw):
        cmax=0.0; rot=0
        for st in steps:
            p=np.array([a for a,b in st]); q=np.array([b for a,b in st])
            X=G[:,p]; Y=G[:,q]
            a=(X*X).sum(0); b=(Y*Y).sum(0); c=(X*Y).sum(0)
            ab=a*b
            act=(c*c>tol*tol*ab)&(a!=0)&(b!=0)
            if not act.any(): continue
            with np.errstate(all='ignore'):
                cos2=np.where(ab>0,c*c/ab,0.0)
            cmax=max(cmax,cos2[act].max()); rot+=act.sum()
            zeta=(b-a)/(2*c+(~act))
            t=np.sign(zeta)/(np.abs(zeta)+np.sqrt(1+zeta*zeta)); t=np.where(zeta==0,1.0,t)
            cs=1/np.sqrt(1+t*t); sn=cs*t
            cs=np.where(act,cs,1.0); sn=np.where(act,sn,0.0)
            G[:,p]=cs*X-sn*Y; G[:,q]=sn*X+cs*Y
        hist.append(np.sqrt(cmax))
        if rot==0 or cmax<=1e-16: break
    return sw+1, hist
rng=np.random.default_rng(0)
# bench-like R-hat
m,n,r,l=6000,2500,640,520
X=rng.standard_normal((m,r))/np.sqrt(m); W=rng.standard_normal((n,r))/np.sqrt(n)
A=(X*np.logspace(1,-3,r))@W.T+1e-6*rng.standard_normal((m,n))
Om=rng.standard_normal((n,l)).astype(np.float32).astype(np.float64)
Y=A@Om; Q,_=np.linalg.qr(Y); Z=A.T@Q; Z,_=np.linalg.qr(Z); Y=A@Z; Q,_=np.linalg.qr(Y); Bt=A.T@Q
Rh=np.linalg.qr(Bt)[1]
for name,G in (("R-hat",Rh),("R-hat^T",Rh.T.copy()),("R-hat cols sorted by norm desc",Rh[:,np.argsort(-np.linalg.norm(Rh,axis=0))]),("R-hat^T cols sorted",Rh.T[:,np.argsort(-np.linalg.norm(Rh.T,axis=0))]),("R-hat cols sorted asc",Rh[:,np.argsort(np.linalg.norm(Rh,axis=0))])):
    t0=time.time(); sw,h=jacobi_sweeps(G); print("%-34s sweeps %2d  max|cos| per sweep: %s  (%.0f s)"%(name,sw," ".join("%.0e"%x for x in h),time.time()-t0),flush=True)
print("--- cheap preconditioners")
import scipy.linalg as sla
R2=np.linalg.qr(Rh.T)[1]            # Rh^T = Q2 R2
R3=np.linalg.qr(R2.T)[1]
Qp,Rp,piv=sla.qr(Rh,pivoting=True)
for name,G in (("R2 = qr(Rh^T).R",R2),("R2^T",R2.T.copy()),("R3 = qr(R2^T).R",R3),("R3^T",R3.T.copy()),("pivoted QR of Rh: R1^T",Rp.T.copy())):
    t0=time.time(); sw,h=jacobi_sweeps(G); print("%-34s sweeps %2d  max|cos| per sweep: %s  (%.0f s)"%(name,sw," ".join("%.0e"%x for x in h),time.time()-t0),flush=True)
