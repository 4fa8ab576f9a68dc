"""Evidence for the GMRES iteration-count criterion: the CPU oracle against itself under
rounding-level perturbations of the right-hand side (bowl_wind, first inversion).
Run: python tools/gmres_sensitivity.py   (about 1.5 minutes of CPU)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nupgcm_b200 import workloads as W
from oracle import krylov
from oracle.stepping import cpu_model_for
w = W.bowl_wind()
ops=W.host_operands(w)
m=cpu_model_for(w,ops,solver='krylov')
# step 1: evolve then the inversion RHS
u_prev,b_prev=m.xu.copy(),m.xb.copy()
m.evolve(u_prev,b_prev)
A=ops['A']; N=A.shape[0]
y=ops['B']@m.xb+ops['b0']
M=np.full(N,ops['pscale'])
rng=np.random.default_rng(0)
res={}
for tag,(yy,orth) in {'mgs':(y,'mgs'),'cgs2':(y,'cgs2'),'mgs+1e-16':(y*(1+1e-16*rng.standard_normal(N)),'mgs'),'mgs+1e-14':(y*(1+1e-14*rng.standard_normal(N)),'mgs'),'mgs+1e-12':(y*(1+1e-12*rng.standard_normal(N)),'mgs')}.items():
    t0=time.time()
    x,st=krylov.gmres(A,yy,x0=np.zeros(N),M=M,atol=1e-6,rtol=1e-6,memory=20,orth=orth)
    res[tag]=np.array(st.residuals)
    print(tag,'niter',st.niter,'solved',st.solved,'t',time.time()-t0,flush=True)
base=res['mgs']
for tag,r in res.items():
    n=min(len(r),len(base))
    d=np.abs(r[:n]-base[:n])/base[:n]
    idx=[100,500,1000,2000,3000,4000,5000]
    print(tag,[f"{d[i]:.1e}" for i in idx if i<n])
