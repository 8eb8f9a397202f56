"""Which of LSMR's tests ends the reference's inner solves?  Runs the unmodified reference (build container only:
needs /root/reference) on a scaled config with a spy around scipy's lsmr and prints istop / iterations / normA / condA.
    python tools/lsmr_spy.py [scale] [config]      # default 0.05 C4: 1 778 cameras, 250 k observations
Round-1 finding (C4 x 0.05): istop = 2 (the atol test) every time, after 119 / 1072 / 870 / 616 / 488 / 400 iterations,
with normA ~ 1.23 sqrt(itn) and normr ~ ||f||: the reference stops when ||A^T res|| <= 1.2e-6 sqrt(itn) ||f||.
"""
import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/reference')
import numpy as np
import bundleAdjuster as ref
from scipy.optimize import least_squares
import scipy.optimize._lsq.trf as T
from meatmodeler_b200 import synth
scale=float(sys.argv[1]) if len(sys.argv)>1 else 0.05
prob=synth.make_config(sys.argv[2] if len(sys.argv)>2 else "C4",hard=True,scale=scale)
ext,K,pts,uv,fi,pi=prob.args()
nc,npts=len(ext),len(pts)
x0=np.hstack((ref.frameParameters(ext),pts.reshape(npts*3)))
A=ref.pointAdjustmentSparsity(nc,npts,fi,pi)
orig=T.lsmr; log=[]
def spy(*a,**k):
    out=orig(*a,**k)
    # x, istop, itn, normr, normar, normA, condA, normx
    log.append(dict(istop=int(out[1]),itn=int(out[2]),normr=float(out[3]),normar=float(out[4]),normA=float(out[5]),condA=float(out[6]),damp=k.get('damp')))
    return out
T.lsmr=spy
costs=[]
t=time.time()
res=least_squares(ref.pointFun,x0,jac_sparsity=A,verbose=0,x_scale="jac",ftol=1e-4,method="trf",args=(K,nc,npts,fi,pi,uv),callback=lambda intermediate_result: costs.append(float(intermediate_result.cost)))
print("sizes",prob.sizes,"nfev",res.nfev,"status",res.status,"cost",res.cost,"time",time.time()-t)
for l,c in zip(log,costs): print(l, "cost", c, "||f||", np.sqrt(2*c), "normar/(normA*normr)", l['normar']/(l['normA']*l['normr']))
