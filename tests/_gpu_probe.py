import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import sdv_loam_b200
from sdv_loam_b200 import synth, api
import orc
from conftest import cached_sequence
seq = cached_sequence(3, 1000, synth.KITTI_K, synth.KITTI_WH)
w,h = synth.KITTI_WH; L = api.pyr_levels(w,h)
f0=orc.Frame(seq.images[0],L); f1=orc.Frame(seq.images[1],L)
pts=synth.select_points(seq.images[0], seq.clouds[0], 2000)
p4=np.concatenate([pts, np.full((len(pts),1),1e-3,np.float32)],1); rh=np.zeros(len(pts),np.int32)
otr=orc.CoarseTracker(w,h,L,synth.KITTI_K); otr.setCoarseTrackingRef(f0,p4,rh)
ctx=api.Context(synth.KITTI_K,w,h)
ctx.makeImages(0,seq.images[0]); ctx.makeImages(1,seq.images[1])
for l in range(L):
    dI,ab=ctx.frameLevel(1,l)
    print("pyr lvl",l,"max|diff| dI", np.abs(dI-f1.dI(l)).max(), "abs", np.abs(ab-f1.absSquaredGrad(l)).max())
tr=api.CoarseTracker(ctx,0); tr.setCoarseTrackingRef(0,p4,rh)
for l in range(L):
    a=tr.cloud(l); b=otr.cloud(l)
    print("cloud",l,len(a[0]),len(b[0]), [float(np.abs(x-y).max()) if len(x)==len(y) else None for x,y in zip(a,b)])
T0=np.array([1,0,0,0,0,0,0.0])
for l in range(L):
    rs_o=otr.calcRes(f1,l,T0,0,0,20.0); H_o,b_o=otr.calcGSSSE(l,T0,0,0)
    rs_g=tr.calcRes(1,l,T0,0,0,20.0); H_g,b_g=tr.calcGSSSE(l)
    print("res",l,rs_o,rs_g, "H rel", np.linalg.norm(H_o-H_g)/np.linalg.norm(H_o), "b rel", np.linalg.norm(b_o-b_g)/np.linalg.norm(b_o), "ms", ctx.last_kernel_ms())
ro=otr.trackNewestCoarse(f1,T0,[0,0],L-1)
t0=time.time(); rg=tr.trackNewestCoarse(1,T0,[0,0]); t1=time.time()
print("oracle", ro['good'], orc.se3_log(ro['T']), ro['ab'], ro['lastResiduals'], ro['iterations'], ro['accepts'])
print("gpu   ", rg['good'], orc.se3_log(rg['T']), rg['ab'], rg['lastResiduals'], rg['iterations'], rg['accepts'], "kernel ms", ctx.last_kernel_ms(), "wall", t1-t0)
print("pose diff", orc.se3_log(orc.se3_mul(rg['T'], orc.se3_inv(ro['T']))))
# batch throughput
for th in (64,128,256):
 for cs in (1,2,8):
  if th==64 and cs>1: continue
  for B in (1,148,592,1184):
    ctx2=api.Context(synth.KITTI_K,w,h,n_tracker_slots=B,max_frames=2*B+2,cluster_size=cs,track_threads=th)
    for i in range(B):
        ctx2.makeImages(2*i,seq.images[0]); ctx2.makeImages(2*i+1,seq.images[1])
        api.CoarseTracker(ctx2,i).setCoarseTrackingRef(2*i,p4,rh)
    for rep in range(3):
        T=np.tile(T0,(B,1)); ab=np.zeros((B,2))
        r=ctx2.trackBatch(list(range(B)),[2*i+1 for i in range(B)],T,ab)
    d=np.abs(T-ro['T']).max()
    ms=ctx2.last_kernel_ms()
    print("threads",th,"cluster",cs,"batch",B,"kernel ms",round(ms,4),"fps",round(B/ms*1e3), "alg GB/s", round(r['evals'].sum()*64/ms/1e6,1), "maxdiff vs oracle", d, flush=True)
    ctx2.close()
