"""tuning probe: device time of the batched track kernel on the bench workload (not a test)"""
import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np
import sdv_loam_b200
from sdv_loam_b200 import synth, api
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
th = int(sys.argv[2]) if len(sys.argv) > 2 else 128
cs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
seq, _ = bench.load_sequence()
w,h = seq.wh
pts=synth.select_points(seq.images[0], seq.clouds[0], 2000)
p4=np.concatenate([pts, np.full((len(pts),1),1e-3,np.float32)],1).astype(np.float32); rh=np.zeros(len(p4),np.int32)
ctx=api.Context(seq.K,w,h,n_tracker_slots=B,max_frames=B+2,cluster_size=cs,track_threads=th)
KF=1<<40
for b in range(B):
    ctx.makeImages(KF, seq.images[0]); api.CoarseTracker(ctx,b).setCoarseTrackingRef(KF,p4,rh); ctx.releaseFrame(KF)
gts, inits = bench.gt_and_inits(seq, synth, B, 6, seed=7)
for b in range(B): ctx.makeImages(b, seq.images[1])
ctx.sync()
ms=[]; ev=0
for s in (0,3,0,3,0,3):       # steps whose init belongs to frame 1 (s%3==0)
    T=inits[s].copy(); ab=np.zeros((B,2))
    r=ctx.trackBatch(list(range(B)), list(range(B)), T, ab)
    ms.append(ctx.last_kernel_ms()); ev=int(r['evals'].sum())
ms=np.array(ms[2:])
print(f"lib={os.environ.get('SDV_B200_LIB','default')} B={B} th={th} cs={cs} kernel ms {ms.mean():.4f} (min {ms.min():.4f}) evals {ev} alg GB/s {ev*64/ms.mean()/1e6:.1f} frac {ev*64/ms.mean()/1e6/6572.9:.3f} good {r['good'].mean():.3f} its {r['iterations'].sum(0)}")
