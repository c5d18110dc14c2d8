import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import sdv_loam_b200
from sdv_loam_b200 import synth, api
from conftest import cached_sequence
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
seq = cached_sequence(3, 1000, synth.KITTI_K, synth.KITTI_WH)
w,h = synth.KITTI_WH
pts=synth.select_points(seq.images[0], seq.clouds[0], 2000)
p4=np.concatenate([pts, np.full((len(pts),1),1e-3,np.float32)],1); rh=np.zeros(len(pts),np.int32)
ctx=api.Context(synth.KITTI_K,w,h,n_tracker_slots=B,max_frames=2*B+2,cluster_size=cs)
for i in range(B):
    ctx.makeImages(2*i,seq.images[0]); ctx.makeImages(2*i+1,seq.images[1])
    api.CoarseTracker(ctx,i).setCoarseTrackingRef(2*i,p4,rh)
T0=np.array([1,0,0,0,0,0,0.0])
for rep in range(3):
    T=np.tile(T0,(B,1)); ab=np.zeros((B,2))
    r=ctx.trackBatch(list(range(B)),[2*i+1 for i in range(B)],T,ab)
print("kernel ms", ctx.last_kernel_ms())
