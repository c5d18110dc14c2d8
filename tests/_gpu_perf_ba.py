"""tuning probe: device time of the batched BA schedule (not a test)"""
import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import sdv_loam_b200
from sdv_loam_b200 import synth, api
import orc
from conftest import cached_sequence
W = int(sys.argv[1]) if len(sys.argv) > 1 else 118
seq = cached_sequence(8, 2000, synth.KITTI_K, synth.KITTI_WH)
w,h = synth.KITTI_WH
ctx=api.Context(synth.KITTI_K,w,h,max_frames=9)
for k in range(8): ctx.makeImages(k, seq.images[k])
wins=[synth.make_ba_window(seq, list(range(7)), n_per_frame=250, seed=3+i%4, pose_noise=(0.005,0.0003), match_noise=0.1, prior_scale=1e-3) for i in range(4)]
t0=time.time()
gbs=[api.EnergyFunctional(ctx, wins[i%4], list(range(7)), window=i) for i in range(W)]
print("setup s", time.time()-t0, "nP", gbs[0].nP, "nR", gbs[0].nR)
for rep in range(3):
    for i in range(W):            # re-upload windows (optimize mutates them)
        gbs[i]=api.EnergyFunctional(ctx, wins[i%4], list(range(7)), window=i)
    ctx.sync(); t0=time.time()
    r=api.optimize_batch(ctx, list(range(W)), 6); t1=time.time()
    print(f"W={W} optimize_batch device ms {r['ms']:.3f} wall ms {(t1-t0)*1e3:.3f} per window us {r['ms']*1e3/W:.1f} its {np.bincount(r['iterations'])} accepts {np.bincount(r['accepts'])}")
frames=[orc.Frame(seq.images[k],4) for k in range(7)]
ob=orc.BAWindow(wins[0], frames); t0=time.time(); ro=ob.optimize(6); print("oracle ms", (time.time()-t0)*1e3, ro)
