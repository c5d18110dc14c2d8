"""tuning probe: clock64 phase breakdown of track_cluster_kernel (needs a library built with -DSDV_TRACK_PROFILE; not a test)"""
import sys, os, ctypes
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np
import sdv_loam_b200
from sdv_loam_b200 import synth, api
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
th = int(sys.argv[2]) if len(sys.argv) > 2 else 128
cs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
seq, _ = bench.load_sequence()
w, h = seq.wh
pts = synth.select_points(seq.images[0], seq.clouds[0], 2000)
p4 = np.concatenate([pts, np.full((len(pts), 1), 1e-3, np.float32)], 1).astype(np.float32); rh = np.zeros(len(p4), np.int32)
ctx = api.Context(seq.K, w, h, n_tracker_slots=B, max_frames=B + 2, cluster_size=cs, track_threads=th)
KF = 1 << 40
for b in range(B):
    ctx.makeImages(KF, seq.images[0]); api.CoarseTracker(ctx, b).setCoarseTrackingRef(KF, p4, rh); ctx.releaseFrame(KF)
gts, inits = bench.gt_and_inits(seq, synth, B, 6, seed=7)
for b in range(B):
    ctx.makeImages(b, seq.images[1])
ctx.sync()
prof = np.zeros(16, np.int64)
f = api.LIB.sdv_debug_track_profile; f.argtypes = [ctypes.c_void_p, ctypes.c_int]
for s in (0, 3, 0, 3):
    T = inits[s].copy(); ab = np.zeros((B, 2))
    f(None, 1)
    r = ctx.trackBatch(list(range(B)), list(range(B)), T, ab)
    f(prof.ctypes.data, 0)
names = ["issue", "flowpass", "first wait+project", "main loop", "drain+sync", "reduce", "control", "-", "(evals)", "ctl: finalize/accept", "ctl: ldlt solve", "ctl: exp/mul", "ctl: eval params+publish"]
n = max(int(prof[8]), 1)
print(f"B={B} th={th} cs={cs} kernel {ctx.last_kernel_ms():.4f} ms, evals (block 0) {n}")
for i, nm in enumerate(names):
    print(f"  {nm:20s} {prof[i]/n:10.0f} cycles/eval  {prof[i]/1.965e3:9.1f} us total")
