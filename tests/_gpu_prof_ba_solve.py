"""tuning probe: clock64 phase breakdown of ba_solve_kernel (needs a library built with -DSDV_BA_PROFILE: SDV_B200_LIB=.../variants/lib_baprof.so; not a test)"""
import sys, ctypes
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np
import sdv_loam_b200
from sdv_loam_b200 import synth, api
from conftest import cached_sequence
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1
seq = cached_sequence(8, 2000, synth.KITTI_K, synth.KITTI_WH); w, h = synth.KITTI_WH
ctx = api.Context(synth.KITTI_K, w, h, max_frames=9)
for k in range(8):
    ctx.makeImages(k, seq.images[k])
wins = [synth.make_ba_window(seq, list(range(7)), n_per_frame=250, seed=3 + i % 4, pose_noise=(0.005, 0.0003), match_noise=0.1, prior_scale=1e-3) for i in range(4)]
f = api.LIB.sdv_debug_ba_profile; f.argtypes = [ctypes.c_void_p, ctypes.c_int]; prof = np.zeros(16, np.int64)
for rep in range(3):
    for i in range(W):
        api.EnergyFunctional(ctx, wins[i % 4], list(range(7)), window=i)
    f(None, 1); r = api.optimize_batch(ctx, list(range(W)), 6); f(prof.ctypes.data, 0)
names = ["stitch", "publish+scale", "ldlt", "ldlt+solve", "ortho", "tail (x, xAd)", "whole kernel", "(calls)", "stitch: bucket products", "stitch: top blocks", "stitch: symmetrise", "stitch: Schur frame-frame"]
n = max(int(prof[7]), 1)
print(f"W={W} optimize device ms {r['ms']:.3f}; solve calls {n}")
for i, nm in enumerate(names):
    print(f"  {nm:28s} {prof[i]/n:10.0f} cycles/call {prof[i]/n/1.965e3:8.1f} us")
