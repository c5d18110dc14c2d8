"""Profiling driver (not a test): the refinement leg of bench.py alone — 592 frames through sdv_tracker_refine_batch — so that an ncu capture does
not have to skip the tracker launches.  Usage: python tests/_gpu_prof_refine.py [B] [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import sdv_loam_b200  # noqa
from sdv_loam_b200 import api, synth
from conftest import cached_sequence
import orc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
seq8 = cached_sequence(8, 2000, synth.KITTI_K, synth.KITTI_WH); w, h = synth.KITTI_WH
pts, hT, hab = synth.make_map(seq8, list(range(7)), n_per_frame=300, seed=2)
ctx = api.Context(synth.KITTI_K, w, h, n_tracker_slots=B, max_frames=B + 16)
kf_ids = [1000 + k for k in range(7)]
for k in range(7):
    ctx.makeImages(kf_ids[k], seq8.images[k])
rp = api.Reprojector(ctx); slots = np.arange(B, dtype=np.int32); ids = np.arange(B, dtype=np.uint64) + np.uint64(5000)
for b in range(B):
    rp.setMap(b, kf_ids, hT, hab, pts); ctx.makeImages(int(ids[b]), seq8.images[7])
gt = np.concatenate([synth._quat_from_R(seq8.R[7]), seq8.t[7]]); rng = np.random.default_rng(11)
order = rng.permutation(rp.n_cells).astype(np.int32)
for rep in range(reps):
    T0 = np.stack([orc.se3_mul(orc.se3_exp(np.concatenate([rng.normal(0, 0.02, 3), rng.normal(0, 0.001, 3)])), gt) for _ in range(B)])
    t0 = time.perf_counter(); r = rp.refineBatch(slots, ids, T0, cell_order=order, max_matches=400); t1 = time.perf_counter()
    print("rep", rep, "device ms", r["ms"], "wall ms", 1e3 * (t1 - t0), "matches", r["n_matches"].mean(), "its", r["iterations"].mean())
ctx.close()
