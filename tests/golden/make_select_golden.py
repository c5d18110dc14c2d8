"""Generates tests/golden/select_small.npz — golden vectors of the keyframe-rate candidate management and the LiDAR front-end produced BY THE REFERENCE ITSELF
(oracle/_ref: the reference's own PixelSelector2.cpp / FullSystem.cpp / CoarseTracker.cpp / main.cpp compiled unmodified), so that a box without /root/reference (the GPU box)
can hold both the oracle and the CUDA path to reference outputs.  Inputs: the images of tracker_small.npz (640x192), LiDAR pixels derived from its point list, a decimated
synthetic sweep.  Run from the repo root where /root/reference exists:  python tests/golden/make_select_golden.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref, orc                                   # noqa: E402
import sdv_loam_b200                              # noqa: E402,F401
from sdv_loam_b200 import synth                   # noqa: E402
from test_ref_pin_ba import _window               # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "tracker_small.npz")); w, h = 640, 192; K = tuple(float(k) for k in g["K"])
win, ob, rb, _keep = _window((0, 1), 5)          # any FullSystem of this image size (makeNewTraces / shiTomasiScore / the node's handler are members / read the global system)
L = ref.set_calib(w, h, K); img0, img1 = g["img0"].astype(np.float32), g["img1"].astype(np.float32); rf = [ref.Frame(img0, (w, h), L), ref.Frame(img1, (w, h), L)]
p4 = g["pts4"]; cloud = np.stack([p4[:, 0] + 0.3, p4[:, 1] + 0.2, 1.0 / p4[:, 2]], 1).astype(np.float64)
out = dict(K=np.array(K), cloud=cloud)
rs = ref.Selector((w, h), owner=rb); out["random_pattern_crc"] = np.array([int(np.bitwise_xor.reduce(rs.randomPattern().astype(np.uint32) * (np.arange(w * h, dtype=np.uint32) | 1)))], np.int64)
a, b = rs.makeHists(rf[0]); out["ths"], out["thsSmoothed"] = a, b
for name, c in (("lidar", cloud), ("dense", None)):
    for pot, dens in ((3, 400.0), (1, 5000.0), (6, 50.0)):
        rs.currentPotential = pot; m, n = rs.makeMaps(rf[0], dens, 1, 1.0, c)
        out[f"maps_{name}_{pot}"] = m.astype(np.uint8); out[f"num_{name}_{pot}"] = np.array([n, rs.currentPotential], np.int32)
sel = np.zeros((h, w), np.float32); lrud = [int(cloud[:, 0].min()), int(cloud[:, 0].max()), int(cloud[:, 1].min()), int(cloud[:, 1].max())]
for k, (pot, add) in enumerate(((3, 1), (2, 0))):           # second call: addFeaturePoint off -> the monocular map of the first keyframe is walked again
    rs.currentPotential = pot; R = ref.make_new_traces(rb, rf[k], cloud, lrud, add, 600.0, sel)
    out[f"traces{k}"] = R; out[f"traces{k}_pot"] = np.array([pot, rs.currentPotential], np.int32)
out["selection_map_final"] = sel.astype(np.uint8); out["lrud"] = np.array(lrud, np.int32)
world = synth.World(3000); Rt = synth.trajectory(1, 3000); sweep = synth.lidar_sweep(world, Rt[0][0], Rt[1][0], seed=0)[::16].copy()
r = ref.lidar_handler(rb, sweep, synth.RLC, synth.TLC, K, [10000, -1, 10000, -1], images=True)
out.update(sweep=sweep, lidar_cloud_px=r["cloud_px"], lidar_lrud=r["lrud"], lidar_flags=np.array([r["addFeaturePoint"], r["n_segmented"]], np.int32),
           lidar_ground_count=np.array([(r["ground"] == 1).sum(), (r["label"] == 999999).sum()], np.int64))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "select_small.npz"), **out)
print({k: (v.shape, str(v.dtype)) for k, v in out.items()})
