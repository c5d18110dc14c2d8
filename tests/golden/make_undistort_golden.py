"""Writes tests/golden/undistort_small.npz from the REFERENCE's own Undistort (oracle/_ref, built from /root/reference): run in the build container only.
    python tests/golden/make_undistort_golden.py
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref
from test_undistort import CALIBS, SMALL

r = ref.Undistort(SMALL)
rng = np.random.default_rng(2024); yy, xx = np.mgrid[0:r.hOrg, 0:r.wOrg]
raw = ((xx * 2 + yy * 3 + rng.integers(0, 40, (r.hOrg, r.wOrg))) % 256).astype(np.uint8)
out = dict(text=SMALL, K4d=r.K4d, remapX=r.remapX, remapY=r.remapY, raw=raw, image=r.undistort(raw))
for name in ("kitti00", "kitti360"):
    f = ref.Undistort(CALIBS[name]); out[name + "_K4d"] = f.K4d
    out[name + "_sum"] = np.array([f.remapX.astype(np.float64).sum(), f.remapY.astype(np.float64).sum()])
np.savez_compressed(os.path.join(HERE, "undistort_small.npz"), **out)
print("wrote", os.path.join(HERE, "undistort_small.npz"), {k: np.asarray(v).shape for k, v in out.items()})
