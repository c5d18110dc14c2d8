"""Probe (not a test): S-STRESS shape (BASELINE.json configs[4]: 1920x1200, 5 pyramid levels) — the step-wise residual/Jacobian kernel
(coarse_res_gs_kernel = calcRes + calcGSSSE fused, "solve on host" mode) on one sequence, algorithmic GB/s = 64 B x points / device time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import sdv_loam_b200  # noqa
from sdv_loam_b200 import api
w, h = 1920, 1200; K = (1000.0, 1000.0, 959.5, 599.5)
rng = np.random.default_rng(0)
yy, xx = np.mgrid[0:h, 0:w]
img = (127 + 60 * np.sin(xx / 23.0) * np.cos(yy / 31.0) + 30 * np.sin((xx + yy) / 7.0)).astype(np.float32)
ctx = api.Context(K, w, h, max_frames=3); L = ctx.levels
ctx.makeImages(0, img); ctx.makeImages(1, np.roll(img, 2, axis=1))
tr = api.CoarseTracker(ctx, 0)
T = np.array([1, 0, 0, 0, 0.01, 0, 0.0])
for lvl in (0, 1):
    wl, hl = w >> lvl, h >> lvl
    for n in (32768, 163840, 1048576):
        u = rng.uniform(4, wl - 5, n).astype(np.float32); v = rng.uniform(4, hl - 5, n).astype(np.float32)
        order = np.lexsort((u, v.astype(np.int32))); u, v = u[order], v[order]              # raster order like makeCoarseDepthL0
        idp = rng.uniform(0.02, 0.2, n).astype(np.float32); col = rng.uniform(0, 255, n).astype(np.float32)
        tr.setCloud(0, lvl, u, v, idp, col)
        ms = []
        for rep in range(12):
            tr.calcRes(1, lvl, T, 0.0, 0.0, 20.0); ms.append(ctx.last_kernel_ms())
        m = float(np.median(ms[2:]))
        print(f"level {lvl} ({wl}x{hl}) points {n:8d}: {m*1e3:8.1f} us per calcRes+calcGSSSE pass -> {64.0*n/(m*1e-3)/1e9:8.1f} GB/s algorithmic")
print("pyramid (makeImages) 1920x1200:"); import time
for rep in range(3):
    t0 = time.perf_counter(); ctx.makeImages(1, img); ctx.sync(); print("  makeImages wall ms %.3f" % (1e3 * (time.perf_counter() - t0)))
ctx.close()
