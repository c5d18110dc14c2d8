"""Probe (not a test): where the mono8 end-to-end step time goes — upload-only period vs upload + tracking."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import sdv_loam_b200  # noqa
from sdv_loam_b200 import api, synth
from conftest import cached_sequence
B = 592; seq = cached_sequence(4, 1000, synth.KITTI_K, synth.KITTI_WH); w, h = synth.KITTI_WH
ctx = api.Context(synth.KITTI_K, w, h, n_tracker_slots=B, max_frames=2 * B + 2)
pts = synth.select_points(seq.images[0], seq.clouds[0], 2000); p4 = np.concatenate([pts, np.full((len(pts), 1), 1e-3, np.float32)], 1).astype(np.float32); rh = np.zeros(len(p4), np.int32)
for b in range(B):
    ctx.makeImages(1 << 40, seq.images[0]); api.CoarseTracker(ctx, b).setCoarseTrackingRef(1 << 40, p4, rh); ctx.releaseFrame(1 << 40)
host_u8 = torch.empty((2, B, h, w), dtype=torch.uint8).pin_memory()
for k in range(2):
    host_u8[k].copy_(torch.from_numpy(seq.images[1 + k].astype(np.uint8)).expand(B, h, w))
ptrs = [np.uint64(host_u8[k].data_ptr()) + np.arange(B, dtype=np.uint64) * np.uint64(h * w) for k in range(2)]
host_f = torch.empty((2, B, h, w), dtype=torch.float32).pin_memory()
for k in range(2):
    host_f[k].copy_(torch.from_numpy(seq.images[1 + k].astype(np.float32)).expand(B, h, w))
fptrs = [np.uint64(host_f[k].data_ptr()) + np.arange(B, dtype=np.uint64) * np.uint64(h * w * 4) for k in range(2)]
ids = [np.arange(B, dtype=np.uint64) * 2 + p for p in (0, 1)]; slots = np.arange(B, dtype=np.int32)
ID7 = np.tile(np.array([1, 0, 0, 0, 0, 0, -0.9]), (B, 1))
def loop(n, track, u8=True):
    P = ptrs if u8 else fptrs
    ctx.makeImagesBatch(ids[0], P[0], u8=u8); ctx.sync(); t0 = time.perf_counter(); marks = []
    for s in range(n):
        ctx.makeImagesBatch(ids[(s + 1) & 1], P[(s + 1) & 1], u8=u8)
        if track:
            T = ID7.copy(); ab = np.zeros((B, 2)); ctx.trackBatch(slots, ids[s & 1], T, ab)
        marks.append(time.perf_counter() - t0)
    ctx.sync(); return (time.perf_counter() - t0) / n, np.diff([0] + marks)
for u8 in (True, False):
    for track in (False, True):
        loop(3, track, u8); p, d = loop(10, track, u8)
        print("u8" if u8 else "f32", "track" if track else "upload-only", "period ms %.2f" % (1e3 * p), "host marks ms", np.round(1e3 * d, 2).tolist())
t0 = time.perf_counter(); ctx.makeImagesBatch(ids[0], ptrs[0], u8=True); t1 = time.perf_counter(); ctx.sync(); t2 = time.perf_counter()
print("single u8 ingest: enqueue ms %.2f, complete ms %.2f" % (1e3 * (t1 - t0), 1e3 * (t2 - t0)))
ctx.close()
