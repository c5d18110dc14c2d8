"""GPU parity (-m gpu) of sdv_tracker_struct_pose[_batch] against the oracle and the golden fixture.
Bound: identical iteration/accept counts, pose within 1e-9 (the sums are order-identical; only device libm sin/cos ulps differ),
against the north_star tolerance of 1e-3 m / 1e-3 rad."""
import os
import numpy as np
import pytest
import orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refine_small.npz")


def _api():
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import api, synth
    return api, synth


def pts6_of(p):
    return np.stack([p["u"], p["v"], p["idepth"], p["host"].astype(np.float32), p["obs_x"], p["obs_y"]], 1).astype(np.float32)


def to_struct(api, p6):
    s = np.zeros(len(p6), api.OVERLAP_PT_DTYPE)
    s["u"], s["v"], s["idepth"], s["host"], s["obs_x"], s["obs_y"] = p6[:, 0], p6[:, 1], p6[:, 2], p6[:, 3].astype(np.int32), p6[:, 4], p6[:, 5]
    return s


def test_struct_pose_matches_oracle():
    api, synth = _api(); w, h = synth.KITTI_WH
    ctx = api.Context(synth.KITTI_K, w, h, max_frames=2); tr = api.CoarseTracker(ctx, 0)
    for seed, n, nH, of in [(0, 300, 5, 0.05), (1, 1500, 8, 0.1), (2, 17, 1, 0.0), (3, 700, 16, 0.3), (4, 256, 3, 0.0), (5, 257, 3, 0.02)]:
        d = synth.make_overlap_points(n, nH, seed, outlier_frac=of)
        o = orc.struct_pose(w, h, np.array(synth.KITTI_K, np.float32), d["host_T7"], pts6_of(d["pts"]), d["T_init"])
        g = tr.structPoseEstimation(d["T_init"], d["pts"], d["host_T7"])
        assert (g["iterations"], g["accepts"]) == (o["iterations"], o["accepts"]), (seed, g, o)
        assert np.abs(g["T"] - o["T"]).max() < 1e-9 and abs(g["res"] - o["res"]) <= 1e-6 * o["res"]
    ctx.close()


def test_struct_pose_batch_ragged_and_empty():
    api, synth = _api(); w, h = synth.KITTI_WH
    ctx = api.Context(synth.KITTI_K, w, h, max_frames=2)
    ds = [synth.make_overlap_points(n, nH, 20 + i) for i, (n, nH) in enumerate([(120, 3), (0, 2), (640, 6), (1, 1), (333, 5)] * 6)]
    for d in ds:
        if len(d["pts"]) == 0:
            d["pts"] = np.zeros(0, api.OVERLAP_PT_DTYPE)
    T = np.stack([d["T_init"] for d in ds])
    r = api.structPoseEstimationBatch(ctx, T, [d["pts"] for d in ds], [d["host_T7"] for d in ds])
    for k, d in enumerate(ds):
        o = orc.struct_pose(w, h, np.array(synth.KITTI_K, np.float32), d["host_T7"], pts6_of(d["pts"]) if len(d["pts"]) else np.zeros((0, 6), np.float32), d["T_init"])
        assert (int(r["iterations"][k]), int(r["accepts"][k])) == (o["iterations"], o["accepts"]), k
        assert np.abs(r["T"][k] - o["T"]).max() < 1e-9
        if len(d["pts"]) == 0:
            assert np.array_equal(r["T"][k], d["T_init"])
    ctx.close()


def test_struct_pose_golden():
    api, _ = _api(); g = np.load(GOLD)
    ctx = api.Context((383.4, 383.4, 312.0, 97.0), 640, 192, max_frames=2); tr = api.CoarseTracker(ctx, 0)
    for k in range(3):
        r = tr.structPoseEstimation(g[f"Tin{k}"], to_struct(api, g[f"pts{k}"]), g[f"host{k}"])
        assert (r["iterations"], r["accepts"]) == (int(g[f"stat{k}"][1]), int(g[f"stat{k}"][2]))
        assert np.abs(r["T"] - g[f"Tout{k}"]).max() < 1e-9
    ctx.close()


def test_struct_pose_rejects_bad_host_index():
    api, synth = _api(); w, h = synth.KITTI_WH
    ctx = api.Context(synth.KITTI_K, w, h, max_frames=2); tr = api.CoarseTracker(ctx, 0)
    d = synth.make_overlap_points(10, 2, 1); d["pts"]["host"][3] = 7
    with pytest.raises(RuntimeError):
        tr.structPoseEstimation(d["T_init"], d["pts"], d["host_T7"])
    ctx.close()
