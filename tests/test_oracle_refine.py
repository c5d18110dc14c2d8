"""CPU tests of the structPoseEstimation oracle (oracle/orc_refine.cpp): behaviour + golden regression pin."""
import os
import numpy as np
import orc
import sdv_loam_b200  # noqa
from sdv_loam_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refine_small.npz")
W, H = synth.KITTI_WH


def pts6_of(p):
    return np.stack([p["u"], p["v"], p["idepth"], p["host"].astype(np.float32), p["obs_x"], p["obs_y"]], 1).astype(np.float32)


def run(d, T=None):
    return orc.struct_pose(d["wh"][0], d["wh"][1], d["K"][:4].astype(np.float32), d["host_T7"], pts6_of(d["pts"]), d["T_init"] if T is None else T)


def test_energy_never_increases_and_pose_improves():
    better = 0
    for seed in range(8):
        d = synth.make_overlap_points(400, 5, seed, outlier_frac=0.0, match_noise=0.2)
        r0 = run(d, d["T_init"]); assert 1 <= r0["iterations"] <= 10 and r0["accepts"] <= r0["iterations"]
        e0 = np.linalg.norm(d["T_init"][4:] - d["T_gt"][4:]); e1 = np.linalg.norm(r0["T"][4:] - d["T_gt"][4:])
        better += e1 < e0
        again = run(d, r0["T"]); assert again["res"] <= r0["res"] * (1 + 1e-6)      # restarting from the result cannot be worse
    assert better >= 6


def test_pose_untouched_without_accepted_step():
    d = synth.make_overlap_points(200, 4, 3, outlier_frac=0.0, match_noise=0.0, pose_noise=(0.0, 0.0))
    r = run(d, d["T_gt"])                                 # already at the optimum (float noise only) -> no accepted step moves it far
    assert np.linalg.norm(r["T"][4:] - d["T_gt"][4:]) < 1e-3
    r = orc.struct_pose(W, H, d["K"][:4].astype(np.float32), d["host_T7"], np.zeros((0, 6), np.float32), d["T_init"])
    assert np.array_equal(r["T"], d["T_init"]) and r["accepts"] == 0 and r["iterations"] == 1     # num == 0: resNew = 1e6, inc = 0 -> break


def test_out_of_image_points_are_skipped():
    d = synth.make_overlap_points(200, 4, 9)
    p = d["pts"].copy(); far = p[:50].copy(); far["idepth"] = -0.05      # behind the camera -> projection falls outside -> skipped like :862
    d2 = dict(d); d2["pts"] = np.concatenate([p, far])
    a, b = run(d), run(d2)
    assert a["iterations"] >= 1 and b["iterations"] >= 1 and np.isfinite(b["res"])


def test_golden_regression():
    g = np.load(GOLD)
    for k in range(3):
        r = orc.struct_pose(640, 192, np.array([383.4, 383.4, 312.0, 97.0], np.float32), g[f"host{k}"], g[f"pts{k}"], g[f"Tin{k}"])
        assert np.allclose(r["T"], g[f"Tout{k}"], rtol=0, atol=1e-12)
        assert r["iterations"] == int(g[f"stat{k}"][1]) and r["accepts"] == int(g[f"stat{k}"][2]) and abs(r["res"] - g[f"stat{k}"][0]) <= 1e-6 * g[f"stat{k}"][0]
