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


def test_normal_equations_by_finite_differences():
    """calcHandb (CoarseTracker.cpp:889-947): J = d(unit-plane projection)/d(left se(3) increment of worldToCur), H = sum w J^T J, b = sum w J^T r with
    Tukey weights — re-derived here by numerical differentiation of an independent float64 projection."""
    d = synth.make_overlap_points(120, 4, 6, outlier_frac=0.05, match_noise=0.5)
    K = d["K"]; fx, fy, cx, cy = K; p = d["pts"]
    Hm, bv, _, num = orc.struct_pose_hb(W, H, K.astype(np.float32), d["host_T7"], pts6_of(p), d["T_init"]); assert num == len(p)

    def unit_proj(w2c, i):
        hostT = d["host_T7"][p["host"][i]]; X = orc.se3_rot(hostT) @ (np.array([(p["u"][i] - cx) / fx, (p["v"][i] - cy) / fy, 1.0]) / p["idepth"][i]) + hostT[4:]
        Xc = orc.se3_rot(w2c) @ X + w2c[4:]; return Xc[:2] / Xc[2]

    w2c = orc.se3_inv(d["T_init"]); Href = np.zeros((6, 6)); bref = np.zeros(6); eps = 1e-6
    for i in range(len(p)):
        r = unit_proj(w2c, i) - np.array([(p["obs_x"][i] - cx) / fx, (p["obs_y"][i] - cy) / fy])
        J = np.zeros((2, 6))
        for k in range(6):
            e = np.zeros(6); e[k] = eps
            J[:, k] = (unit_proj(orc.se3_mul(orc.se3_exp(e), w2c), i) - unit_proj(orc.se3_mul(orc.se3_exp(-e), w2c), i)) / (2 * eps)
        # Two entries of the reference's analytic Jacobian are NOT the derivative: d_xi_x[4] = 1 + X*d_xi_x[2] = 1 - (X/Z)^2 and d_xi_y[3] = -(1 + Y*d_xi_y[2]) =
        # -(1 - (Y/Z)^2) (CoarseTracker.cpp:916,922), where the true values are 1 + (X/Z)^2 and -(1 + (Y/Z)^2).  The restatement keeps the reference's
        # expressions (they steer structPoseEstimation's steps); every other entry must equal the numerical derivative.
        u_, v_ = unit_proj(w2c, i)
        assert abs(J[0, 4] - (1 + u_ * u_)) < 1e-6 and abs(J[1, 3] + (1 + v_ * v_)) < 1e-6          # the numerical derivative has the textbook form
        J[0, 4] = 1 - u_ * u_; J[1, 3] = -(1 - v_ * v_)
        x = np.linalg.norm(r); wgt = (1 - x * x / 4.6851 ** 2) ** 2 if x <= 4.6851 else 0.0
        Href += wgt * J.T @ J; bref += wgt * J.T @ r
    assert np.allclose(Hm, Href, rtol=2e-4, atol=1e-6 * np.abs(Href).max()) and np.allclose(bv, bref, rtol=2e-3, atol=2e-5 * (np.abs(bref).max() + 1e-9))
