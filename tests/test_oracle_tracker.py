"""Oracle tracker path vs independent numpy formulations, the committed golden pin, and ground truth (CPU only)."""
import os
import numpy as np
import pytest
import orc
from conftest import SMALL_K, SMALL_WH

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tracker_small.npz")


def np_make_images(img, levels):
    """Independent numpy statement of FrameHessian::makeImages (HessianBlocks.cpp:107-167) incl. flat-index wrap."""
    out = []
    I = img.astype(np.float32)
    for l in range(levels):
        if l > 0:
            P = out[-1][0]
            I = np.float32(0.25) * (((P[0::2, 0::2] + P[0::2, 1::2]) + P[1::2, 0::2]) + P[1::2, 1::2])
        h, w = I.shape; f = I.reshape(-1)
        dx = np.zeros(w * h, np.float32); dy = np.zeros(w * h, np.float32)
        idx = np.arange(w, w * (h - 1))
        dx[idx] = np.float32(0.5) * (f[idx + 1] - f[idx - 1]); dy[idx] = np.float32(0.5) * (f[idx + w] - f[idx - w])
        out.append((I, dx.reshape(h, w), dy.reshape(h, w), (dx * dx + dy * dy).reshape(h, w)))
    return out


def test_make_images_matches_numpy(small_seq):
    img = small_seq.images[0]; L = 4
    f = orc.Frame(img, L); ref = np_make_images(img, L)
    for l in range(L):
        dI = f.dI(l)
        assert np.array_equal(dI[..., 0], ref[l][0]) and np.array_equal(dI[..., 1], ref[l][1]) and np.array_equal(dI[..., 2], ref[l][2])
        assert np.array_equal(f.absSquaredGrad(l), ref[l][3])


def test_makeK_levels():
    tr = orc.CoarseTracker(1200, 360, 4, (718.856, 718.856, 607.1928, 185.2157))
    for l in range(4):
        fx, fy, cx, cy = tr.K(l)
        assert np.isclose(fx, 718.856 / 2 ** l, rtol=1e-7) and np.isclose(cx, (607.1928 + 0.5) / 2 ** l - 0.5, rtol=1e-6)   # CoarseTracker.cpp:87-95
        Ki = tr.Ki(l); Km = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
        assert np.allclose(Ki, np.linalg.inv(Km), rtol=1e-6)


def test_coarse_depth_properties(small_seq):
    w, h = SMALL_WH; L = 4
    f0 = orc.Frame(small_seq.images[0], L)
    tr = orc.CoarseTracker(w, h, L, SMALL_K)
    # two splats on one pixel (weighted mean), one isolated splat, with both rounding rules
    pts = np.array([[100.7, 50.2, 0.10, 1e-3], [100.1, 50.9, 0.30, 1e-3 / 4], [300.6, 120.6, 0.05, 1e-3]], np.float32)
    tr.setCoarseTrackingRef(f0, pts, np.array([0, 0, 1], np.int32))
    u, v, idp, col = tr.cloud(0)
    got = {(int(a), int(b)): c for a, b, c in zip(u, v, idp)}
    w1, w2 = 1.0, 2.0                                                    # sqrt(1e-3/HdiF): 1 and 2
    assert np.isclose(got[(100, 50)], (0.10 * w1 + 0.30 * w2) / (w1 + w2), rtol=1e-5)
    assert np.isclose(got[(301, 121)], 0.05, rtol=1e-6)                 # +0.5 rounding (CoarseTracker.cpp:285-286)
    for d in ((1, 1), (-1, -1), (1, -1), (-1, 1)):                       # diagonal 1-px dilation on level 0 (:343-346)
        assert np.isclose(got[(301 + d[0], 121 + d[1])], 0.05, rtol=1e-6)
    assert (302, 121) not in got
    assert len(u) == 2 * 5
    assert np.array_equal(col, f0.dI(0)[v.astype(int), u.astype(int), 0])
    assert all(np.all(np.diff(tr.cloud(l)[1] * 10000 + tr.cloud(l)[0]) > 0) for l in range(L))   # raster order


def test_calcres_identity_pose_is_zero_residual(small_seq):
    """Size-independent property: tracking a frame against itself at identity gives r == 0 for every point."""
    w, h = SMALL_WH; L = 4
    f0 = orc.Frame(small_seq.images[0], L)
    from sdv_loam_b200 import synth
    pts = synth.select_points(small_seq.images[0], small_seq.clouds[0], 500)
    p4 = np.concatenate([pts, np.full((len(pts), 1), 1e-3, np.float32)], 1)
    tr = orc.CoarseTracker(w, h, L, SMALL_K); tr.setCoarseTrackingRef(f0, p4, np.zeros(len(p4), np.int32))
    T = np.array([1, 0, 0, 0, 0, 0, 0.0])
    for l in range(L):
        rs = tr.calcRes(f0, l, T, 0, 0, 20.0)
        assert rs[0] < 1e-3 * rs[1] and rs[5] == 0 and rs[1] > 0.8 * len(tr.cloud(l)[0])
        W = tr.warped(); n = int(rs[1])
        assert np.abs(W[5, :n]).max() < 2e-2 and np.all(W[6, :n] == 1.0)      # residual ~ 0 (bilinear at integer pixel), hw = 1
        H, b = tr.calcGSSSE(l, T, 0, 0)
        assert np.allclose(H, H.T) and np.all(np.linalg.eigvalsh(H) > -1e-6 * np.abs(H).max())


def test_gs_matches_float64_normal_equations(small_seq):
    """calcGSSSE against a float64 numpy accumulation of the same J,w,r (MatrixAccumulators.h:1040-1115, CoarseTracker.cpp:442-483)."""
    w, h = SMALL_WH; L = 4
    f0, f1 = orc.Frame(small_seq.images[0], L), orc.Frame(small_seq.images[1], L)
    from sdv_loam_b200 import synth
    pts = synth.select_points(small_seq.images[0], small_seq.clouds[0], 500)
    p4 = np.concatenate([pts, np.full((len(pts), 1), 1e-3, np.float32)], 1)
    tr = orc.CoarseTracker(w, h, L, SMALL_K); tr.setCoarseTrackingRef(f0, p4, np.zeros(len(p4), np.int32), 0.0, 2.0)
    T = orc.se3_exp([0.0, 0.0, -0.9, 0, 0, 0])
    for l in range(L):
        tr.calcRes(f1, l, T, 0.01, 0.5, 20.0)
        H, b = tr.calcGSSSE(l, T, 0.01, 0.5)
        idp, u, v, dx, dy, r, hw, rc = tr.warped().astype(np.float64)
        fx, fy, _, _ = tr.K(l).astype(np.float64)
        a = np.exp(0.01); dxf, dyf = dx * fx, dy * fy
        J = np.stack([idp * dxf, idp * dyf, -idp * (u * dxf + v * dyf), -(u * v * dxf + dyf * (1 + v * v)), u * v * dyf + dxf * (1 + u * u),
                      u * dyf - v * dxf, a * (2.0 - rc), -np.ones_like(u), r], 0)
        M = (J * hw) @ J.T / len(u)
        sc = np.array([1, 1, 1, .5, .5, .5, 10, 1000.0])
        assert np.allclose(H, M[:8, :8] * sc[:, None] * sc[None, :], rtol=2e-4, atol=1e-6 * np.abs(H).max())
        assert np.allclose(b, M[:8, 8] * sc, rtol=2e-4, atol=1e-6 * np.abs(b).max())


def test_golden_pin():
    """The oracle reproduces its committed fixture (tests/golden/make_golden.py) — guards against silent drift."""
    g = np.load(GOLD)
    w, h = SMALL_WH; L = 4; K = tuple(g["K"])
    f0, f1 = orc.Frame(g["img0"].astype(np.float32), L), orc.Frame(g["img1"].astype(np.float32), L)
    tr = orc.CoarseTracker(w, h, L, K); tr.setCoarseTrackingRef(f0, g["pts4"], g["round_half"], *g["ref_ab"])
    assert np.array_equal([len(tr.cloud(l)[0]) for l in range(L)], g["pc_n"])
    for l in range(L):
        assert np.array_equal(np.stack(tr.cloud(l)), g[f"cloud{l}"])
        assert np.isclose(f1.dI(l).astype(np.float64).sum(), g["pyr_checksum"][l], rtol=1e-12)
        rs = tr.calcRes(f1, l, g["T_eval"], 0.02, 1.0, 20.0); H, b = tr.calcGSSSE(l, g["T_eval"], 0.02, 1.0)
        assert np.allclose(rs, g[f"rs{l}"], rtol=1e-12) and np.allclose(H, g[f"H{l}"], rtol=1e-12) and np.allclose(b, g[f"b{l}"], rtol=1e-12)
    r = tr.trackNewestCoarse(f1, np.array([1, 0, 0, 0, 0, 0, 0.0]), [0.0, 0.0], L - 1)
    assert r["good"] == bool(g["track_good"]) and np.array_equal(r["iterations"], g["track_iterations"]) and np.array_equal(r["accepts"], g["track_accepts"])
    assert np.allclose(r["T"], g["track_T"], atol=1e-12) and np.allclose(r["lastResiduals"], g["track_lastRes"], rtol=1e-10, equal_nan=True)


def test_tracker_converges_to_ground_truth(kitti_seq):
    """End-to-end sanity of the restated LM on KITTI-shape input: identity start -> ground-truth relative pose."""
    from sdv_loam_b200 import synth
    w, h = synth.KITTI_WH; L = 4
    f0, f1 = orc.Frame(kitti_seq.images[0], L), orc.Frame(kitti_seq.images[1], L)
    pts = synth.select_points(kitti_seq.images[0], kitti_seq.clouds[0], 2000)
    p4 = np.concatenate([pts, np.full((len(pts), 1), 1e-3, np.float32)], 1)
    tr = orc.CoarseTracker(w, h, L, synth.KITTI_K); tr.setCoarseTrackingRef(f0, p4, np.zeros(len(p4), np.int32))
    r = tr.trackNewestCoarse(f1, np.array([1, 0, 0, 0, 0, 0, 0.0]), [0.0, 0.0], L - 1)
    Tgt = orc.se3_from_rt(*synth.rel_pose(kitti_seq.R[0], kitti_seq.t[0], kitti_seq.R[1], kitti_seq.t[1]))
    err = orc.se3_log(orc.se3_mul(r["T"], orc.se3_inv(Tgt)))
    assert r["good"] and np.linalg.norm(err[:3]) < 5e-3 and np.linalg.norm(err[3:]) < 5e-4
    # abort rule: an impossible minResForAbort stops after the coarsest level (CoarseTracker.cpp:810)
    r2 = tr.trackNewestCoarse(f1, np.array([1, 0, 0, 0, 0, 0, 0.0]), [0.0, 0.0], L - 1, minRes=np.full(5, 1e-3))
    assert not r2["good"] and np.isnan(r2["lastResiduals"][0]) and np.isfinite(r2["lastResiduals"][L - 1])
    assert np.array_equal(r2["T"], [1, 0, 0, 0, 0, 0, 0])               # outputs untouched on abort
