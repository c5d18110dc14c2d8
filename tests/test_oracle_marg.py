"""Oracle keyframe hand-over (marginalizePointsF / marginalizeFrame / flagPointsForRemoval numeric part) against
independent float64 numpy formulations — CPU only."""
import numpy as np
import pytest
import orc
from conftest import SMALL_K, SMALL_WH, cached_sequence


@pytest.fixture(scope="module")
def window():
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import synth
    seq = cached_sequence(5, 3000, SMALL_K, SMALL_WH)
    win = synth.make_ba_window(seq, [0, 1, 2, 3, 4], n_per_frame=120, seed=5, pose_noise=(0.004, 0.0003), match_noise=0.15, prior_scale=1e-2)
    frames = [orc.Frame(seq.images[k], 4) for k in win["kf_idx"]]
    return seq, win, frames


def select(win, host_to_marg=0, extra_every=7):
    sel = (win["host"] == host_to_marg).astype(np.int32)
    sel[::extra_every] = 1
    sel[win["host"] == win["nF"] - 1] = 0                   # FullSystem.cpp:751: the newest keyframe's points are never touched
    return sel


def test_flag_points_fix_linearization(window):
    seq, win, frames = window
    ba = orc.BAWindow(win, frames); ba.optimize(4)
    sel = select(win); pts_before = ba.points()
    st = ba.flagPointsForRemoval(sel)
    assert set(np.unique(st[sel == 1])) <= {1, 2} and np.all(st[sel == 0] == 0)
    assert np.array_equal(st[sel == 1] == 2, pts_before["idepth_hessian"][sel == 1] > 50)
    r2z, lin = ba.res_to_zero(); rs = ba.residuals()
    selr = sel[win["r_point"]] == 1
    assert np.array_equal(lin == 1, selr & (rs["active"] == 1))
    # res_toZeroF = resF - J*delta  (EnergyFunctionalStructs.cpp:46-55)
    J = rs["efJ"]; k = np.where(lin == 1)[0]
    for r in k[:40]:
        p, h, t = win["r_point"][r], win["r_host"][r], win["r_target"][r]
        pc = ba.precalc(h, t); cal_v, _ = ba.calib()
        dd = ba.points()["idepth"][p] - win["idepth_zero"][p]
        # deltaF is idepth - idepth_zero at window set-up; doStepFromBackup keeps idepth_zero == idepth afterwards -> 0 here (EFPoint::takeData only at insert)
        d = J[r, 2:8] @ pc["adHTdelta"], J[r, 8:14] @ pc["adHTdelta"]
        assert abs((J[r, 0] - r2z[r, 0]) - d[0]) < 1e-3 + abs(J[r, 22] * dd) and abs((J[r, 1] - r2z[r, 1]) - d[1]) < 1e-3 + abs(J[r, 23] * dd)


def dense_marg_system(win, ba, status):
    """Explicit float64 normal equations over [calib(4), frames(6nF), idepths of the marg points] from efJ/res_toZeroF, then a numpy Schur complement."""
    rs = ba.residuals(); r2z, _ = ba.res_to_zero(); nF = win["nF"]; N = 4 + 6 * nF
    mp = np.where(status == 2)[0]; col = {p: N + i for i, p in enumerate(mp)}
    H = np.zeros((N + len(mp), N + len(mp))); b = np.zeros(N + len(mp))
    prior = np.where(win["hasDepthPrior"] == 1, 2500.0 * 600 * 600, 0.0)
    for p in mp:
        if win["isFromSensor"][p]:
            continue                                                         # LiDAR points never reach the Schur accumulators (AccumulatedSCHessian.cpp:38-39)
    for r in range(len(rs["active"])):
        p = win["r_point"][r]
        if status[p] != 2 or not rs["active"][r]:
            continue
        h, t = win["r_host"][r], win["r_target"][r]; pc = ba.precalc(h, t); J = rs["efJ"][r].astype(np.float64)
        Jr = np.zeros((2, N + len(mp)))
        for a in range(2):
            Jx = J[2 + 6 * a: 8 + 6 * a]; Jc = J[14 + 4 * a: 18 + 4 * a]
            Jr[a, :4] = Jc
            Jr[a, 4 + 6 * h: 10 + 6 * h] += pc["adHost"] @ Jx
            Jr[a, 4 + 6 * t: 10 + 6 * t] += pc["adTarget"] @ Jx
            Jr[a, col[p]] = J[22 + a]
        H += Jr.T @ Jr; b += Jr.T @ r2z[r].astype(np.float64)
    for p in mp:
        H[col[p], col[p]] += prior[p]
    return H, b, mp, col


def test_marginalize_points_matches_dense_schur(window):
    seq, win, frames = window
    ba = orc.BAWindow(win, frames); ba.optimize(4)
    st = ba.flagPointsForRemoval(select(win)); assert (st == 2).sum() > 30
    HM0, bM0 = ba.prior()
    out = ba.marginalizePointsF(st)
    HM1, bM1 = ba.prior(); N = len(bM0)
    H, b, mp, col = dense_marg_system(win, ba, st)
    # top part: M equals the frame/calib block of the dense system
    scale = np.abs(H[:N, :N]).max()
    assert np.allclose(out["M"], H[:N, :N], rtol=1e-4, atol=1e-5 * scale) and np.allclose(out["Mb"], b[:N], rtol=1e-4, atol=1e-5 * np.abs(b[:N]).max())
    # Schur part: only vision points are eliminated against frames; LiDAR points contribute nothing to Msc
    vis = [p for p in mp if not win["isFromSensor"][p]]
    S = np.zeros((N, N)); Sb = np.zeros(N)
    for p in vis:
        c = col[p]; S += np.outer(H[:N, c], H[:N, c]) / H[c, c]; Sb += H[:N, c] * b[c] / H[c, c]
    assert np.allclose(out["Msc"], S, rtol=1e-3, atol=1e-5 * scale) and np.allclose(out["Mbsc"], Sb, rtol=1e-3, atol=1e-5 * np.abs(b[:N]).max() + 1e-6)
    assert np.allclose(HM1 - HM0, 0.25 * (out["M"] - out["Msc"]), rtol=1e-12, atol=1e-9 * scale) and np.allclose(bM1 - bM0, 0.25 * (out["Mb"] - out["Mbsc"]), atol=1e-9 * (1 + np.abs(out["Mb"]).max()))
    D = out["M"] - out["Msc"]; assert np.allclose(D, D.T, atol=1e-6 * scale)
    assert np.linalg.eigvalsh(0.5 * (D + D.T)).min() > -1e-6 * scale         # a Schur complement of a PSD system is PSD


@pytest.mark.parametrize("idx", [0, 2, 4])
def test_marginalize_frame_is_a_schur_complement(window, idx):
    seq, win, frames = window
    ba = orc.BAWindow(win, frames); ba.optimize(2)
    st = ba.flagPointsForRemoval(select(win, host_to_marg=idx)); ba.marginalizePointsF(st)
    HM, bM = ba.prior(); N = len(bM)
    keep = [i for i in range(N) if not (4 + 6 * idx <= i < 10 + 6 * idx)]; drop = list(range(4 + 6 * idx, 10 + 6 * idx))
    prior = np.zeros(6)
    if win["frameID"][idx] == 0:
        prior[:3] = 1e10; prior[3:] = 1e11                                   # FrameHessian::getPrior (first frame): setting_initialTransPrior / RotPrior
    dprior = ba.frames()["state"][idx][:6]                                  # delta_prior = (state - priorZero).head<6>()  (EnergyFunctionalStructs.cpp:28-35)
    Dm = HM[np.ix_(drop, drop)] + np.diag(prior); B = HM[np.ix_(keep, drop)]
    ref = HM[np.ix_(keep, keep)] - B @ np.linalg.solve(Dm, B.T); refb = bM[keep] - B @ np.linalg.solve(Dm, bM[drop] + prior * dprior)
    ba.marginalizeFrame(idx)
    H2, b2 = ba.prior()
    assert H2.shape == (N - 6, N - 6) and np.array_equal(H2, H2.T)
    sc = np.abs(ref).max()
    assert np.allclose(H2, 0.5 * (ref + ref.T), rtol=1e-7, atol=1e-9 * sc) and np.allclose(b2, refb, rtol=1e-5, atol=1e-7 * (1 + np.abs(refb).max()))   # cancellation: cond(D) ~ 1e5


def _golden_setup():
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "handover_small.npz"))
    b = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ba_small.npz"))
    win = {k[4:]: b[k] for k in b.files if k.startswith("win_")}; win["nF"] = int(win["nF"]); win["wh"] = SMALL_WH; win["kf_idx"] = list(range(win["nF"]))
    imgs = [b["images"][k].astype(np.float32) for k in range(5)]
    return g, win, imgs


def test_golden_handover_regression():
    """tests/golden/handover_small.npz pins the oracle's keyframe hand-over and reprojection outputs (make_golden.py::main_handover) against drift."""
    g, win, imgs = _golden_setup()
    frames = [orc.Frame(im, 4) for im in imgs]
    ba = orc.BAWindow(win, frames); ba.optimize(4)
    st = ba.flagPointsForRemoval(g["sel"]); assert np.array_equal(st, g["status"])
    m = ba.marginalizePointsF(st); assert np.array_equal(m["M"], g["M"]) and np.array_equal(m["Msc"], g["Msc"])
    H1, b1 = ba.prior(); assert np.array_equal(H1, g["HM1"]) and np.array_equal(b1, g["bM1"])
    ba.marginalizeFrame(0); H2, b2 = ba.prior(); assert np.array_equal(H2, g["HM2"]) and np.array_equal(b2, g["bM2"])
    mp = g["map_pts"]; pts = np.zeros(len(mp), [("u", np.float32), ("v", np.float32), ("idepth", np.float32), ("host", np.int32), ("type", np.int32)])
    pts["u"], pts["v"], pts["idepth"], pts["host"], pts["type"] = mp[:, 0], mp[:, 1], mp[:, 2], mp[:, 3].astype(np.int32), mp[:, 4].astype(np.int32)
    idx, px = orc.reproject_map(SMALL_WH[0], SMALL_WH[1], 4, SMALL_K, frames[:4], g["map_T7"], np.zeros((4, 2)), frames[4], g["cur_T7"], [0.0, 0.0], pts,
                                cell_order=g["order"], max_matches=60)
    assert np.array_equal(idx, g["match_idx"]) and np.array_equal(px, g["match_px"])
