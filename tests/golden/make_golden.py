"""Generates tests/golden/*.npz.

The reference ships no golden vectors and cannot be built or imported here, so these fixtures are
REGRESSION PINS OF THE ORACLE (oracle/; the oracle itself is pinned on oracle/_ref by tests/test_ref_pin*.py), not reference outputs: they freeze the oracle's results on small
seeded inputs so that (a) any drift of the oracle is caught on CPU and (b) the CUDA path is compared against a committed
vector as well as against the live oracle.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sdv_loam_b200  # noqa
from sdv_loam_b200 import synth
import orc

WH = (640, 192); K = (383.4, 383.4, 312.0, 97.0)


def main():
    seq = synth.Sequence(2, seed=4242, K=K, wh=WH)
    w, h = WH; L = orc.lib().orc_pyr_levels(w, h)
    img0, img1 = seq.images
    pts = synth.select_points(img0, seq.clouds[0], 600, seed=1)
    p4 = np.concatenate([pts, np.full((len(pts), 1), 2e-3, np.float32)], 1).astype(np.float32)
    rh = (np.arange(len(pts)) % 3 == 0).astype(np.int32)
    f0, f1 = orc.Frame(img0, L), orc.Frame(img1, L)
    tr = orc.CoarseTracker(w, h, L, K); tr.setCoarseTrackingRef(f0, p4, rh, 0.01, -1.5)
    out = dict(img0=img0.astype(np.uint8), img1=img1.astype(np.uint8), pts4=p4, round_half=rh, K=np.array(K), ref_ab=np.array([0.01, -1.5]))
    out["pyr_checksum"] = np.array([f1.dI(l).astype(np.float64).sum() for l in range(L)])
    out["abs_checksum"] = np.array([f1.absSquaredGrad(l).astype(np.float64).sum() for l in range(L)])
    out["pc_n"] = np.array([len(tr.cloud(l)[0]) for l in range(L)])
    for l in range(L):
        u, v, idp, col = tr.cloud(l); out[f"cloud{l}"] = np.stack([u, v, idp, col])
    T = orc.se3_exp([0.01, -0.005, -0.45, 0.002, -0.003, 0.001])
    for l in range(L):
        rs = tr.calcRes(f1, l, T, 0.02, 1.0, 20.0); H, b = tr.calcGSSSE(l, T, 0.02, 1.0)
        out[f"rs{l}"] = rs; out[f"H{l}"] = H; out[f"b{l}"] = b
    out["T_eval"] = T
    r = tr.trackNewestCoarse(f1, np.array([1, 0, 0, 0, 0, 0, 0.0]), [0.0, 0.0], L - 1)
    out.update(track_good=np.array(r["good"]), track_T=r["T"], track_ab=r["ab"], track_lastRes=r["lastResiduals"], track_flow=r["flow"],
               track_iterations=r["iterations"], track_accepts=r["accepts"], track_evals=r["evals"])
    Rgt, tgt = synth.rel_pose(seq.R[0], seq.t[0], seq.R[1], seq.t[1])
    out["T_gt"] = orc.se3_from_rt(Rgt, tgt)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tracker_small.npz"), **out)
    print("track:", r["good"], orc.se3_log(r["T"]), "gt", orc.se3_log(out["T_gt"]), r["iterations"], r["accepts"], out["pc_n"])


if __name__ == "__main__":
    main()


def main_ba():
    """Back-end fixture: same window as tests/test_oracle_ba.py::window (seed-identical), frozen oracle outputs of optimize()."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import cached_sequence
    seq = cached_sequence(5, 3000, K, WH)
    win = synth.make_ba_window(seq, [0, 1, 2, 3, 4], n_per_frame=120, seed=5, pose_noise=(0.004, 0.0003), match_noise=0.15, prior_scale=1e-2)
    frames = [orc.Frame(seq.images[k], 4) for k in win["kf_idx"]]
    ba = orc.BAWindow(win, frames); r = ba.optimize(6)
    fr = ba.frames(); pts = ba.points(); rs = ba.residuals()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ba_small.npz"), iterations=r["iterations"], accepts=r["accepts"], rmse=r["rmse"],
                        T_eval=fr["T_eval"], state=fr["state"], frameEnergyTH=fr["frameEnergyTH"], idepth=pts["idepth"], res_state=rs["state"],
                        images=np.stack([seq.images[k] for k in win["kf_idx"]]).astype(np.uint8),
                        **{"win_" + k: np.asarray(v) for k, v in win.items() if k not in ("wh", "kf_idx")})
    print("ba:", r, np.bincount(rs["state"], minlength=3))


if __name__ == "__main__" and "--ba" in sys.argv or __name__ == "__main__":
    main_ba()


def pts6_of(p):
    return np.stack([p["u"], p["v"], p["idepth"], p["host"].astype(np.float32), p["obs_x"], p["obs_y"]], 1).astype(np.float32)


def main_refine():
    """structPoseEstimation fixture (§8 a11): three seeded overlap sets, frozen oracle outputs."""
    out = {}
    for k, (n, nH, seed) in enumerate([(300, 5, 2), (900, 7, 11), (40, 2, 5)]):
        d = synth.make_overlap_points(n, nH, seed, K=K, wh=WH)
        r = orc.struct_pose(WH[0], WH[1], np.array(K, np.float32), d["host_T7"], pts6_of(d["pts"]), d["T_init"])
        out.update({f"pts{k}": pts6_of(d["pts"]), f"host{k}": d["host_T7"], f"Tin{k}": d["T_init"], f"Tout{k}": r["T"],
                    f"stat{k}": np.array([r["res"], r["iterations"], r["accepts"]])})
        print("refine:", k, r["res"], r["iterations"], r["accepts"])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "refine_small.npz"), **out)


if __name__ == "__main__":
    main_refine()


def main_handover():
    """Keyframe hand-over + reprojection fixtures (§8 b9, a10): frozen oracle outputs on the small window / sequence."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import cached_sequence
    seq = cached_sequence(5, 3000, K, WH)
    win = synth.make_ba_window(seq, [0, 1, 2, 3, 4], n_per_frame=120, seed=5, pose_noise=(0.004, 0.0003), match_noise=0.15, prior_scale=1e-2)
    frames = [orc.Frame(seq.images[k], 4) for k in win["kf_idx"]]
    ba = orc.BAWindow(win, frames); ba.optimize(4)
    sel = (win["host"] == 0).astype(np.int32); sel[::7] = 1; sel[win["host"] == win["nF"] - 1] = 0
    st = ba.flagPointsForRemoval(sel); m = ba.marginalizePointsF(st); HM1, bM1 = ba.prior()
    ba.marginalizeFrame(0); HM2, bM2 = ba.prior()
    pts, hT, hab = synth.make_map(seq, [0, 1, 2, 3], n_per_frame=250, seed=2)
    cur_T7 = np.concatenate([synth._quat_from_R(seq.R[4]), seq.t[4]]); cur_T7[4:] += [0.02, -0.01, 0.03]
    order = np.random.default_rng(9).permutation(int(np.ceil(WH[0] / 25)) * int(np.ceil(WH[1] / 25))).astype(np.int32)
    idx, px = orc.reproject_map(WH[0], WH[1], 4, K, frames[:4], hT, hab, frames[4], cur_T7, [0.0, 0.0], pts, cell_order=order, max_matches=60)
    p6 = np.stack([pts["u"][idx], pts["v"][idx], pts["idepth"][idx], pts["host"][idx].astype(np.float32), px[:, 0].astype(np.float32), px[:, 1].astype(np.float32)], 1).astype(np.float32)
    sp = orc.struct_pose(WH[0], WH[1], np.array(K, np.float32), hT, p6, cur_T7)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "handover_small.npz"), sel=sel, status=st, M=m["M"], Msc=m["Msc"], HM1=HM1, bM1=bM1, HM2=HM2, bM2=bM2,
                        map_pts=np.stack([pts["u"], pts["v"], pts["idepth"], pts["host"].astype(np.float32), pts["type"].astype(np.float32)], 1), map_T7=hT, cur_T7=cur_T7, order=order,
                        match_idx=idx, match_px=px, refined_T7=sp["T"], refine_stats=np.array([sp["iterations"], sp["accepts"]]))
    print("handover:", np.bincount(st, minlength=3), len(idx), sp["iterations"], sp["accepts"])


if __name__ == "__main__":
    main_handover()
