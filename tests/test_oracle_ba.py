"""Oracle back-end (oracle/orc_ba.*) against independent float64 numpy formulations — CPU only.
These pin the restated formulas without reference goldens (there are none): numerical differentiation of
the projection for the Jacobians, dense assembly of the normal equations, and the explicit (un-marginalised) system for the Schur complement."""
import os
import numpy as np
import pytest
import orc
from conftest import SMALL_K, SMALL_WH, cached_sequence

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ba_small.npz")


@pytest.fixture(scope="module")
def window():
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import synth
    seq = cached_sequence(5, 3000, SMALL_K, SMALL_WH)
    win = synth.make_ba_window(seq, [0, 1, 2, 3, 4], n_per_frame=120, seed=5, pose_noise=(0.004, 0.0003), match_noise=0.15, prior_scale=1e-2)
    frames = [orc.Frame(seq.images[k], 4) for k in win["kf_idx"]]
    return seq, win, frames


def R_of(T7): return orc.se3_rot(T7)


def project(win, T_h, T_t, uv, idepth, K):
    fx, fy, cx, cy = K
    rel = orc.se3_mul(T_t, orc.se3_inv(T_h)); R = R_of(rel); t = rel[4:]
    p = R @ np.array([(uv[0] - cx) / fx, (uv[1] - cy) / fy, 1.0]) + t * idepth
    return np.array([fx * p[0] / p[2] + cx, fy * p[1] / p[2] + cy])


def test_adjoints_and_precalc(window):
    seq, win, frames = window
    ba = orc.BAWindow(win, frames)
    for h, t in ((0, 1), (2, 4), (4, 1)):
        pc = ba.precalc(h, t)
        rel = orc.se3_mul(win["T_eval"][t], orc.se3_inv(win["T_eval"][h]))
        AH = -orc.se3_adj(rel).T; AH[:3] *= 0.5
        AT = np.eye(6); AT[:3] *= 0.5
        assert np.allclose(pc["adHost"], AH, atol=1e-12) and np.allclose(pc["adTarget"], AT)      # EnergyFunctional.cpp:36-48
        K = np.array([[win["K"][0], 0, win["K"][2]], [0, win["K"][1], win["K"][3]], [0, 0, 1]])
        assert np.allclose(pc["KRKi"], K @ R_of(rel) @ np.linalg.inv(K), rtol=2e-5, atol=1e-4)
        assert np.allclose(pc["Kt"], K @ rel[4:], rtol=1e-5, atol=1e-4) and np.allclose(pc["R0"], R_of(rel), atol=1e-6) and np.allclose(pc["aff"], [1, 0])


def test_linearize_jacobians_by_finite_differences(window):
    """Jpdxi / Jpdd / Jpdc (Residuals.cpp:99-134) are d(Ku,Kv)/d(left pose increment of host->target), d/d(idepth), d/d(calib*SCALE)."""
    seq, win, frames = window
    ba = orc.BAWindow(win, frames); ba.reset_oob(); ba.linearizeAll(False)
    rs = ba.residuals(); K = win["K"]
    rng = np.random.default_rng(0); ok = np.where((rs["new_state"] != 1))[0]
    for r in rng.choice(ok, 25, replace=False):
        p, h, t = win["r_point"][r], win["r_host"][r], win["r_target"][r]
        uv, idp = win["uv"][p].astype(np.float64), float(win["idepth"][p])
        J = rs["J"][r].astype(np.float64)
        base = project(win, win["T_eval"][h], win["T_eval"][t], uv, idp, K)
        res = base - win["r_matcher"][r]
        nrm = np.linalg.norm(res); hw = 1.0 if nrm < 6 else np.sqrt(6 / nrm)
        assert np.allclose(J[:2], res * hw, rtol=2e-4, atol=2e-3)
        assert np.allclose(rs["center"][r][:2], base, atol=5e-3)
        rel = orc.se3_mul(win["T_eval"][t], orc.se3_inv(win["T_eval"][h]))
        def proj_rel(relT, uv_, idp_, K_):
            fx, fy, cx, cy = K_; R = R_of(relT); tt = relT[4:]
            q = R @ np.array([(uv_[0] - cx) / fx, (uv_[1] - cy) / fy, 1.0]) + tt * idp_
            return np.array([fx * q[0] / q[2] + cx, fy * q[1] / q[2] + cy])
        Jxi = np.zeros((2, 6))
        for k in range(6):
            e = np.zeros(6); e[k] = 1e-6
            Jxi[:, k] = (proj_rel(orc.se3_mul(orc.se3_exp(e), rel), uv, idp, K) - proj_rel(orc.se3_mul(orc.se3_exp(-e), rel), uv, idp, K)) / 2e-6
        assert np.allclose(np.stack([J[2:8], J[8:14]]) / hw, Jxi, rtol=2e-3, atol=2e-2)
        Jd = (proj_rel(rel, uv, idp + 1e-7, K) - proj_rel(rel, uv, idp - 1e-7, K)) / 2e-7
        assert np.allclose(J[22:24] / hw, Jd, rtol=2e-3, atol=1e-2)
        Jc = np.zeros((2, 4))
        for k in range(4):
            dK = np.zeros(4); dK[k] = 1e-4
            Jc[:, k] = (proj_rel(rel, uv, idp, K + dK) - proj_rel(rel, uv, idp, K - dK)) / 2e-4 * 50.0      # SCALE_F / SCALE_C
        assert np.allclose(np.stack([J[14:18], J[18:22]]) / hw, Jc, rtol=5e-3, atol=5e-2)


def _dense_system(win, rs, pts_isSensor, ba):
    """float64 assembly of the full (frames+calib+idepth) normal equations from the per-residual Jacobians."""
    nF = win["nF"]; n = 4 + 6 * nF; nP = len(win["uv"])
    act = np.where(rs["active"] == 1)[0]
    Jf = np.zeros((2 * len(act), n)); Jd = np.zeros((2 * len(act), nP)); rv = np.zeros(2 * len(act))
    for i, r in enumerate(act):
        J = rs["efJ"][r].astype(np.float64); h, t, p = win["r_host"][r], win["r_target"][r], win["r_point"][r]
        pc = ba.precalc(h, t)
        for row, (jx, jc, jd, rr) in enumerate(((J[2:8], J[14:18], J[22], J[0]), (J[8:14], J[18:22], J[23], J[1]))):
            Jf[2 * i + row, :4] = jc
            Jf[2 * i + row, 4 + 6 * h:10 + 6 * h] += pc["adHost"] @ jx
            Jf[2 * i + row, 4 + 6 * t:10 + 6 * t] += pc["adTarget"] @ jx
            Jd[2 * i + row, p] = jd; rv[2 * i + row] = rr
    return Jf, Jd, rv, act


def test_accumulate_and_schur_against_dense_algebra(window):
    seq, win, frames = window
    ba = orc.BAWindow(win, frames); ba.reset_oob(); ba.linearizeAll(False); ba.applyRes()
    HA, bA, Hsc, bsc = ba.accumulate()
    rs = ba.residuals(); Jf, Jd, rv, act = _dense_system(win, rs, win["isFromSensor"], ba)
    n = HA.shape[0]
    prior = np.zeros(n); prior[:4] = 5e9; prior[4:7] = 1e10; prior[7:10] = 1e11                 # cPrior + frame-0 prior (settings.cpp:23-27)
    Hd = Jf.T @ Jf + np.diag(prior)
    assert np.allclose(HA, Hd, rtol=2e-4, atol=1e-6 * np.abs(Hd).max())                          # AccumulatedTopHessian stitch
    assert np.allclose(bA, Jf.T @ rv, rtol=2e-4, atol=1e-5 * np.abs(bA).max())
    assert np.allclose(HA, HA.T)
    # Schur complement over the free (non-sensor) inverse depths with the idepth prior on the diagonal
    free = np.where((win["isFromSensor"] == 0) & (np.bincount(win["r_point"][act], minlength=len(win["uv"])) > 0))[0]
    Hdd = (Jd[:, free] ** 2).sum(0) + np.where(win["hasDepthPrior"][free] == 1, 2500.0, 0.0)
    Hfd = Jf.T @ Jd[:, free]; bd = Jd[:, free].T @ rv
    assert np.allclose(Hsc, (Hfd / Hdd) @ Hfd.T, rtol=5e-4, atol=1e-6 * np.abs(Hsc).max())      # AccumulatedSCHessian stitch
    assert np.allclose(bsc, (Hfd / Hdd) @ bd, rtol=5e-4, atol=1e-5 * np.abs(bsc).max())


def test_solve_resubstitute_and_orthogonalize(window):
    seq, win, frames = window
    ba = orc.BAWindow(win, frames); ba.reset_oob(); ba.linearizeAll(False); ba.applyRes(); ba.backupState()
    x, HS, bS = ba.solveSystem(0, 0.1)
    n = len(x); Hl = HS.copy(); Hl[np.diag_indices(n)] *= 1.1
    assert np.allclose(x, np.linalg.solve(Hl, bS), rtol=1e-7, atol=1e-12)                         # EnergyFunctional.cpp:741-743
    fr = ba.frames(); assert np.allclose(fr["step"][:, :6].reshape(-1), -x[4:]) and np.allclose(ba.calib()[1], -x[:4])
    pts = ba.points(); assert np.all(pts["step"][win["isFromSensor"] == 1] == 0)                  # LiDAR points keep their depth
    x2, _, _ = ba.solveSystem(2, 0.1)                                                             # iteration >= 2: x made orthogonal to gauge directions
    assert np.linalg.norm(x2 - x) > 0
    # nullspace of a left-multiplied rigid motion / scale: check orthogonality through the difference being in span(N)
    d = x - x2
    x3, _, _ = ba.solveSystem(3, 0.1)
    assert np.allclose(x3, x2, atol=1e-14)
    assert abs(d @ x2) < 1e-8 * np.linalg.norm(d) * np.linalg.norm(x2) + 1e-18                    # projector: removed part is orthogonal to what is kept


def test_optimize_reduces_energy_and_golden_pin(window):
    seq, win, frames = window
    ba = orc.BAWindow(win, frames)
    ba.reset_oob(); e0 = ba.linearizeAll(False)
    ba2 = orc.BAWindow(win, frames); r = ba2.optimize(6)
    assert 1 <= r["iterations"] <= 6 and r["accepts"] >= 1 and np.isfinite(r["rmse"])
    g = np.load(GOLD)
    fr = ba2.frames(); pts = ba2.points(); rs = ba2.residuals()
    assert r["iterations"] == int(g["iterations"]) and r["accepts"] == int(g["accepts"]) and np.isclose(r["rmse"], float(g["rmse"]), rtol=1e-6)
    assert np.allclose(fr["T_eval"], g["T_eval"], atol=1e-12) and np.allclose(fr["state"], g["state"], atol=1e-12)
    assert np.allclose(pts["idepth"], g["idepth"], atol=1e-9) and np.array_equal(rs["state"], g["res_state"])
    assert e0 > 0
