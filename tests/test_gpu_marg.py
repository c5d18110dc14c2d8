"""GPU parity (-m gpu) of the keyframe hand-over: sdv_ba_flag_points / sdv_ba_marginalize_points / sdv_ba_marginalize_frame
against the oracle.  Float accumulators (reference tiers and order) and every fp64 sum are evaluated in the oracle's order -> BIT-EXACT."""
import numpy as np
import pytest
import orc
from conftest import cached_sequence

pytestmark = pytest.mark.gpu


def _mods():
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import api, synth
    return api, synth


def _pair(api, synth, seq, kfs, **kw):
    win = synth.make_ba_window(seq, kfs, **kw)
    L = api.pyr_levels(*synth.KITTI_WH)
    frames = [orc.Frame(seq.images[k], L) for k in kfs]
    ctx = api.Context(synth.KITTI_K, *synth.KITTI_WH, max_frames=len(kfs) + 1)
    for i, k in enumerate(kfs):
        ctx.makeImages(500 + i, seq.images[k])
    return win, ctx, orc.BAWindow(win, frames), api.EnergyFunctional(ctx, win, [500 + i for i in range(len(kfs))])


def _same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def _select(win, marg_host, every):
    sel = (win["host"] == marg_host).astype(np.int32); sel[::every] = 1
    sel[win["host"] == win["nF"] - 1] = 0
    return sel


@pytest.mark.parametrize("cfg", [dict(kfs=[0, 1, 2, 3, 4, 5, 6], n_per_frame=250, prior_scale=1e-3, marg=0, every=5),
                                 dict(kfs=[0, 1, 2, 3, 4, 5, 6, 7], n_per_frame=400, prior_scale=1e-2, sensor_frac=0.3, marg=3, every=3),
                                 dict(kfs=[1, 3, 5], n_per_frame=150, prior_scale=0.0, sensor_frac=0.0, marg=1, every=2),
                                 dict(kfs=[0, 2, 4, 6], n_per_frame=200, prior_scale=1e-3, sensor_frac=1.0, marg=3, every=4)])
def test_keyframe_handover_bit_exact(cfg):
    api, synth = _mods()
    seq = cached_sequence(8, 2000, synth.KITTI_K, synth.KITTI_WH)
    kfs, marg, every = cfg.pop("kfs"), cfg.pop("marg"), cfg.pop("every")
    win, ctx, ob, gb = _pair(api, synth, seq, kfs, seed=3, pose_noise=(0.005, 0.0003), match_noise=0.1, **cfg)
    ro, rg = ob.optimize(4), gb.optimize(4)
    assert (ro["iterations"], ro["accepts"]) == (rg["iterations"], rg["accepts"])
    sel = _select(win, marg, every)
    so, sg = ob.flagPointsForRemoval(sel), gb.flagPointsForRemoval(sel)
    assert _same(so, sg) and (so == 2).sum() > 0
    (zo, lo), (zg, lg) = ob.res_to_zero(), gb.linearized()
    assert _same(lo, lg) and _same(zo[lo == 1], zg[lg == 1])
    a, b = ob.residuals(), gb.residuals()
    assert _same(a["active"], b["active"]) and _same(a["efJ"], b["efJ"]) and _same(a["state"], b["state"])
    mo = ob.marginalizePointsF(so); mg = gb.marginalizePointsF()                 # device-resident status of flag_points
    for k in ("M", "Mb", "Msc", "Mbsc"):
        assert _same(mo[k], mg[k]), k
    (Ho, bo), (Hg, bg) = ob.prior(), gb.prior()
    assert _same(Ho, Hg) and _same(bo, bg)
    po, pg = ob.points(), gb.points()
    assert _same(po["idepth_hessian"][so == 2], pg["idepth_hessian"][so == 2]) and _same(po["HdiF"][so == 2], pg["HdiF"][so == 2])
    ob.marginalizeFrame(marg); gb.marginalizeFrame(marg)
    (Ho, bo), (Hg, bg) = ob.prior(), gb.prior()
    assert Ho.shape == Hg.shape == (4 + 6 * (len(kfs) - 1),) * 2
    assert _same(Ho, Hg) and _same(bo, bg)
    if len(kfs) > 3:                                                             # a second frame, now at a shifted index
        ob.marginalizeFrame(0); gb.marginalizeFrame(0)
        (Ho, bo), (Hg, bg) = ob.prior(), gb.prior()
        assert _same(Ho, Hg) and _same(bo, bg)
    ctx.close()


def test_marginalize_points_host_status_and_errors():
    api, synth = _mods()
    seq = cached_sequence(8, 2000, synth.KITTI_K, synth.KITTI_WH)
    win, ctx, ob, gb = _pair(api, synth, seq, [0, 1, 2, 3, 4], seed=9, n_per_frame=150, pose_noise=(0.004, 0.0003), match_noise=0.1, prior_scale=1e-3)
    ob.optimize(3); gb.optimize(3)
    sel = _select(win, 0, 6); so = ob.flagPointsForRemoval(sel); sg = gb.flagPointsForRemoval(sel)
    st = so.copy(); st[np.where(st == 2)[0][::3]] = 1                            # the host may still demote points (PS_DROP) before marginalising
    mo = ob.marginalizePointsF(st); mg = gb.marginalizePointsF(st)
    assert _same(mo["M"], mg["M"]) and _same(mo["Msc"], mg["Msc"]) and _same(ob.prior()[0], gb.prior()[0])
    with pytest.raises(api.SdvError):
        gb.marginalizeFrame(9)
    gb.marginalizeFrame(0)
    with pytest.raises(api.SdvError):
        gb.flagPointsForRemoval(sel[:1])                                         # window has no points until sdv_ba_set_points is called again
    ctx.close()


def test_marginalize_nothing_and_two_frame_window():
    """No point reaches PS_MARGINALIZE -> HM,bM unchanged (M = Msc = 0); frame elimination down to a single keyframe."""
    api, synth = _mods()
    seq = cached_sequence(8, 2000, synth.KITTI_K, synth.KITTI_WH)
    win, ctx, ob, gb = _pair(api, synth, seq, [0, 3], seed=4, n_per_frame=120, pose_noise=(0.004, 0.0003), match_noise=0.1, prior_scale=1e-3)
    ob.optimize(3); gb.optimize(3)
    H0, b0 = gb.prior()
    st = np.zeros(len(win["uv"]), np.int32); st[::4] = 1                          # only PS_DROP
    mo = ob.marginalizePointsF(st); mg = gb.marginalizePointsF(st)
    assert not mg["M"].any() and not mg["Msc"].any() and _same(mo["M"], mg["M"])
    H1, b1 = gb.prior(); assert _same(H0, H1) and _same(b0, b1)
    ob.marginalizeFrame(0); gb.marginalizeFrame(0)
    (Ho, bo), (Hg, bg) = ob.prior(), gb.prior()
    assert Hg.shape == (10, 10) and _same(Ho, Hg) and _same(bo, bg)
    with pytest.raises(api.SdvError):
        gb.marginalizeFrame(0)                                                    # a window keeps at least one keyframe
    ctx.close()


def test_golden_handover_on_gpu():
    """CUDA path against the committed fixture (tests/golden/handover_small.npz): marginalisation prior before/after frame elimination, reprojection matches,
    refined pose — without running the oracle."""
    import os
    api, synth = _mods()
    from conftest import SMALL_K, SMALL_WH
    here = os.path.dirname(os.path.abspath(__file__))
    g = np.load(os.path.join(here, "golden", "handover_small.npz")); b = np.load(os.path.join(here, "golden", "ba_small.npz"))
    win = {k[4:]: b[k] for k in b.files if k.startswith("win_")}; win["nF"] = int(win["nF"]); win["wh"] = SMALL_WH; win["kf_idx"] = list(range(win["nF"]))
    w, h = SMALL_WH; ctx = api.Context(SMALL_K, w, h, max_frames=8)
    for k in range(5):
        ctx.makeImages(500 + k, b["images"][k].astype(np.float32))
    gb = api.EnergyFunctional(ctx, win, [500 + k for k in range(5)]); gb.optimize(4)
    st = gb.flagPointsForRemoval(g["sel"]); assert np.array_equal(st, g["status"])
    m = gb.marginalizePointsF(); assert _same(m["M"], g["M"]) and _same(m["Msc"], g["Msc"])
    H1, b1 = gb.prior(); assert _same(H1, g["HM1"]) and _same(b1, g["bM1"])
    gb.marginalizeFrame(0); H2, b2 = gb.prior(); assert _same(H2, g["HM2"]) and _same(b2, g["bM2"])
    mp = g["map_pts"]; pts = np.zeros(len(mp), api.MAP_PT_DTYPE)
    pts["u"], pts["v"], pts["idepth"], pts["host"], pts["type"] = mp[:, 0], mp[:, 1], mp[:, 2], mp[:, 3].astype(np.int32), mp[:, 4].astype(np.int32)
    rp = api.Reprojector(ctx); rp.setMap(0, [500, 501, 502, 503], g["map_T7"], None, pts)
    idx, px = rp.reprojectMap(0, 504, g["cur_T7"], cell_order=g["order"], max_matches=60)
    assert np.array_equal(idx, g["match_idx"]) and np.array_equal(px, g["match_px"])
    r = rp.refineBatch([0], [504], g["cur_T7"][None], cell_order=g["order"], max_matches=60)
    assert (int(r["iterations"][0]), int(r["accepts"][0])) == tuple(int(x) for x in g["refine_stats"]) and np.abs(r["T"][0] - g["refined_T7"]).max() < 1e-9
    ctx.close()
