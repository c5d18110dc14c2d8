"""GPU parity tests of the back-end (-m gpu): CUDA path through the C-ABI vs the CPU oracle.

The accumulation kernels walk items in the reference's order with the reference's float tiers and every fp64 reduction is
evaluated in the oracle's order, so the bar here is BIT-EXACT for every intermediate (J, energies, H, b, Schur, x, steps) and for the
outcome of the whole FullSystem::optimize loop — far inside north_star's 1e-4 energy / 1e-3 pose tolerance."""
import os
import numpy as np
import pytest
import orc
from conftest import SMALL_K, SMALL_WH, cached_sequence

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ba_small.npz")


def _mods():
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import api, synth
    return api, synth


def _pair(api, synth, seq, K, wh, kfs, **kw):
    win = synth.make_ba_window(seq, kfs, **kw)
    L = api.pyr_levels(*wh)
    frames = [orc.Frame(seq.images[k], L) for k in kfs]
    ctx = api.Context(K, wh[0], wh[1], max_frames=len(kfs) + 1)
    for i, k in enumerate(kfs):
        ctx.makeImages(500 + i, seq.images[k])
    return win, frames, ctx, orc.BAWindow(win, frames), api.EnergyFunctional(ctx, win, [500 + i for i in range(len(kfs))])


def _same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


@pytest.mark.parametrize("cfg", [dict(kfs=[0, 1, 2, 3, 4, 5, 6], n_per_frame=250, prior_scale=1e-3), dict(kfs=[0, 2, 4], n_per_frame=150, prior_scale=0.0),
                                 dict(kfs=[0, 1, 2, 3, 4, 5, 6, 7], n_per_frame=120, prior_scale=1e-2, sensor_frac=0.0), dict(kfs=[1, 3], n_per_frame=80, sensor_frac=1.0)])
def test_stepwise_bit_exact(cfg):
    """T6/T7/T8: linearize, applyRes, energies, accumulate (top + Schur), solve, resubstitute, step — every array identical."""
    api, synth = _mods()
    seq = cached_sequence(8, 2000, synth.KITTI_K, synth.KITTI_WH)
    kfs = cfg.pop("kfs")
    win, frames, ctx, ob, gb = _pair(api, synth, seq, synth.KITTI_K, synth.KITTI_WH, kfs, seed=3, pose_noise=(0.005, 0.0003), match_noise=0.1, **cfg)
    nF = len(kfs)
    for h, t in ((0, nF - 1), (nF - 1, 0)):
        po, pg = ob.precalc(h, t), gb.precalc(h, t)
        assert all(_same(po[k], pg[k]) for k in po), (h, t)
    ob.reset_oob(); gb.reset_oob()
    assert ob.linearizeAll(False) == gb.linearizeAll(False)
    ro, rg = ob.residuals(), gb.residuals()
    assert _same(ro["new_state"], rg["new_state"]) and _same(ro["J"], rg["J"]) and _same(ro["energies"].astype(np.float32), rg["energies"]) and _same(ro["center"], rg["center"])
    assert _same(ob.frames()["frameEnergyTH"], gb.frames()["frameEnergyTH"])                       # nth_element threshold
    assert (ob.calcLEnergy(), ob.calcMEnergy()) == gb.energies()
    ob.applyRes(); gb.applyRes()
    ro, rg = ob.residuals(), gb.residuals()
    assert _same(ro["active"], rg["active"]) and _same(ro["efJ"], rg["efJ"]) and _same(ro["JpJdF"], rg["JpJdF"]) and _same(ro["state"], rg["state"])
    ob.backupState(); gb.backupState()
    for it, lam in ((0, 0.1), (1, 0.025), (2, 0.00625)):
        HAo, bAo, Hsco, bsco = ob.accumulate(); xo, HSo, bSo = ob.solveSystem(it, lam)
        xg, HSg, bSg, (HAg, bAg, Hscg, bscg) = gb.solveSystem(it, lam)
        assert _same(HAo, HAg) and _same(bAo, bAg) and _same(Hsco, Hscg) and _same(bsco, bscg) and _same(HSo, HSg) and _same(bSo, bSg)
        assert _same(xo, xg), (it, np.abs(xo - xg).max())
        po, pg = ob.points(), gb.points()
        assert all(_same(po[k], pg[k]) for k in ("HdiF", "bdSumF", "step", "idepth_hessian"))
        assert _same(ob.frames()["step"], gb.frames()["step"])
    assert ob.doStepFromBackup(1.0) == gb.doStepFromBackup(1.0)
    fo, fg = ob.frames(), gb.frames()
    assert _same(fo["state"], fg["state"]) and _same(fo["PRE_worldToCam"], fg["PRE_worldToCam"]) and _same(ob.points()["idepth"], gb.points()["idepth"])
    assert ob.linearizeAll(False) == gb.linearizeAll(False)
    ob.loadStateBackup(); gb.loadStateBackup()
    assert _same(ob.frames()["state"], gb.frames()["state"]) and ob.linearizeAll(True) == gb.linearizeAll(True)
    ro, rg = ob.residuals(), gb.residuals()
    assert _same(ro["toRemove"], rg["toRemove"]) and _same(ro["state"], rg["state"])
    po, pg = ob.points(), gb.points()
    assert _same(po["maxRelBaseline"], pg["maxRelBaseline"]) and _same(po["numGood"], pg["numGood"])
    ctx.close()


@pytest.mark.parametrize("nkf,noise", [(7, (0.005, 0.0003)), (7, (0.02, 0.002)), (5, (0.0, 0.0)), (3, (0.01, 0.0005)), (2, (0.003, 0.0002))])
def test_optimize_matches_oracle(nkf, noise):
    """T9: FullSystem::optimize — same iteration count and accept pattern, identical final poses / depths / residual states."""
    api, synth = _mods()
    seq = cached_sequence(8, 2000, synth.KITTI_K, synth.KITTI_WH)
    win, frames, ctx, ob, gb = _pair(api, synth, seq, synth.KITTI_K, synth.KITTI_WH, list(range(nkf)), n_per_frame=250, seed=11, pose_noise=noise, match_noise=0.2, prior_scale=1e-3)
    ro, rg = ob.optimize(6), gb.optimize(6)
    assert (ro["iterations"], ro["accepts"]) == (rg["iterations"], rg["accepts"]) and ro["rmse"] == rg["rmse"]
    fo, fg = ob.frames(), gb.frames()
    assert _same(fo["T_eval"], fg["T_eval"]) and _same(fo["state"], fg["state"]) and _same(fo["frameEnergyTH"], fg["frameEnergyTH"]) and _same(ob.calib()[0], fg["calib_value"])
    assert _same(ob.points()["idepth"], gb.points()["idepth"])
    so, sg = ob.residuals(), gb.residuals()
    assert _same(so["state"], sg["state"]) and _same(so["toRemove"], sg["toRemove"]) and _same(so["active"], sg["active"])
    # north_star tolerances, for the record: pose 1e-3 m / 1e-3 rad, energy 1e-4 rel
    for i in range(nkf):
        e = orc.se3_log(orc.se3_mul(fg["PRE_worldToCam"][i], orc.se3_inv(fo["PRE_worldToCam"][i])))
        assert np.abs(e).max() < 1e-3
    ctx.close()


def test_golden_fixture_on_gpu():
    api, synth = _mods()
    g = np.load(GOLD)
    win = {k[4:]: g[k] for k in g.files if k.startswith("win_")}
    win["nF"] = int(win["nF"]); win["wh"] = SMALL_WH
    w, h = SMALL_WH
    ctx = api.Context(tuple(win["K"]), w, h, max_frames=win["nF"] + 1)
    for i in range(win["nF"]):
        ctx.makeImages(i, g["images"][i].astype(np.float32))
    gb = api.EnergyFunctional(ctx, win, list(range(win["nF"])))
    r = gb.optimize(6)
    assert r["iterations"] == int(g["iterations"]) and r["accepts"] == int(g["accepts"]) and np.isclose(r["rmse"], float(g["rmse"]), rtol=1e-6)
    fr = gb.frames()
    assert np.allclose(fr["T_eval"], g["T_eval"], atol=1e-12) and np.allclose(fr["state"], g["state"], atol=1e-12) and np.allclose(fr["frameEnergyTH"], g["frameEnergyTH"])
    assert np.allclose(gb.points()["idepth"], g["idepth"], atol=1e-9) and np.array_equal(gb.residuals()["state"], g["res_state"])
    ctx.close()


def test_window_validation_errors():
    """The flattening contract is enforced: points out of host order / residuals inconsistent with their point are refused."""
    api, synth = _mods()
    seq = cached_sequence(5, 3000, SMALL_K, SMALL_WH)
    win = synth.make_ba_window(seq, [0, 1, 2], n_per_frame=40, seed=1)
    ctx = api.Context(SMALL_K, *SMALL_WH, max_frames=4)
    for i in range(3):
        ctx.makeImages(i, seq.images[i])
    bad = dict(win); bad["host"] = win["host"][::-1].copy()
    with pytest.raises(api.SdvError):
        api.EnergyFunctional(ctx, bad, [0, 1, 2])
    bad = dict(win); bad["r_target"] = win["r_host"].copy()
    with pytest.raises(api.SdvError):
        api.EnergyFunctional(ctx, bad, [0, 1, 2])
    with pytest.raises(api.SdvError):
        api.EnergyFunctional(ctx, win, [0, 1, 99])
    ctx.close()


def test_optimize_batch_device_resident_schedule():
    """Batched mode: windows of different shapes (7, 5, 3 and 2 keyframes; with/without prior) optimised in ONE device-resident schedule
    give exactly the per-window oracle results (each window takes its own accept/reject path and stops on its own break test)."""
    api, synth = _mods()
    seq = cached_sequence(8, 2000, synth.KITTI_K, synth.KITTI_WH)
    w, h = synth.KITTI_WH; L = api.pyr_levels(w, h)
    ctx = api.Context(synth.KITTI_K, w, h, max_frames=9)
    for k in range(8):
        ctx.makeImages(k, seq.images[k])
    frames = [orc.Frame(seq.images[k], L) for k in range(8)]
    cfgs = [dict(kfs=[0, 1, 2, 3, 4, 5, 6], noise=(0.005, 0.0003), prior=1e-3), dict(kfs=[1, 2, 4, 5, 7], noise=(0.02, 0.002), prior=0.0),
            dict(kfs=[0, 3, 6], noise=(0.01, 0.0005), prior=1e-2), dict(kfs=[2, 3], noise=(0.003, 0.0002), prior=0.0),
            dict(kfs=[0, 1, 2, 3, 4, 5, 6, 7], noise=(0.0, 0.0), prior=1e-3)]
    gbs, obs = [], []
    for i, cf in enumerate(cfgs):
        win = synth.make_ba_window(seq, cf["kfs"], n_per_frame=150 + 20 * i, seed=20 + i, pose_noise=cf["noise"], match_noise=0.2, prior_scale=cf["prior"])
        gbs.append(api.EnergyFunctional(ctx, win, cf["kfs"], window=i))
        obs.append(orc.BAWindow(win, [frames[k] for k in cf["kfs"]]))
    r = api.optimize_batch(ctx, list(range(len(cfgs))), 6)
    for i, (gb, ob) in enumerate(zip(gbs, obs)):
        ro = ob.optimize(6)
        assert (ro["iterations"], ro["accepts"]) == (int(r["iterations"][i]), int(r["accepts"][i])), (i, ro, r)
        assert np.float32(ro["rmse"]) == r["rmse"][i]
        fo, fg = ob.frames(), gb.frames()
        assert _same(fo["T_eval"], fg["T_eval"]) and _same(fo["state"], fg["state"]) and _same(fo["frameEnergyTH"], fg["frameEnergyTH"])
        assert _same(ob.points()["idepth"], gb.points()["idepth"])
        so, sg = ob.residuals(), gb.residuals()
        assert _same(so["state"], sg["state"]) and _same(so["toRemove"], sg["toRemove"])
    ctx.close()
