"""Pins the oracle restatement (oracle/orc_*.cpp) on the reference's own code.

oracle/_ref/libsdvref.so is the UNMODIFIED hot-path sources of /root/reference/src compiled against stand-in headers (oracle/Makefile `ref`,
oracle/ref_stub/, oracle/ref_shim.cpp).  Everything the reference computes with its own scalar / SSE code — makeImages, makeK, makeCoarseDepthL0,
calcRes, calcGSSSE + Accumulator9, the LM loop of trackNewestCoarse — must equal the restatement BIT FOR BIT on the same inputs; this is the
anchor that lets the GPU tests compare against `orc` at full speed on the GPU box (where /root/reference does not exist).
Tracker rows here; back-end rows in test_ref_pin_ba.py.
"""
import numpy as np
import pytest
import orc
import ref
from conftest import cached_sequence, SMALL_K, SMALL_WH

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference absent")

T_ID = np.array([1, 0, 0, 0, 0, 0, 0.0])


def _frames(seq, wh, K, idx=(0, 1)):
    L = ref.set_calib(wh[0], wh[1], K); ref.settings()
    rf = [ref.Frame(seq.images[i], wh, L) for i in idx]; of = [orc.Frame(seq.images[i], L) for i in idx]
    return L, rf, of


def _cloud_inputs(seq, n, seed=0):
    from sdv_loam_b200 import synth
    pts = synth.select_points(seq.images[0], seq.clouds[0], n, seed=seed)
    rng = np.random.default_rng(seed)
    p4 = np.concatenate([pts, rng.uniform(1e-4, 1e-2, (len(pts), 1)).astype(np.float32)], 1).astype(np.float32)
    return p4


@pytest.mark.parametrize("cfg", ["small", "kitti", "k360", "stress"])
def test_pyramid_and_calibration_bit_exact(cfg):
    from sdv_loam_b200 import synth
    K, wh = {"small": (SMALL_K, SMALL_WH), "kitti": (synth.KITTI_K, synth.KITTI_WH), "k360": (synth.K360_K, synth.K360_WH), "stress": (synth.STRESS_K, synth.STRESS_WH)}[cfg]
    rng = np.random.default_rng(3); img = np.rint(rng.uniform(0, 255, (wh[1], wh[0]))).astype(np.float32)
    L = ref.set_calib(wh[0], wh[1], K)
    rf = ref.Frame(img, wh, L); of = orc.Frame(img, L)
    rt = ref.CoarseTracker(); ot = orc.CoarseTracker(wh[0], wh[1], L, K)
    for l in range(L):
        a, b = rf.dI(l), of.dI(l)
        assert np.array_equal(a[..., 0], b[..., 0])                       # intensities, every pixel
        assert np.array_equal(a[1:-1], b[1:-1])                           # gradients: the reference never writes rows 0 / h-1 (uninitialised there)
        assert np.array_equal(rf.absSquaredGrad(l)[1:-1], of.absSquaredGrad(l)[1:-1])
        assert np.array_equal(rt.K(l), ot.K(l)) and np.array_equal(rt.Ki(l), ot.Ki(l))
        assert np.array_equal(ref.global_K(l)[0], ot.K(l))


def test_reference_cloud_bit_exact_with_collisions_and_old_keyframe_points(small_seq):
    L, rf, of = _frames(small_seq, SMALL_WH, SMALL_K, (0, 1))
    p4 = _cloud_inputs(small_seq, 900)
    # colliding splats (same pixel, different depth) in both classes + old-keyframe rows (+0.5 rounding) first, as the reference walks them
    rng = np.random.default_rng(5)
    old = p4[:200].copy(); old[:, :2] += rng.uniform(-0.4, 0.4, (200, 2)).astype(np.float32); old = np.concatenate([old, old[:40] * np.float32([1, 1, 1.1, 2])])
    new = np.concatenate([p4[200:], p4[200:260] * np.float32([1, 1, 0.9, 0.5])])
    pts = np.concatenate([old, new]).astype(np.float32); rh = np.concatenate([np.ones(len(old), np.int32), np.zeros(len(new), np.int32)])
    rt = ref.CoarseTracker(); ot = orc.CoarseTracker(SMALL_WH[0], SMALL_WH[1], L, SMALL_K)
    rt.setCoarseTrackingRef(rf[0], pts, rh, old=rf[1]); ot.setCoarseTrackingRef(of[0], pts, rh)
    for l in range(L):
        a, b = rt.cloud(l), ot.cloud(l)
        assert len(a[0]) == len(b[0]) > 0
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("seed", [0, 1])
def test_calcres_and_gs_bit_exact(small_seq, seed):
    L, rf, of = _frames(small_seq, SMALL_WH, SMALL_K, (0, 1))
    p4 = _cloud_inputs(small_seq, 1200, seed); rh = np.zeros(len(p4), np.int32)
    rt = ref.CoarseTracker(); ot = orc.CoarseTracker(SMALL_WH[0], SMALL_WH[1], L, SMALL_K)
    rt.setCoarseTrackingRef(rf[0], p4, rh, 0.02, 1.5); ot.setCoarseTrackingRef(of[0], p4, rh, 0.02, 1.5)
    rng = np.random.default_rng(seed)
    poses = [T_ID, orc.se3_exp(np.array([0.02, -0.01, -1.0, 0.002, -0.004, 0.001])), orc.se3_exp(rng.normal(0, [0.3, 0.1, 0.5, 0.02, 0.03, 0.02])), orc.se3_exp(np.array([0, 0, -3.0, 0, 0.2, 0]))]
    for T in poses:
        for l in range(L):
            for cutoff, a, b in ((20.0, 0.0, 0.0), (40.0, 0.03, -2.0), (5.0, -0.1, 4.0)):
                r1 = rt.calcRes(rf[1], l, T, a, b, cutoff); r2 = ot.calcRes(of[1], l, T, a, b, cutoff)
                assert np.array_equal(r1, r2, equal_nan=True), (l, cutoff, r1, r2)
                assert np.array_equal(rt.warped(), ot.warped())                                   # the 8 buf_warped_* arrays, padded to x4
                H1, b1 = rt.calcGSSSE(l, T, a, b); H2, b2 = ot.calcGSSSE(l, T, a, b)               # Accumulator9 SSE lanes + 1k/1M tiers (MatrixAccumulators.h:937-1292)
                assert np.array_equal(H1, H2, equal_nan=True) and np.array_equal(b1, b2, equal_nan=True)


@pytest.mark.parametrize("modes", [(0.0, 0.0), (-1.0, -1.0), (-1.0, 0.0), (0.0, -1.0), (1e5, 1e8)])
def test_track_newest_coarse_bit_exact(small_seq, modes):
    """The whole coarse-to-fine LM: identical iterates, hence identical final pose / affine / residuals / flow — for every affine mode of CoarseTracker.cpp:726-748."""
    L, rf, of = _frames(small_seq, SMALL_WH, SMALL_K, (0, 1, 2))
    ref.settings(6.0, 20.0, *modes)
    p4 = _cloud_inputs(small_seq, 1500); rh = np.zeros(len(p4), np.int32)
    rt = ref.CoarseTracker(); ot = orc.CoarseTracker(SMALL_WH[0], SMALL_WH[1], L, SMALL_K); ot.settings(6.0, 20.0, *modes)
    rt.setCoarseTrackingRef(rf[0], p4, rh); ot.setCoarseTrackingRef(of[0], p4, rh)
    rng = np.random.default_rng(11)
    inits = [T_ID, orc.se3_exp(np.array([0, 0, -0.9, 0, 0, 0.0])), orc.se3_exp(rng.normal(0, [0.2, 0.1, 0.6, 0.01, 0.02, 0.01])), orc.se3_exp(np.array([0.5, 0, 2.0, 0, 0.1, 0]))]
    for k in (1, 2):
        for T0 in inits:
            for minRes in (None, np.array([1.0, 1.0, 1.0, 1.0, np.nan])):
                a = rt.trackNewestCoarse(rf[k], T0, [0.0, 0.0], L - 1, minRes); b = ot.trackNewestCoarse(of[k], T0, [0.0, 0.0], L - 1, minRes)
                assert a["good"] == b["good"]
                assert np.array_equal(a["T"], b["T"]) and np.array_equal(a["ab"], b["ab"])
                assert np.array_equal(a["lastResiduals"], b["lastResiduals"], equal_nan=True) and np.array_equal(a["flow"], b["flow"], equal_nan=True)
    ref.settings()


def test_interpolation_and_afflight_bit_exact():
    rng = np.random.default_rng(0); img = rng.uniform(0, 255, (40, 64, 3)).astype(np.float32)
    L = 1; of = None
    for _ in range(200):
        x, y = float(rng.uniform(2, 60)), float(rng.uniform(2, 37))
        r = ref.interp33(img, np.float32(x), np.float32(y))
        ix, iy = int(np.float32(x)), int(np.float32(y)); dx = np.float32(x) - np.float32(ix); dy = np.float32(y) - np.float32(iy); dxdy = np.float32(dx * dy)
        e = (dxdy * img[iy + 1, ix + 1] + np.float32(dy - dxdy) * img[iy + 1, ix]) + np.float32(dx - dxdy) * img[iy, ix + 1] + np.float32(np.float32(np.float32(1) - dx) - dy + dxdy) * img[iy, ix]
        assert np.allclose(r, e, rtol=1e-6)
    for eF, eT, aF, bF, aT, bT in [(1, 1, 0, 0, 0, 0), (0.5, 2.0, 0.1, 3.0, -0.2, 7.0), (0, 1, 0.3, 1, 0.1, 2)]:
        o = np.zeros(2); orc.lib().orc_aff_from_to(eF, eT, aF, bF, aT, bT, o)
        assert np.array_equal(ref.aff_from_to(eF, eT, aF, bF, aT, bT), o)
