"""Pins the immature-point restatement (oracle/orc_trace.cpp: ImmaturePoint constructor + traceOn) on the reference's own compiled
ImmaturePoint.cpp (oracle/_ref): every field of every candidate after construction and after one / two / three consecutive traces must agree BIT FOR BIT,
over candidates that end in every status of the machine (GOOD, OOB, OUTLIER, SKIPPED, BADCONDITION)."""
import numpy as np
import pytest
import orc
import ref
from conftest import cached_sequence, SMALL_K, SMALL_WH

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference absent")
FLOAT_FIELDS = ["u", "v", "idepth_min", "idepth_max", "color", "weights", "gradH", "energyTH", "quality", "lastTraceUV", "lastTracePixelInterval"]


def flat(P):
    return np.concatenate([np.atleast_1d(P[f]).astype(np.float32).reshape(-1) for f in FLOAT_FIELDS])


def candidates(seq, n, seed):
    """integer pixels on image gradients of frame 0 (what PixelSelector hands to makeNewTraces), a few near the border"""
    rng = np.random.default_rng(seed); w, h = seq.wh; img = seq.images[0]
    g = np.abs(np.gradient(img)[0]) + np.abs(np.gradient(img)[1]); ys, xs = np.nonzero(g[8:h - 8, 8:w - 8] > 12)
    pick = rng.choice(len(xs), n, replace=False); uv = np.stack([xs[pick] + 8, ys[pick] + 8], 1).astype(np.int32)
    uv[:6] = [[5, 5], [w - 7, h - 7], [6, h // 2], [w // 2, 5], [w - 6, 9], [9, h - 6]]
    return uv


@pytest.mark.parametrize("wh,K,seed", [(SMALL_WH, SMALL_K, 3000), ((1200, 360), None, 2000)])
def test_constructor_and_trace_bit_exact(wh, K, seed):
    from sdv_loam_b200 import synth
    K = K or synth.KITTI_K
    seq = cached_sequence(5 if wh == SMALL_WH else 8, seed, K, wh); L = ref.set_calib(wh[0], wh[1], K)
    of = [orc.Frame(im, L) for im in seq.images[:4]]; rf = [ref.Frame(im, wh, L) for im in seq.images[:4]]
    uv = candidates(seq, 300, seed)
    P = orc.immature_init(of[0], uv); R = [ref.ImmaturePoint(rf[0], u, v) for u, v in uv]
    for i, r in enumerate(R):
        o, st = r.record()
        assert st == P["lastTraceStatus"][i] == orc.IPS_UNINITIALIZED and np.array_equal(flat(P[i])[:25], o[:25], equal_nan=True), i       # ctor: everything but the fields it leaves untouched
    poses = [np.concatenate([synth._quat_from_R(seq.R[i]), seq.t[i]]) for i in range(4)]
    seen = set()
    for k, (ab, ex) in zip((1, 2, 3), (((0.0, 0.0), 1.0), ((0.02, -1.5), 1.1), ((-0.01, 2.0), 0.9))):     # three consecutive frames, as traceNewCoarse does
        KRKi, Kt, aff = orc.trace_geometry(K, poses[0], poses[k], 1.0, ex, (0.0, 0.0), ab)
        if k == 2:                                                                                   # exercise the finite-range branches with intervals of different widths
            for i, r in enumerate(R):
                if P["lastTraceStatus"][i] == orc.IPS_GOOD and i % 3 == 0:
                    mid = 0.5 * (P["idepth_min"][i] + P["idepth_max"][i]); half = (0.02 if i % 2 else 1e-4) * max(mid, 1e-3)
                    P["idepth_min"][i] = np.float32(mid - half); P["idepth_max"][i] = np.float32(mid + half)
                    r.set_range(P["idepth_min"][i], P["idepth_max"][i], orc.IPS_GOOD)
        so = orc.immature_trace(of[k], P, KRKi, Kt, aff)
        for i, r in enumerate(R):
            sr = r.traceOn(rf[k], KRKi, Kt, aff); o, st = r.record()
            assert sr == so[i] == st == P["lastTraceStatus"][i], (k, i, sr, so[i])
            assert np.array_equal(flat(P[i]), o, equal_nan=True), (k, i, sr, flat(P[i]) - o)
        seen |= set(int(s) for s in so)
    assert {orc.IPS_GOOD, orc.IPS_OOB, orc.IPS_OUTLIER, orc.IPS_SKIPPED} <= seen, seen
    good = P[P["lastTraceStatus"] == orc.IPS_GOOD]
    assert len(good) > 50 and np.all(good["idepth_max"] >= good["idepth_min"])


def test_optimize_immature_point_bit_exact():
    """FullSystem::optimizeImmaturePoint (+ ImmaturePoint::linearizeResidual) on a 5-keyframe window: status, activated inverse depth and the final state of every temporary
    residual identical to the reference, for candidates whose depth interval comes from tracing (so the well / badly constrained, outlier and sensor branches all occur)."""
    from sdv_loam_b200 import synth
    from test_ref_pin_ba import _window
    win, ob, rb, (of, rf) = _window((0, 1, 2, 3, 4), 5)
    seq = cached_sequence(5, 3000, SMALL_K, SMALL_WH); nF = 5
    for host in (0, 2):
        pre, cal = rb.immature_pre(host, nF)
        uv = candidates(seq, 250, 70 + host) if host == 0 else None
        if uv is None:                                                                               # candidates on THIS host's gradients
            seq_h = type(seq).__new__(type(seq)); seq_h.wh = seq.wh; seq_h.images = [seq.images[host]]; uv = candidates(seq_h, 250, 70 + host)
        P = orc.immature_init(of[host], uv)
        # depth intervals: trace into a neighbouring keyframe with the true geometry (gives GOOD candidates with a tight range), widen / corrupt some
        poses = [np.concatenate([synth._quat_from_R(seq.R[i]), seq.t[i]]) for i in range(nF)]
        tgt = 1 if host == 0 else 3
        KRKi, Kt, aff = orc.trace_geometry(SMALL_K, poses[host], poses[tgt])
        orc.immature_trace(of[tgt], P, KRKi, Kt, aff)
        rng = np.random.default_rng(host); sensor = (rng.uniform(size=len(P)) < 0.15)
        bad = ~np.isfinite(P["idepth_max"]); P["idepth_max"][bad] = 0.2; P["idepth_min"][bad] = 0.0    # never-traced candidates get an arbitrary interval
        wide = rng.uniform(size=len(P)) < 0.2; P["idepth_min"][wide] *= np.float32(0.5); P["idepth_max"][wide] *= np.float32(1.7)
        targets = [of[t] for t in range(nF) if t != host]
        for min_obs in (1, 3):
            so, io, ro = orc.immature_optimize(P, sensor, targets, pre, cal, min_obs)
            for i in range(len(P)):
                sr, ir, rr = rb.optimizeImmaturePoint(host, uv[i][0], uv[i][1], float(P["idepth_min"][i]), float(P["idepth_max"][i]), bool(sensor[i]), min_obs, nF)
                assert sr == so[i], (host, i, sr, so[i])
                if sr == 1:
                    assert np.float32(ir) == io[i] and np.array_equal(rr, ro[i]), (host, i, ir, io[i], rr, ro[i])
            assert {0, 1} <= set(int(s) for s in so) or {-1, 1} <= set(int(s) for s in so)
    assert (so == 1).sum() > 30
