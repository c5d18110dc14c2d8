"""Sequence-level parity of the immature-point tracing on the reference's OWN running pipeline: the reference's FullSystem (oracle/_ref) runs the
synthetic KITTI drive; around every frame that does NOT become a keyframe (makeNonKeyFrame -> traceNewCoarse -> ImmaturePoint::traceOn for every candidate of every keyframe of
the window) all candidates are read back before and after, and the restatement, started from the "before" records with the geometry the reference forms (K R K^-1, K t of
host -> new frame from the tracked pose), must reproduce the "after" records BIT FOR BIT — every field, every status — for thousands of candidates in whatever state the
running system left them (fresh, converged, skipped, out of bounds, outliers)."""
import numpy as np
import pytest
import orc
import ref
from conftest import cached_sequence
import seq_replay as sr

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference absent")
N_FRAMES = 60


def test_oracle_follows_reference_trace_new_coarse():
    from sdv_loam_b200 import synth
    seq = cached_sequence(200, 1000, synth.KITTI_K, synth.KITTI_WH, step=0.5); w, h = seq.wh
    import ctypes as C
    libc = C.CDLL(None); libc.mallopt(-6, 0xFF)      # M_PERTURB: malloc'ed memory reads as zero — the reference reads uninitialised heap in a few places, which makes a
    try:                                             # second pipeline run in one process depend on what the first one left on the heap (the bootstrap can fail)
        _run(seq, synth, w, h)
    finally:
        libc.mallopt(-6, 0)


def _run(seq, synth, w, h):
    L = ref.set_calib(w, h, seq.K); run = sr.ReferenceRun(seq, L); S = run.S
    frames = 0; cands = 0; seen = set(); aff = np.array([1.0, 0.0], np.float32)      # perfect-image mode: exposures 1, affine parameters fixed at 0 -> fromToVecExposure = (1, 0)
    lrud = np.array([10000, -1, 10000, -1], np.int32)
    for i in range(N_FRAMES):
        cloud = sr.frame_cloud(seq, i); ku, kv = cloud[:, 0].astype(np.float32), cloud[:, 1].astype(np.float32)     # what lidarCloudHandler leaves in the FullSystem (main.cpp:834-854):
        lrud = np.array([min(lrud[0], int(ku.min())), max(lrud[1], int(ku.max())), min(lrud[2], int(kv.min())), max(lrud[3], int(kv.max()))], np.int32)
        S.set_lidar_state(lrud, 1)                                               # the running pixel box and addFeaturePoint (uninitialised members otherwise) -> monocular candidates exist
        pre = S.immature_dump() if i >= 3 else None
        nkf = S.num_keyframes(); _, _, res = run.step(); assert res["rc"] == 0, i
        if pre is None or S.num_keyframes() != nkf: continue                        # bootstrap, or the frame became a keyframe (candidates get activated / deleted there)
        post = S.immature_dump(); assert [p[0] for p in pre] == [p[0] for p in post] and [len(p[1]) for p in pre] == [len(p[1]) for p in post], i
        fo = orc.Frame(seq.images[i], L)
        for hidx, ((kid, rec0, st0), (_, rec1, st1)) in enumerate(zip(pre, post)):
            if len(rec0) == 0: continue
            KRKi, Kt = S.trace_geometry(hidx, res["camToWorld"])
            P = np.zeros(len(rec0), orc.IMM_DTYPE); P.view(np.float32).reshape(len(rec0), -1)[:, :29] = rec0; P["lastTraceStatus"] = st0
            so = orc.immature_trace(fo, P, KRKi, Kt, aff)
            got = P.view(np.float32).reshape(len(rec0), -1)[:, :29]
            assert np.array_equal(so, st1) and np.array_equal(P["lastTraceStatus"], st1), (i, kid, np.nonzero(so != st1)[0][:5])
            assert np.array_equal(got.view(np.uint32), rec1.view(np.uint32)), (i, kid, np.argwhere(got.view(np.uint32) != rec1.view(np.uint32))[:5])
            cands += len(rec0); seen |= set(int(s) for s in st1)
        frames += 1
    assert frames >= 15 and cands > 10000 and {orc.IPS_GOOD, orc.IPS_OOB, orc.IPS_SKIPPED} <= seen, (frames, cands, seen)
    print(f"traceNewCoarse: {frames} non-keyframe frames of the reference run, {cands} candidate traces identical; statuses seen {sorted(seen)}")
