"""Sequence-level parity of CoarseDistanceMap and the activation walk (caller half of SURVEY §8f rank 2) on the LIVE windows of the reference's own running pipeline: while the
reference's FullSystem (oracle/_ref) runs the synthetic KITTI drive, every few frames the reference's own makeDistanceMap is run on its current window (sources = the ACTIVE
points of the other keyframes, forward-warped with the window's real poses and the bundle-adjusted intrinsics) and compared with the restatement fed from the same state
(points read back from the window, K R K^-1 / K t as the reference forms them); then the same random candidates walk both maps (the reference side on its own BFS).
Distance maps and decisions: BIT FOR BIT."""
import numpy as np
import pytest
import orc
import ref
from conftest import cached_sequence
import seq_replay as sr

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference absent")
N_FRAMES = 60


def test_oracle_follows_reference_distance_maps():
    from sdv_loam_b200 import synth
    import ctypes as C
    seq = cached_sequence(200, 1000, synth.KITTI_K, synth.KITTI_WH, step=0.5)
    libc = C.CDLL(None); libc.mallopt(-6, 0xFF)      # M_PERTURB: zero heap, so that this pipeline run does not depend on what earlier runs in the process left behind (see test_sequence_trace.py)
    try:
        _run(seq)
    finally:
        libc.mallopt(-6, 0)


def _run(seq):
    w, h = seq.wh
    L = ref.set_calib(w, h, seq.K); run = sr.ReferenceRun(seq, L); S = run.S; rng = np.random.default_rng(11)
    checked = 0; sources = 0; accepted = 0; lrud = np.array([10000, -1, 10000, -1], np.int32)
    for i in range(N_FRAMES):
        cloud = sr.frame_cloud(seq, i); ku, kv = cloud[:, 0].astype(np.float32), cloud[:, 1].astype(np.float32)
        lrud = np.array([min(lrud[0], int(ku.min())), max(lrud[1], int(ku.max())), min(lrud[2], int(kv.min())), max(lrud[3], int(kv.max()))], np.int32)
        S.set_lidar_state(lrud, 1)                                               # the members lidarCloudHandler would have set (main.cpp:834-854)
        _, _, res = run.step(); assert res["rc"] == 0, i
        if i < 6 or i % 3: continue
        snap = S.tracker_snapshot(); nF = len(snap["kf_ids"]); p5 = snap["map_pts"]
        if nF < 3 or len(p5) < 50: continue
        newest = nF - 1; rd = ref.DistMap(S, newest, (w, h)); rd.make()                     # RefSys and RefBA both start with their FullSystem*: the distance-map entries of the shim take either
        hosts = [k for k in range(nF) if k != newest]; geo = [rd.geometry(k) for k in hosts]
        uvid = []; pb = [0]
        for k in hosts:
            m = p5[:, 3].astype(np.int32) == k; uvid.append(p5[m, :3]); pb.append(pb[-1] + int(m.sum()))
        od = orc.DistMap(w >> 1, h >> 1); od.make(pb, np.stack([g[0] for g in geo]), np.stack([g[1] for g in geo]), np.concatenate(uvid).astype(np.float32))
        a, b = od.get(), rd.get(); assert np.array_equal(a, b), (i, int((a != b).sum()))
        # candidates on every keyframe of the window (the newest included), judged by the greedy walk against both maps
        cb = [0]; cand = []
        for k in range(nF):
            n = 300; cand.append(np.stack([rng.integers(4, w - 5, n), rng.integers(4, h - 5, n), rng.uniform(0.02, 0.5, n), rng.choice([1.0, 2.0, 4.0], n)], 1).astype(np.float32)); cb.append(cb[-1] + n)
        cand = np.concatenate(cand); geo_all = [rd.geometry(k) for k in range(nF)]; minDist = float(rng.choice([0.5, 1.5, 3.0]))
        do = od.activateSelect(cb, np.stack([g[0] for g in geo_all]), np.stack([g[1] for g in geo_all]), cand, minDist)
        dr = rd.activateSelect(list(range(nF)), cb, cand, minDist)
        assert np.array_equal(do, dr), (i, minDist) ; assert np.array_equal(od.get(), rd.get()), i
        checked += 1; sources += pb[-1]; accepted += int((do == 1).sum())
    assert checked >= 10 and sources > 5000 and accepted > 500, (checked, sources, accepted)
    print(f"CoarseDistanceMap: {checked} live windows of the reference run, {sources} source points, {accepted} accepted candidates, maps and decisions identical")
