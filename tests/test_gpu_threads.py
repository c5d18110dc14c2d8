"""One context, two host threads — the reference's threading model: the tracking thread calls the tracker-slot / frame / policy entries (FullSystem::trackMutex)
while the mapping thread calls the back-end entries (mapMutex).  The context runs the two domains on separate streams under separate locks; results must be
bit-identical to the same calls made one after the other."""
import threading
import numpy as np
import pytest
from conftest import cached_sequence

pytestmark = pytest.mark.gpu
ID7 = np.array([1, 0, 0, 0, 0, 0, 0.0])


def test_tracking_and_mapping_threads_share_one_context():
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import api, synth
    K, wh = synth.KITTI_K, synth.KITTI_WH; w, h = wh; B = 6
    seq = cached_sequence(8, 2000, K, wh)
    ctx = api.Context(K, w, h, n_tracker_slots=B, max_frames=3 * B + 16)
    pts = synth.select_points(seq.images[0], seq.clouds[0], 1500); p4 = np.concatenate([pts, np.full((len(pts), 1), 1e-3, np.float32)], 1).astype(np.float32)
    for b in range(B):
        ctx.makeImages(1000 + b, seq.images[0]); api.CoarseTracker(ctx, b).setCoarseTrackingRef(1000 + b, p4, np.zeros(len(p4), np.int32))
    kfs = list(range(7)); win = synth.make_ba_window(seq, kfs, n_per_frame=250, seed=3, pose_noise=(0.005, 0.0003), match_noise=0.1, prior_scale=1e-3)
    kf_ids = [2000 + k for k in kfs]
    for k in kfs:
        ctx.makeImages(2000 + k, seq.images[k])
    u8 = [np.ascontiguousarray(seq.images[1 + (i % 3)].astype(np.uint8)) for i in range(3)]
    slots = np.arange(B, dtype=np.int32); N_TRACK, N_BA = 12, 6

    def tracking(out):
        for i in range(N_TRACK):                                               # upload (frame table + ingest streams) and track, like the per-frame path
            ids = np.arange(B, dtype=np.uint64) + 100 * (i & 1)
            ctx.makeImagesBatch(ids, [u8[i % 3].ctypes.data] * B, u8=True)
            T = np.tile(ID7, (B, 1)); T[:, 6] = -0.9; ab = np.zeros((B, 2))
            r = ctx.trackBatch(slots, ids, T, ab)
            out.append((T.copy(), r["lastResiduals"].copy(), r["iterations"].copy()))

    def mapping(out):
        for i in range(N_BA):                                                  # set_window touches the frame table (pins, level-0 texels) from the mapping thread
            ef = api.EnergyFunctional(ctx, win, kf_ids, window=0)
            r = ef.optimize(6)
            out.append((r["rmse"], r["iterations"], r["accepts"], ef.frames()["state"].copy(), ef.points()["idepth"].copy()))

    serial_t, serial_m = [], []; tracking(serial_t); mapping(serial_m)
    par_t, par_m = [], []; errs = []

    def guard(fn, out):
        try:
            fn(out)
        except Exception as e:                                                 # noqa: BLE001 — reported below
            errs.append(e)
    ta = threading.Thread(target=guard, args=(tracking, par_t)); tb = threading.Thread(target=guard, args=(mapping, par_m))
    ta.start(); tb.start(); ta.join(); tb.join()
    assert not errs, errs
    assert len(par_t) == N_TRACK and len(par_m) == N_BA
    for a, b in zip(serial_t, par_t):
        assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))
    for a, b in zip(serial_m, par_m):
        assert a[:3] == b[:3] and np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
    # a frame referenced by the resident window cannot be released from the tracking thread meanwhile
    with pytest.raises(api.SdvError):
        ctx.releaseFrame(2000)
    ctx.close()
