"""GPU parity (-m gpu) of sdv_track_new_coarse_batch (FullSystem::trackNewCoarse as a batched policy) against the oracle
restatement orc.track_new_coarse: same number of tries, same winner, final pose within the tracker's parity bound (1e-6 << 1e-3 m / rad)."""
import numpy as np
import pytest
import orc
from conftest import cached_sequence

pytestmark = pytest.mark.gpu


def _mods():
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import api, synth
    return api, synth


def test_track_new_coarse_batch_matches_oracle():
    api, synth = _mods(); K, wh = synth.KITTI_K, synth.KITTI_WH; w, h = wh; L = api.pyr_levels(w, h)
    seq = cached_sequence(8, 2000, K, wh); kfs = [0, 1, 2, 3]
    pts, hT, hab = synth.make_map(seq, kfs, n_per_frame=400, seed=2)
    poses = np.array([np.concatenate([synth._quat_from_R(seq.R[i]), seq.t[i]]) for i in range(8)])
    ctx = api.Context(K, w, h, max_frames=10, n_tracker_slots=4)
    for i in range(8):
        ctx.makeImages(100 + i, seq.images[i])
    frames = [orc.Frame(seq.images[i], L) for i in range(8)]
    ref_pts = synth.select_points(seq.images[3], seq.clouds[3], 2000); p4 = np.concatenate([ref_pts, np.full((len(ref_pts), 1), 1e-3, np.float32)], 1).astype(np.float32)
    rh = np.zeros(len(p4), np.int32)
    rp = api.Reprojector(ctx); otr = orc.CoarseTracker(w, h, L, K); otr.setCoarseTrackingRef(frames[3], p4, rh)
    for s in range(4):
        api.CoarseTracker(ctx, s).setCoarseTrackingRef(103, p4, rh); rp.setMap(s, [100 + k for k in kfs], hT, hab, pts)
    order = np.random.default_rng(4).permutation(rp.n_cells).astype(np.int32)
    base = dict(frame=106, sprelast_c2w=poses[4], slast_c2w=poses[5], lastF_c2w=poses[3], aff_last=[0.0, 0.0])
    jobs = [dict(base, slot=0, poses_valid=1, lastCoarseRMSE=[100.0] * 5),                       # healthy: one try
            dict(base, slot=1, poses_valid=1, lastCoarseRMSE=[1e-3] * 5),                        # immediate-accept rule never met: all 31 hypotheses
            dict(base, slot=2, poses_valid=0, lastCoarseRMSE=[100.0] * 5, frame=104),            # invalid history: identity hypothesis only
            dict(base, slot=3, poses_valid=2, lastCoarseRMSE=[1e-3] * 5, frame=104)]             # second frame of a sequence: identity + 52 pure rotations, all tried
    res = api.trackNewCoarseBatch(ctx, jobs, cell_order=order)
    kf_frames = [frames[k] for k in kfs]
    for j, g in zip(jobs, res):
        o = orc.track_new_coarse(otr, frames[j["frame"] - 100], K, kf_frames, hT, hab, pts, j["sprelast_c2w"], j["slast_c2w"], j["lastF_c2w"], j["aff_last"], j["poses_valid"],
                                 j["lastCoarseRMSE"], cell_order=order)
        assert g["tries"] == o["tries"] and g["have_one_good"] == o["have_one_good"], (g["tries"], o["tries"])
        assert np.allclose(g["lastCoarseRMSE"], o["lastCoarseRMSE"], rtol=1e-4, equal_nan=True) and np.allclose(g["aff_g2l"], o["aff_g2l"], atol=2e-4)       # same bound as the LM parity tests (b is in grey levels)
        assert g["n_matches"] == o["n_matches"] and (g["refine_iterations"], g["refine_accepts"]) == (o["refine_iterations"], o["refine_accepts"])
        assert np.abs(g["camToWorld"] - o["camToWorld"]).max() < 1e-6 and np.abs(g["camToTrackingRef"] - o["camToTrackingRef"]).max() < 1e-6
    assert res[0]["tries"] == 1 and res[1]["tries"] == 31 and res[2]["tries"] == 1 and res[3]["tries"] == 53
    gt = poses[6]; assert np.linalg.norm(res[0]["camToWorld"][4:] - gt[4:]) < 0.05            # metres, against the synthetic ground truth
    ctx.close()
