"""GPU parity (-m gpu) of sdv_map_set / sdv_reproject_map_batch against the Reprojector oracle.
Geometry is fp64 in the oracle's operation order and the alignment sums are accumulated in pixel order -> matches are compared EXACTLY
(same points in the same order, identical aligned pixels)."""
import numpy as np
import pytest
import orc
from conftest import SMALL_K, SMALL_WH, cached_sequence

pytestmark = pytest.mark.gpu


def _mods():
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import api, synth
    return api, synth


def _scene(api, synth, K, wh, n, seed, kfs, n_per_frame, **kw):
    seq = cached_sequence(n, seed, K, wh); L = api.pyr_levels(*wh)
    pts, host_T7, host_ab = synth.make_map(seq, kfs, n_per_frame=n_per_frame, seed=2, **kw)
    ctx = api.Context(K, wh[0], wh[1], max_frames=n + 1)
    for i in range(n):
        ctx.makeImages(100 + i, seq.images[i])
    frames = [orc.Frame(seq.images[i], L) for i in range(n)]
    poses = np.array([np.concatenate([synth._quat_from_R(seq.R[i]), seq.t[i]]) for i in range(n)])
    return seq, L, pts, host_T7, host_ab, ctx, frames, poses


def _eq(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_reproject_map_exact_small():
    api, synth = _mods(); w, h = SMALL_WH; kfs = [0, 1, 2, 3]
    seq, L, pts, hT, hab, ctx, frames, poses = _scene(api, synth, SMALL_K, SMALL_WH, 5, 3000, kfs, 250)
    hab = np.array([[0.01, -0.5], [0.0, 0.3], [-0.02, 0.0], [0.015, 1.0]])
    rp = api.Reprojector(ctx); rp.setMap(0, [100 + k for k in kfs], hT, hab, pts)
    kf_frames = [frames[k] for k in kfs]
    for cur_ab, order_seed, cap in (([0.0, 0.0], None, 1200), ([0.03, -2.0], 5, 1200), ([0.0, 0.0], 7, 15)):
        order = None if order_seed is None else np.random.default_rng(order_seed).permutation(rp.n_cells).astype(np.int32)
        o = orc.reproject_map(w, h, L, SMALL_K, kf_frames, hT, hab, frames[4], poses[4], cur_ab, pts, cell_order=order, max_matches=cap)
        g = rp.reprojectMap(0, 104, poses[4], cur_ab, cell_order=order, max_matches=cap)
        assert len(o[0]) > 10 and _eq(o, g), (len(o[0]), len(g[0]))
    ctx.close()


def test_reproject_map_exact_kitti_batch():
    """KITTI-sized frames, 7 keyframes, noisy depths (pushes some matches to coarser search levels / failures), batched with backprojectMap jobs
    and the <= 2 keyframe rule in a second slot."""
    api, synth = _mods(); K, wh = synth.KITTI_K, synth.KITTI_WH; w, h = wh; kfs = [0, 1, 2, 3, 4, 5, 6]
    seq, L, pts, hT, hab, ctx, frames, poses = _scene(api, synth, K, wh, 8, 2000, kfs, 400, idepth_noise=0.02)
    rp = api.Reprojector(ctx); rp.setMap(0, [100 + k for k in kfs], hT, hab, pts)
    two = pts[pts["host"] < 2]; rp.setMap(1, [100, 101], hT[:2], hab[:2], two)
    noisy = poses[7].copy(); noisy[4:] += [0.03, -0.01, 0.05]
    jobs = [dict(slot=0, cur=107, T=poses[7], kf=-1, oh=-1, bk=0), dict(slot=0, cur=107, T=noisy, kf=-1, oh=-1, bk=0),
            dict(slot=0, cur=100, T=poses[0], kf=0, oh=6, bk=1), dict(slot=0, cur=106, T=poses[6], kf=6, oh=0, bk=1),
            dict(slot=1, cur=107, T=poses[7], kf=-1, oh=-1, bk=0), dict(slot=1, cur=100, T=poses[0], kf=0, oh=1, bk=1)]
    res = rp.reprojectMapBatch([j["slot"] for j in jobs], [j["cur"] for j in jobs], np.stack([j["T"] for j in jobs]), None,
                               [j["kf"] for j in jobs], [j["oh"] for j in jobs], [j["bk"] for j in jobs])
    kf_frames = [frames[k] for k in kfs]
    for j, g in zip(jobs, res):
        if j["slot"] == 0:
            o = orc.reproject_map(w, h, L, K, kf_frames, hT, hab, frames[j["cur"] - 100], j["T"], [0.0, 0.0], pts, cur_kf_index=j["kf"], only_host=j["oh"], backup=bool(j["bk"]))
        else:
            o = orc.reproject_map(w, h, L, K, kf_frames[:2], hT[:2], hab[:2], frames[j["cur"] - 100], j["T"], [0.0, 0.0], two, cur_kf_index=j["kf"], only_host=j["oh"], backup=bool(j["bk"]))
        assert len(o[0]) > 20 and _eq(o, g), (j, len(o[0]), len(g[0]))
    ctx.close()


def test_reproject_then_struct_pose_recovers_pose():
    """D4 end to end on the device path: matches of reprojectMap feed structPoseEstimation; the refined pose is closer to ground truth."""
    api, synth = _mods(); K, wh = synth.KITTI_K, synth.KITTI_WH; kfs = [0, 1, 2, 3, 4, 5, 6]
    seq, L, pts, hT, hab, ctx, frames, poses = _scene(api, synth, K, wh, 8, 2000, kfs, 400)
    rp = api.Reprojector(ctx); rp.setMap(0, [100 + k for k in kfs], hT, hab, pts)
    T0 = poses[7].copy(); T0[4:] += [0.04, -0.02, 0.06]
    idx, px = rp.reprojectMap(0, 107, T0)
    ov = np.zeros(len(idx), api.OVERLAP_PT_DTYPE)
    for k in ("u", "v", "idepth", "host"):
        ov[k] = pts[k][idx]
    ov["obs_x"], ov["obs_y"] = px[:, 0], px[:, 1]
    r = api.CoarseTracker(ctx, 0).structPoseEstimation(T0, ov, hT)
    e0 = np.linalg.norm(T0[4:] - poses[7][4:]); e1 = np.linalg.norm(r["T"][4:] - poses[7][4:])
    assert len(idx) > 200 and r["accepts"] >= 1 and e1 < 0.5 * e0
    ctx.close()


def test_fused_refine_equals_separate_calls_and_oracle():
    api, synth = _mods(); K, wh = synth.KITTI_K, synth.KITTI_WH; w, h = wh; kfs = [0, 1, 2, 3, 4, 5, 6]
    seq, L, pts, hT, hab, ctx, frames, poses = _scene(api, synth, K, wh, 8, 2000, kfs, 400, idepth_noise=0.01)
    rp = api.Reprojector(ctx); rp.setMap(0, [100 + k for k in kfs], hT, hab, pts); rp.setMap(1, [100 + k for k in kfs[:4]], hT[:4], hab[:4], pts[pts["host"] < 4])
    rng = np.random.default_rng(3); order = rng.permutation(rp.n_cells).astype(np.int32)
    T0 = np.stack([poses[7] + np.concatenate([np.zeros(4), rng.normal(0, 0.04, 3)]) for _ in range(4)])
    slots = [0, 1, 0, 1]
    f = rp.refineBatch(slots, [107] * 4, T0, cell_order=order, max_matches=150)
    tr = api.CoarseTracker(ctx, 0); kf_frames = [frames[k] for k in kfs]
    for k in range(4):
        P = pts if slots[k] == 0 else pts[pts["host"] < 4]; nH = 7 if slots[k] == 0 else 4
        idx, px = rp.reprojectMap(slots[k], 107, T0[k], cell_order=order, max_matches=150)
        ov = np.zeros(len(idx), api.OVERLAP_PT_DTYPE)
        for key in ("u", "v", "idepth", "host"):
            ov[key] = P[key][idx]
        ov["obs_x"], ov["obs_y"] = px[:, 0], px[:, 1]
        r = tr.structPoseEstimation(T0[k], ov, hT[:nH])
        assert f["n_matches"][k] == len(idx) == 151 and np.array_equal(f["T"][k], r["T"]) and (f["iterations"][k], f["accepts"][k]) == (r["iterations"], r["accepts"])
        oi, opx = orc.reproject_map(w, h, L, K, kf_frames[:nH], hT[:nH], hab[:nH], frames[7], T0[k], [0.0, 0.0], P, cell_order=order, max_matches=150)
        p6 = np.stack([P["u"][oi], P["v"][oi], P["idepth"][oi], P["host"][oi].astype(np.float32), opx[:, 0].astype(np.float32), opx[:, 1].astype(np.float32)], 1).astype(np.float32)
        o = orc.struct_pose(w, h, np.array(K, np.float32), hT[:nH], p6, T0[k])
        assert (o["iterations"], o["accepts"]) == (int(f["iterations"][k]), int(f["accepts"][k])) and np.abs(o["T"] - f["T"][k]).max() < 1e-9
    ctx.close()


def test_reproject_errors():
    api, synth = _mods(); kfs = [0, 1, 2]
    seq, L, pts, hT, hab, ctx, frames, poses = _scene(api, synth, SMALL_K, SMALL_WH, 5, 3000, kfs, 50)
    rp = api.Reprojector(ctx)
    with pytest.raises(api.SdvError):
        rp.reprojectMap(0, 104, poses[4])                                   # slot not set
    bad = pts.copy(); bad["host"][3] = 9; rp.setMap(0, [100, 101, 102], hT, hab, bad)
    with pytest.raises(api.SdvError):
        rp.reprojectMap(0, 104, poses[4])
    with pytest.raises(api.SdvError):
        rp.setMap(0, [100, 555, 102], hT, hab, pts)                         # unknown keyframe handle
    ctx.close()


def test_reproject_edge_cases_empty_map_and_no_candidates():
    """Empty map, a map whose points all project outside the target frame, and a target equal to a keyframe (cur_kf_index skips its own points)."""
    api, synth = _mods(); kfs = [0, 1, 2]; w, h = SMALL_WH
    seq, L, pts, hT, hab, ctx, frames, poses = _scene(api, synth, SMALL_K, SMALL_WH, 5, 3000, kfs, 120)
    rp = api.Reprojector(ctx)
    rp.setMap(0, [100, 101, 102], hT, hab, pts[:0])
    idx, px = rp.reprojectMap(0, 104, poses[4]); assert len(idx) == 0
    r = rp.refineBatch([0], [104], poses[4:5]); assert r["n_matches"][0] == 0 and np.array_equal(r["T"][0], poses[4]) and r["accepts"][0] == 0
    far = poses[4].copy(); far[4:] += [500.0, 0, 0]                                   # camera 500 m to the side: nothing lands in the image
    rp.setMap(0, [100, 101, 102], hT, hab, pts)
    o = orc.reproject_map(w, h, L, SMALL_K, [frames[k] for k in kfs], hT, hab, frames[4], far, [0.0, 0.0], pts)
    g = rp.reprojectMap(0, 104, far); assert _eq(o, g)
    o = orc.reproject_map(w, h, L, SMALL_K, [frames[k] for k in kfs], hT, hab, frames[1], poses[1], [0.0, 0.0], pts, cur_kf_index=1)
    g = rp.reprojectMap(0, 101, poses[1], cur_kf_index=1); assert _eq(o, g) and len(g[0]) > 5 and np.all(pts["host"][g[0]] != 1)
    ctx.close()
