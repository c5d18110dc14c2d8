"""CPU tests of the Reprojector oracle (oracle/orc_reproject.cpp): matches land on the ground-truth projections of the
synthetic world, one match per grid cell, cell order respected, both alignment branches exercised."""
import numpy as np
import pytest
import orc
from conftest import SMALL_K, SMALL_WH, cached_sequence


@pytest.fixture(scope="module")
def scene():
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import synth
    seq = cached_sequence(5, 3000, SMALL_K, SMALL_WH)
    kfs = [0, 1, 2, 3]; w, h = SMALL_WH; L = 4
    pts, host_T7, host_ab = synth.make_map(seq, kfs, n_per_frame=250, seed=2)
    frames = [orc.Frame(seq.images[k], L) for k in kfs]; cur = orc.Frame(seq.images[4], L)
    cur_T7 = np.concatenate([synth._quat_from_R(seq.R[4]), seq.t[4]])
    return synth, seq, kfs, pts, host_T7, host_ab, frames, cur, cur_T7


def gt_projection(synth, seq, kfs, pts, cur_idx):
    fx, fy, cx, cy = seq.K; out = np.zeros((len(pts), 2))
    for i, p in enumerate(pts):
        k = kfs[p["host"]]; X = seq.R[k] @ (np.array([(p["u"] - cx) / fx, (p["v"] - cy) / fy, 1.0]) / p["idepth"]) + seq.t[k]
        Xc = seq.R[cur_idx].T @ (X - seq.t[cur_idx]); out[i] = [fx * Xc[0] / Xc[2] + cx, fy * Xc[1] / Xc[2] + cy]
    return out


def test_reproject_map_matches_are_near_ground_truth(scene):
    synth, seq, kfs, pts, host_T7, host_ab, frames, cur, cur_T7 = scene
    w, h = SMALL_WH
    idx, px = orc.reproject_map(w, h, 4, SMALL_K, frames, host_T7, host_ab, cur, cur_T7, [0.0, 0.0], pts)
    assert len(idx) > 60
    gt = gt_projection(synth, seq, kfs, pts, 4)[idx]
    err = np.linalg.norm(px - gt, axis=1)
    assert np.median(err) < 0.35 and np.mean(err < 1.0) > 0.75           # sub-pixel alignment; weakest-gradient-first candidates and 1-D edgelet alignment leave a tail
    cells = (px[:, 1] // 25).astype(int) * int(np.ceil(w / 25)) + (px[:, 0] // 25).astype(int)
    cand_cells = (gt[:, 1] // 25).astype(int) * int(np.ceil(w / 25)) + (gt[:, 0] // 25).astype(int)
    assert len(np.unique(cand_cells)) == len(cand_cells)                 # one match per candidate cell (cell of the projected, un-aligned pixel)
    assert np.all(np.diff(cand_cells) > 0)                               # identity cell order
    assert {0, 1} <= set(pts["type"][idx].tolist())                      # align2D and align1D both produced matches
    del cells


def test_cell_order_and_cap(scene):
    synth, seq, kfs, pts, host_T7, host_ab, frames, cur, cur_T7 = scene
    w, h = SMALL_WH; ncells = int(np.ceil(w / 25)) * int(np.ceil(h / 25))
    a_idx, a_px = orc.reproject_map(w, h, 4, SMALL_K, frames, host_T7, host_ab, cur, cur_T7, [0.0, 0.0], pts)
    order = np.random.default_rng(1).permutation(ncells).astype(np.int32)
    b_idx, b_px = orc.reproject_map(w, h, 4, SMALL_K, frames, host_T7, host_ab, cur, cur_T7, [0.0, 0.0], pts, cell_order=order)
    assert sorted(a_idx.tolist()) == sorted(b_idx.tolist())             # same matches, different order
    ia = np.argsort(a_idx); ib = np.argsort(b_idx); assert np.array_equal(a_px[ia], b_px[ib])
    c_idx, _ = orc.reproject_map(w, h, 4, SMALL_K, frames, host_T7, host_ab, cur, cur_T7, [0.0, 0.0], pts, max_matches=20)
    assert len(c_idx) == 21 and np.array_equal(c_idx, a_idx[:21])        # `n_matches_ > cap` breaks after cap+1 matches (:151)


def test_backproject_single_host_and_two_keyframe_rule(scene):
    synth, seq, kfs, pts, host_T7, host_ab, frames, cur, cur_T7 = scene
    w, h = SMALL_WH
    # backprojectMap(ref_frame = KF 0, frame = KF 2): only KF 2's points, projected into KF 0
    idx, px = orc.reproject_map(w, h, 4, SMALL_K, frames, host_T7, host_ab, frames[0], host_T7[0], [0.0, 0.0], pts, cur_kf_index=0, only_host=2, backup=True)
    assert len(idx) > 20 and np.all(pts["host"][idx] == 2)
    gt = gt_projection(synth, seq, kfs, pts, kfs[0])[idx]
    assert np.median(np.linalg.norm(px - gt, axis=1)) < 0.5
    # <= 2 keyframes: the reference patch always comes from frameHessians_[0], whatever the host (Reprojector.cpp:242-250)
    two = pts[pts["host"] < 2]
    idx2, px2 = orc.reproject_map(w, h, 4, SMALL_K, frames[:2], host_T7[:2], host_ab[:2], cur, cur_T7, [0.0, 0.0], two)
    assert len(idx2) > 20


def test_alignment_recovers_known_subpixel_shift():
    """findMatchDirect on analytic images: cur(x, y) = ref(x - dx, y - dy), identical camera poses (affine warp = identity, search level 0).
    align2D must move the projected pixel by (dx, dy); align1D (EDGELET) along the reference gradient direction."""
    import sdv_loam_b200  # noqa
    w, h = 320, 192; L = 3; K = (300.0, 300.0, 159.5, 95.5)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    tex = lambda x, y: 120 + 50 * np.sin(x / 5.3) * np.cos(y / 6.1) + 30 * np.sin((x + 2 * y) / 9.7)
    T = np.array([1, 0, 0, 0, 0, 0, 0.0])
    for (dx, dy) in ((0.37, -0.22), (-0.6, 0.45)):
        ref = orc.Frame(tex(xx, yy).astype(np.float32), L); cur = orc.Frame(tex(xx - dx, yy - dy).astype(np.float32), L)
        pts = np.zeros(40, [("u", np.float32), ("v", np.float32), ("idepth", np.float32), ("host", np.int32), ("type", np.int32)])
        rng = np.random.default_rng(3)
        pts["u"] = rng.integers(2, 10, 40) + 30 * (np.arange(40) % 10); pts["v"] = rng.integers(2, 10, 40) + 40 * (np.arange(40) // 10) + 20
        pts["idepth"] = 0.1; pts["type"] = 0
        idx, px = orc.reproject_map(w, h, L, K, [ref, ref, ref], np.stack([T] * 3), np.zeros((3, 2)), cur, T, [0.0, 0.0], pts)   # 3 keyframes: reference patch = host
        assert len(idx) >= 30
        err = px - np.stack([pts["u"][idx] + dx, pts["v"][idx] + dy], 1)
        assert np.abs(err).max() < 0.1 and np.median(np.abs(err)) < 0.03, np.abs(err).max()   # uint8 patch truncation + 0.03 px convergence threshold
    # edgelets on a high-contrast image that varies along x only: 1-D alignment recovers dx and leaves y untouched
    edge = lambda x: 120 + 90 * np.sin(x / 4.0)
    dx = 0.41; ref = orc.Frame(edge(xx).astype(np.float32), L); cur = orc.Frame(edge(xx - dx).astype(np.float32), L)
    pts["type"] = 1
    idx, px = orc.reproject_map(w, h, L, K, [ref, ref, ref], np.stack([T] * 3), np.zeros((3, 2)), cur, T, [0.0, 0.0], pts)
    gx = np.abs(edge(pts["u"][idx] + 0.5) - edge(pts["u"][idx] - 0.5)); strong = gx > 8.0         # an edgelet needs an edge (uint8 patches: error ~ 0.5 grey / gradient)
    assert len(idx) >= 25 and strong.sum() >= 10 and np.abs(px[strong, 0] - (pts["u"][idx][strong] + dx)).max() < 0.15 and np.median(np.abs(px[strong, 0] - (pts["u"][idx][strong] + dx))) < 0.05 and np.abs(px[:, 1] - pts["v"][idx]).max() < 1e-6
