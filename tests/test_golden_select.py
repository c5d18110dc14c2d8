"""Golden vectors of the keyframe-rate candidate management and the LiDAR front-end, produced by the REFERENCE ITSELF (tests/golden/make_select_golden.py runs the reference's own
compiled PixelSelector2.cpp / FullSystem.cpp / main.cpp from oracle/_ref): thresholds, selection maps at three potentials (LiDAR and monocular), two consecutive makeNewTraces
calls (the second one walking the stale monocular map), one decimated sweep through lidarCloudHandler.  The oracle (CPU) and the CUDA path (-m gpu) must reproduce them bit for
bit — also on a box that has neither /root/reference nor oracle/_ref."""
import os
import numpy as np
import pytest
import orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "select_small.npz")
TRACK = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tracker_small.npz")
W, H = 640, 192
CASES = ((3, 400.0), (1, 5000.0), (6, 50.0))


def _inputs():
    g = np.load(GOLD); t = np.load(TRACK)
    return g, t["img0"].astype(np.float32), t["img1"].astype(np.float32), tuple(float(k) for k in g["K"])


def _rows(T):
    return np.stack([T["u"], T["v"], T["my_type"], T["score"], T["idepth_fromSensor"], T["isFromSensor"].astype(np.float32), T["type"].astype(np.float32)], 1)


def test_oracle_reproduces_reference_goldens():
    g, img0, img1, K = _inputs(); L = orc.lib().orc_pyr_levels(W, H); f = [orc.Frame(img0, L), orc.Frame(img1, L)]; rp = orc.libc_random_pattern(W, H)
    assert int(np.bitwise_xor.reduce(rp.astype(np.uint32) * (np.arange(W * H, dtype=np.uint32) | 1))) == int(g["random_pattern_crc"][0])
    s = orc.Selector(W, H, rp); a, b = s.makeHists(f[0]); assert np.array_equal(a, g["ths"]) and np.array_equal(b, g["thsSmoothed"])
    for name, c in (("lidar", g["cloud"]), ("dense", None)):
        for pot, dens in CASES:
            s.currentPotential = pot; m, n = s.makeMaps(f[0], dens, 1, 1.0, c)
            assert np.array_equal(m.astype(np.uint8), g[f"maps_{name}_{pot}"]) and [n, s.currentPotential] == list(g[f"num_{name}_{pot}"]), (name, pot)
    sel = np.zeros((H, W), np.float32); lrud = list(g["lrud"])
    for k, (pot, add) in enumerate(((3, 1), (2, 0))):
        s.currentPotential = pot; T, _, _ = s.makeNewTraces(f[k], g["cloud"], orc.lidar_density(lrud, (W, H), 600.0), 600.0, add, sel)
        assert np.array_equal(_rows(T), g[f"traces{k}"]) and s.currentPotential == g[f"traces{k}_pot"][1], k
    assert np.array_equal(sel.astype(np.uint8), g["selection_map_final"])
    from sdv_loam_b200 import synth
    o = orc.LidarFrontEnd().handle(g["sweep"], synth.RLC, synth.TLC, K, (W, H), [10000, -1, 10000, -1])
    assert np.array_equal(o["cloud_px"], g["lidar_cloud_px"]) and np.array_equal(o["lrud"], g["lidar_lrud"]) and [o["addFeaturePoint"], o["n_segmented"]] == list(g["lidar_flags"])


@pytest.mark.gpu
def test_gpu_reproduces_reference_goldens():
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import api, synth
    g, img0, img1, K = _inputs(); ctx = api.Context(K, W, H, max_frames=4); ctx.makeImages(0, img0); ctx.makeImages(1, img1); rp = api.random_pattern(W, H)
    ps = api.PixelSelector(ctx, 1, rp); a, b = ps.makeHists(0); assert np.array_equal(a, g["ths"]) and np.array_equal(b, g["thsSmoothed"])
    for name, c in (("lidar", g["cloud"]), ("dense", None)):
        for pot, dens in CASES:
            ps.potential(0, pot); maps, num = ps.makeMapsBatch([0], [0], dens, 1, 1.0, None if c is None else [c])
            assert np.array_equal(np.asarray(maps[0]).reshape(g[f"maps_{name}_{pot}"].shape), g[f"maps_{name}_{pot}"]) and [int(num[0]), ps.potential(0)] == list(g[f"num_{name}_{pot}"]), (name, pot)
    ps = api.PixelSelector(ctx, 1, rp); lrud = list(g["lrud"])                                # fresh slot: empty persistent map
    for k, (pot, add) in enumerate(((3, 1), (2, 0))):
        ps.potential(0, pot); (T, I), _ = ps.makeNewTracesBatch([0], [k], [g["cloud"]], api.lidar_density(lrud, (W, H), 600.0), 600.0, add)[0][0], None
        assert np.array_equal(_rows(T), g[f"traces{k}"]) and ps.potential(0) == g[f"traces{k}_pot"][1], k
    assert np.array_equal(ps.selectionMap(0), g["selection_map_final"])
    o = api.LidarFrontEnd(ctx).handle([g["sweep"]], synth.RLC, synth.TLC, K, [[10000, -1, 10000, -1]])[0]
    assert np.array_equal(o["cloud_px"], g["lidar_cloud_px"]) and np.array_equal(o["lrud"], g["lidar_lrud"]) and [o["addFeaturePoint"], o["n_segmented"]] == list(g["lidar_flags"])
    ctx.close()
