"""The candidate-management CUDA source (sdv-loam_b200/csrc/sdv_select_core.cuh: kernels + host engine) run on the CPU through tests/emu/cuda_emu.hpp and compared with
the oracle (oracle/orc_select.cpp, itself pinned on the reference: tests/test_ref_pin_select.py).  The build container has no GPU; this is how the parallel decomposition
(per-cell direction masks, n2 resolution, 4x4-block replay, raster-order greedy suppression, BFS rings, activation walk) and the host orchestration are checked before the
-m gpu tests (tests/test_gpu_select.py) run the same source on the B200.  Everything is discrete or a short float expression: BIT FOR BIT."""
import numpy as np
import pytest
import orc
import select_emu as se
from conftest import cached_sequence, SMALL_K, SMALL_WH

W, H = SMALL_WH


@pytest.fixture(scope="module")
def scene():
    seq = cached_sequence(3, 3000, SMALL_K, SMALL_WH); L = orc.lib().orc_pyr_levels(W, H); rp = orc.libc_random_pattern(W, H)
    of = [orc.Frame(im, L) for im in seq.images]; F = [se.FrameImgs(f) for f in of]
    return seq, rp, of, F


def test_hists(scene):
    seq, rp, of, F = scene; E = se.Engine(W, H, rp); os_ = orc.Selector(W, H, rp)
    for k in range(2):
        a, b = E.makeHists(F[k]); oa, ob = os_.makeHists(of[k]); assert np.array_equal(a, oa) and np.array_equal(b, ob)


@pytest.mark.parametrize("lidar,dirDist", [(True, 1), (True, 0), (False, 1)])
def test_make_maps_batch(scene, lidar, dirDist):
    """six makeMaps calls with different start potentials / densities in ONE batch (recursion up and down, sub-selection), small scratch budget so the chunking runs"""
    seq, rp, of, F = scene; E = se.Engine(W, H, rp, dirDist); se.lib().emu_engine_max_scratch(E.p, 3 << 20)
    os_ = orc.Selector(W, H, rp); cloud = seq.clouds[0] if lidar else None
    pots = [3, 3, 1, 8, 2, 5, 4]; dens = [500., 60., 3000., 1500., 1e5, 333., 900.]; recs = [1, 1, 1, 1, 1, 1, 0]; ths = [1.0, 1.0, 1.0, 1.0, 1.0, 2.0, 1.0]
    if not lidar:                                                            # the dense pass runs one emulated warp per 4x4-cell block: keep the potentials >= 2 here (the GPU test runs all seven)
        keep = [0, 1, 3, 5, 6]; pots, dens, recs, ths = ([x[k] for k in keep] for x in (pots, dens, recs, ths))
    maps, num, pot, passes = E.makeMaps(F[0], pots, dens, recs, ths, cloud)
    if dirDist:
        for j in range(len(pots)):
            os_.currentPotential = pots[j]; m, n = os_.makeMaps(of[0], dens[j], recs[j], ths[j], cloud)
            assert n == num[j] and os_.currentPotential == pot[j] and np.array_equal(m.astype(np.uint8), maps[j]), j
        assert set(passes) == {1, 2}
    else:                                                                    # direction-free mode (dirNorm = gradient magnitude): self-consistency only
        assert all(np.count_nonzero(maps[j]) == num[j] for j in range(len(pots)))


def test_cloud_point_zero_quirk(scene):
    """`bestIdx > 0`: cloud row 0 can never be selected even when it is the strongest pixel of its cell"""
    seq, rp, of, F = scene; E = se.Engine(W, H, rp); os_ = orc.Selector(W, H, rp)
    g = of[0].absSquaredGrad(0); ys, xs = np.nonzero(g[8:H - 8, 8:W - 8] > 2000); cloud = seq.clouds[0].copy()
    cloud[0, 0] = xs[0] + 8 + 0.25; cloud[0, 1] = ys[0] + 8                    # a strong-gradient pixel first
    os_.currentPotential = 2; m, n = os_.makeMaps(of[0], 1e5, 0, 1.0, cloud)
    maps, num, pot, _ = E.makeMaps(F[0], [2], [1e5], [0], [1.0], cloud)
    assert m[0] == 0 and np.array_equal(m.astype(np.uint8), maps[0]) and n == num[0]


@pytest.mark.parametrize("add_feature", [0, 1])
def test_make_new_traces_batch(scene, add_feature):
    """three keyframes of three 'sequences' in one batch + a second round on the same selector slots (state carried over, stale monocular map when addFeaturePoint is off)"""
    seq, rp, of, F = scene; E = se.Engine(W, H, rp)
    slots = [se.Slot(p) for p in (3, 2, 4)]; osel = [orc.Selector(W, H, rp) for _ in range(3)]; omap = [np.zeros((H, W), np.float32) for _ in range(3)]
    for o, p in zip(osel, (3, 2, 4)): o.currentPotential = p
    dens = [600.0, 1500.0, 200.0]
    for rnd in range(2):
        order = [0, 1, 2] if rnd == 0 else [1, 2, 0]                          # round 2: other frames on the same slots
        clouds = [seq.clouds[k] for k in order]; lr = [[int(c[:, 0].min()), int(c[:, 0].max()), int(c[:, 1].min()), int(c[:, 1].max())] for c in clouds]
        dl = [orc.lidar_density(lr[j], SMALL_WH, dens[j]) for j in range(3)]; add = [add_feature and not (rnd == 1 and j == 1) for j in range(3)]
        outs, imms, num, passes = E.makeNewTraces(slots, [F[k] for k in order], clouds, dl, dens, add)
        for j in range(3):
            T, onum, opass = osel[j].makeNewTraces(of[order[j]], clouds[j], dl[j], dens[j], add[j], omap[j])
            assert len(T) == len(outs[j]) and np.array_equal(onum, num[j]) and np.array_equal(opass, passes[j]) and osel[j].currentPotential == slots[j].currentPotential, (rnd, j)
            assert T.tobytes() == outs[j].tobytes(), (rnd, j)
            assert np.array_equal(omap[j].astype(np.uint8), slots[j].map(W, H))
            P = orc.immature_init(of[order[j]], np.stack([T["u"], T["v"]], 1).astype(np.int32))
            for f in ("u", "v", "color", "weights", "gradH", "energyTH", "quality", "idepth_min"):
                assert np.array_equal(P[f], imms[j][f]), f
            assert np.isnan(imms[j]["idepth_max"]).all()
            if add_feature and add[j]: assert (T["isFromSensor"] == 0).sum() > 0


def _distmap_inputs(seq, newest, hosts, rng, n_pts=150):
    from sdv_loam_b200 import synth
    K0, K1 = orc.distmap_geometry(SMALL_K, None, None); KRKi, Kt = [], []
    Ki0 = np.linalg.inv(K0.astype(np.float64)).astype(np.float32)
    for hh in hosts:
        R, t = synth.rel_pose(seq.R[hh], seq.t[hh], seq.R[newest], seq.t[newest]); KRKi.append((K1 @ R.astype(np.float32) @ Ki0).astype(np.float32)); Kt.append((K1 @ t.astype(np.float32)).astype(np.float32))
    uvid, begin = [], [0]
    for hh in hosts:
        c = seq.clouds[hh]; pick = rng.choice(len(c), n_pts, replace=False); uvid.append(np.stack([np.floor(c[pick, 0]), np.floor(c[pick, 1]), 1.0 / c[pick, 2]], 1)); begin.append(begin[-1] + n_pts)
    return np.array(begin, np.int32), np.stack(KRKi), np.stack(Kt), np.concatenate(uvid).astype(np.float32)


def test_distance_map_and_activation_walk(scene):
    seq, rp, of, F = scene; E = se.Engine(W, H, rp); rng = np.random.default_rng(3)
    pb, KRKi, Kt, uvid = _distmap_inputs(seq, 2, [0, 1], rng)
    od = orc.DistMap(W >> 1, H >> 1); od.make(pb, KRKi, Kt, uvid)
    _, m = E.activate(pb, KRKi, Kt, uvid)
    assert np.array_equal(m[0], od.get()) and (m[0] == 0).sum() > 50
    # candidates of three hosts (the newest included, identity geometry), some projecting outside
    cb = [0]; cand = []
    for n in (500, 400, 150):
        u = rng.integers(4, W - 5, n); v = rng.integers(4, H - 5, n); idm = rng.uniform(0.01, 0.4, n).astype(np.float32); idm[:8] = 6.0
        cand.append(np.stack([u, v, idm, rng.choice([1.0, 2.0, 4.0], n)], 1).astype(np.float32)); cb.append(cb[-1] + n)
    cand = np.concatenate(cand); K0, K1 = orc.distmap_geometry(SMALL_K, None, None)
    cK = np.concatenate([KRKi, [(K1 @ np.linalg.inv(K0.astype(np.float64)).astype(np.float32)).astype(np.float32)]]); ct = np.concatenate([Kt, np.zeros((1, 3), np.float32)])
    for minDist, copies in ((1.0, 1), (2.5, 2)):
        se.lib().emu_engine_fuse_map(E.p, int(copies == 2))                     # both ways of building the map: 41 launches, or inside the walk kernel (SDV_FUSE_MAP)                              # the emulated walk is slow (one OS thread per CUDA thread); the GPU test runs four distances
        od.make(pb, KRKi, Kt, uvid); do = od.activateSelect(cb, cK, ct, cand, minDist)
        dec, m = E.activate(pb, KRKi, Kt, uvid, cb, cK, ct, cand, minDist, copies=copies)
        assert all(np.array_equal(dec[c], do) for c in range(copies)), minDist
        assert all(np.array_equal(m[c], od.get()) for c in range(copies))
        assert (do == 1).sum() > 15 and (do == -1).sum() >= 8
    assert (do == 0).sum() > 100


# ---------------------------------------------------------------------------------------------- LiDAR front-end (sdv_lidar_core.cuh)
def test_atan2f_matches_libm():
    """the device carries the C library's atan2f (fdlibm, float operations only): identical bits to this box's libm on random and special arguments"""
    import ctypes as C
    libm = C.CDLL("libm.so.6"); libm.atan2f.restype = C.c_float; libm.atan2f.argtypes = [C.c_float, C.c_float]; L = se.lib()
    rng = np.random.default_rng(0); y = (rng.uniform(-1, 1, 60000) * 10.0 ** rng.uniform(-6, 3, 60000)).astype(np.float32); x = (rng.uniform(-1, 1, 60000) * 10.0 ** rng.uniform(-6, 3, 60000)).astype(np.float32)
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-38, 3e38], np.float32); y[:81] = np.repeat(sp, 9); x[:81] = np.tile(sp, 9)
    a = np.array([libm.atan2f(float(p), float(q)) for p, q in zip(y, x)], np.float32); b = np.array([L.emu_atan2f(float(p), float(q)) for p, q in zip(y, x)], np.float32)
    assert np.array_equal(a.view(np.uint32)[~np.isnan(a)], b.view(np.uint32)[~np.isnan(a)]) and np.array_equal(np.isnan(a), np.isnan(b))


def test_lidar_front_end_batch():
    """three raw sweeps in one batch + a second call carrying the pixel box over: pixel rows {Ku, Kv, depth}, box, counts, addFeaturePoint identical to the oracle"""
    from test_ref_pin_lidar import sweeps
    synth, S = sweeps(3); fe = orc.LidarFrontEnd(); E = se.LidarEngine()
    lr0 = [[10000, -1, 10000, -1]] * 3
    for rnd in range(2):
        tlc = synth.TLC if rnd == 0 else np.array([0.0, -0.08, 0.35])
        G = E.handle(S, synth.RLC, tlc, SMALL_K, SMALL_WH, lr0)
        for j in range(3):
            o = fe.handle(S[j], synth.RLC, tlc, SMALL_K, SMALL_WH, lr0[j])
            assert np.array_equal(o["cloud_px"], G[j]["cloud_px"]) and len(o["cloud_px"]) > 3000, (rnd, j, len(o["cloud_px"]), len(G[j]["cloud_px"]))
            assert np.array_equal(o["lrud"], G[j]["lrud"]) and o["numGround"] == G[j]["numGround"] and o["n_segmented"] == G[j]["n_segmented"] and o["addFeaturePoint"] == G[j]["addFeaturePoint"]
        lr0 = [g["lrud"] for g in G]
    for cloud in (np.zeros((0, 4), np.float32), np.array([[np.nan, 0, 0, 0], [0.01, 0.01, 0, 0]], np.float32), S[0][::7]):
        o = fe.handle(cloud, synth.RLC, synth.TLC, SMALL_K, SMALL_WH, [10000, -1, 10000, -1]); g = E.handle([cloud], synth.RLC, synth.TLC, SMALL_K, SMALL_WH, [[10000, -1, 10000, -1]])[0]
        assert np.array_equal(o["cloud_px"], g["cloud_px"]) and np.array_equal(o["lrud"], g["lrud"]) and o["addFeaturePoint"] == g["addFeaturePoint"] and o["n_segmented"] == g["n_segmented"]


def test_lidar_front_end_128_beams():
    """S-STRESS sensor (BASELINE config #5: 128 beams over the same elevation span): sensor constants are parameters here (the reference hard-codes N_SCAN = 64, main.cpp:103)"""
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import synth
    world = synth.World(3000); R, t = synth.trajectory(1, 3000); sw = synth.lidar_sweep(world, R[0], t[0], beams=128, seed=3)
    args = dict(n_scan=128, horizon=1800, ang_res_x=0.2, ang_res_y=0.427 / 2, ang_bottom=24.9, groundScanInd=100)
    o = orc.LidarFrontEnd(**args).handle(sw, synth.RLC, synth.TLC, SMALL_K, SMALL_WH, [10000, -1, 10000, -1]); g = se.LidarEngine(**args).handle([sw], synth.RLC, synth.TLC, SMALL_K, SMALL_WH, [[10000, -1, 10000, -1]])[0]
    assert np.array_equal(o["cloud_px"], g["cloud_px"]) and len(o["cloud_px"]) > 6000 and np.array_equal(o["lrud"], g["lrud"]) and o["n_segmented"] == g["n_segmented"] and o["numGround"] == g["numGround"]


def test_make_new_traces_without_lidar_pixels(scene):
    """a keyframe whose sweep left no pixel in the image (empty cloud): the LiDAR selection recurses on zero picks, the monocular selection still runs"""
    seq, rp, of, F = scene; E = se.Engine(W, H, rp); slot = se.Slot(3); osel = orc.Selector(W, H, rp); omap = np.zeros((H, W), np.float32)
    empty = np.zeros((0, 3)); outs, imms, num, passes = E.makeNewTraces([slot], [F[0]], [empty], [300.0], [600.0], [1])
    T, onum, opass = osel.makeNewTraces(of[0], empty, 300.0, 600.0, 1, omap)
    assert T.tobytes() == outs[0].tobytes() and np.array_equal(onum, num[0]) and osel.currentPotential == slot.currentPotential and len(T) > 100 and (T["isFromSensor"] == 0).all()
    assert np.array_equal(omap.astype(np.uint8), slot.map(W, H))
