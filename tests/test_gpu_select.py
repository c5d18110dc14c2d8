"""f4 + the caller half of f2 on the GPU (-m gpu): PixelSelector (makeHists, makeMapsFromLidar, makeMaps), FullSystem::makeNewTraces, CoarseDistanceMap and the candidate
walk of activatePointsMT through the C-ABI vs the oracle (tests/test_ref_pin_select.py pins it bit for bit on the reference's own compiled code).  Selection maps,
counts, potentials, Shi-Tomasi scores, point types, immature-point records, distance maps and decisions must be IDENTICAL.  The same CUDA source runs on the CPU
(emulated) in tests/test_select_emu_cpu.py; here it runs on the B200, batched, at two image sizes."""
import numpy as np
import pytest
import orc
from conftest import cached_sequence, SMALL_K, SMALL_WH

pytestmark = pytest.mark.gpu


def _mods():
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import api, synth
    return api, synth


def _scene(wh, K, seed, nfr=3):
    api, synth = _mods(); K = K or synth.KITTI_K; w, h = wh; L = api.pyr_levels(w, h)
    seq = cached_sequence(nfr, seed, K, wh); of = [orc.Frame(im, L) for im in seq.images[:nfr]]
    ctx = api.Context(K, w, h, max_frames=8)
    for i in range(nfr): ctx.makeImages(10 + i, seq.images[i])
    rp = api.random_pattern(w, h)
    return api, seq, of, ctx, rp


@pytest.mark.parametrize("wh,K,seed", [(SMALL_WH, SMALL_K, 3000), ((1200, 360), None, 1000)])
def test_hists_and_make_maps(wh, K, seed):
    api, seq, of, ctx, rp = _scene(wh, K, seed); w, h = wh
    assert np.array_equal(rp, orc.libc_random_pattern(w, h))
    pots = [3, 3, 1, 8, 2, 5, 4]; dens = [500., 60., 3000., 1500., 1e5, 333., 900.]; recs = [1, 1, 1, 1, 1, 1, 0]; ths = [1.0, 1.0, 1.0, 1.0, 1.0, 2.0, 1.0]; n = len(pots)
    ps = api.PixelSelector(ctx, n, rp); os_ = orc.Selector(w, h, rp)
    for k in range(2):
        a, b = ps.makeHists(10 + k); oa, ob = os_.makeHists(of[k]); assert np.array_equal(a, oa) and np.array_equal(b, ob)
    for lidar in (True, False):
        for j in range(n): ps.potential(j, pots[j])
        clouds = [seq.clouds[0]] * n if lidar else None
        maps, num = ps.makeMapsBatch(list(range(n)), [10] * n, dens, recs, ths, clouds)
        for j in range(n):
            os_.currentPotential = pots[j]; m, nh = os_.makeMaps(of[0], dens[j], recs[j], ths[j], seq.clouds[0] if lidar else None)
            assert nh == num[j] and os_.currentPotential == ps.potential(j), (lidar, j, nh, num[j])
            assert np.array_equal(m.astype(np.uint8), maps[j]), (lidar, j)
            if not lidar: assert np.array_equal(ps.selectionMap(j), maps[j])
    ctx.close()


@pytest.mark.parametrize("wh,K,seed", [(SMALL_WH, SMALL_K, 3000), ((1200, 360), None, 1000)])
@pytest.mark.parametrize("add_feature", [0, 1])
def test_make_new_traces_batch(wh, K, seed, add_feature):
    api, seq, of, ctx, rp = _scene(wh, K, seed); w, h = wh
    ps = api.PixelSelector(ctx, 3, rp); osel = [orc.Selector(w, h, rp) for _ in range(3)]; omap = [np.zeros((h, w), np.float32) for _ in range(3)]
    for j, p in enumerate((3, 2, 4)): ps.potential(j, p); osel[j].currentPotential = p
    dens = [600.0, 1500.0, 200.0]
    for rnd in range(2):
        order = [0, 1, 2] if rnd == 0 else [1, 2, 0]
        clouds = [seq.clouds[k] for k in order]; lr = [[int(c[:, 0].min()), int(c[:, 0].max()), int(c[:, 1].min()), int(c[:, 1].max())] for c in clouds]
        dl = [api.lidar_density(lr[j], wh, dens[j]) for j in range(3)]; add = [int(add_feature and not (rnd == 1 and j == 1)) for j in range(3)]
        res, num = ps.makeNewTracesBatch([0, 1, 2], [10 + k for k in order], clouds, dl, dens, add, cap=1 << 15)
        for j in range(3):
            T, onum, _ = osel[j].makeNewTraces(of[order[j]], clouds[j], dl[j], dens[j], add[j], omap[j]); G, I = res[j]
            assert len(T) == len(G) and np.array_equal(onum, num[j]) and osel[j].currentPotential == ps.potential(j), (rnd, j, len(T), len(G))
            assert T.tobytes() == G.tobytes(), (rnd, j)
            assert np.array_equal(omap[j].astype(np.uint8), ps.selectionMap(j))
            P = orc.immature_init(of[order[j]], np.stack([T["u"], T["v"]], 1).astype(np.int32))
            assert np.array_equal(P.view(np.uint8), I.view(np.uint8)), (rnd, j)                    # the whole ImmaturePoint record, byte for byte
            assert (T["type"] == 0).sum() > 0 and (T["type"] == 1).sum() > 0
            if add[j]: assert (T["isFromSensor"] == 0).sum() > 0
    assert ctx.last_kernel_ms() > 0
    ctx.close()


def test_distance_map_and_activation_walk_batch():
    from test_select_emu_cpu import _distmap_inputs
    api, seq, of, ctx, rp = _scene(SMALL_WH, SMALL_K, 3000); w, h = SMALL_WH; rng = np.random.default_rng(3)
    api.PixelSelector(ctx, 1, rp)
    pb, KRKi, Kt, uvid = _distmap_inputs(seq, 2, [0, 1], rng)
    od = orc.DistMap(w >> 1, h >> 1); od.make(pb, KRKi, Kt, uvid)
    dec, m = api.activateSelectBatch(ctx, [dict(pt_begin=pb, KRKi=KRKi, Kt=Kt, uvid=uvid)], want_maps=True)
    assert np.array_equal(m[0], od.get()) and (m[0] == 0).sum() > 50
    cb = [0]; cand = []
    for n in (500, 400, 150):
        u = rng.integers(4, w - 5, n); v = rng.integers(4, h - 5, n); idm = rng.uniform(0.01, 0.4, n).astype(np.float32); idm[:8] = 6.0
        cand.append(np.stack([u, v, idm, rng.choice([1.0, 2.0, 4.0], n)], 1).astype(np.float32)); cb.append(cb[-1] + n)
    cand = np.concatenate(cand); K0, K1 = orc.distmap_geometry(SMALL_K, None, None)
    cK = np.concatenate([KRKi, [(K1 @ np.linalg.inv(K0.astype(np.float64)).astype(np.float32)).astype(np.float32)]]); ct = np.concatenate([Kt, np.zeros((1, 3), np.float32)])
    dists = (0.0, 1.0, 2.5, 4.0)
    seqs = [dict(pt_begin=pb, KRKi=KRKi, Kt=Kt, uvid=uvid, cand_begin=cb, cKRKi=cK, cKt=ct, cand4=cand, minActDist=d) for d in dists]     # four sequences in ONE call
    decs, maps = api.activateSelectBatch(ctx, seqs, want_maps=True)
    for j, d in enumerate(dists):
        od.make(pb, KRKi, Kt, uvid); do = od.activateSelect(cb, cK, ct, cand, d)
        assert np.array_equal(decs[j], do), d
        assert np.array_equal(maps[j], od.get()), d
        assert (do == 1).sum() > 15 and (do == -1).sum() >= 8
    ctx.close()


def test_lidar_front_end_batch():
    """the node's lidarCloudHandler on raw XYZI sweeps (three sequences per call, pixel box carried over a second call): rows {Ku, Kv, depth}, box, ground count, segmented
    size and addFeaturePoint identical to the oracle (pinned on the reference's own src/main.cpp, tests/test_ref_pin_lidar.py); then the rows feed makeNewTraces"""
    from test_ref_pin_lidar import sweeps
    api, seq, of, ctx, rp = _scene(SMALL_WH, SMALL_K, 3000); synth, S = sweeps(3)
    fe = orc.LidarFrontEnd(); G = api.LidarFrontEnd(ctx); lr0 = [[10000, -1, 10000, -1]] * 3
    for rnd in range(2):
        tlc = synth.TLC if rnd == 0 else np.array([0.0, -0.08, 0.35])
        res = G.handle(S, synth.RLC, tlc, SMALL_K, lr0)
        for j in range(3):
            o = fe.handle(S[j], synth.RLC, tlc, SMALL_K, SMALL_WH, lr0[j])
            assert np.array_equal(o["cloud_px"], res[j]["cloud_px"]) and len(o["cloud_px"]) > 3000, (rnd, j, len(o["cloud_px"]), len(res[j]["cloud_px"]))
            assert np.array_equal(o["lrud"], res[j]["lrud"]) and o["numGround"] == res[j]["numGround"] and o["n_segmented"] == res[j]["n_segmented"] and o["addFeaturePoint"] == res[j]["addFeaturePoint"]
        lr0 = [r["lrud"] for r in res]
    for cloud in (np.zeros((0, 4), np.float32), np.array([[np.nan, 0, 0, 0], [0.01, 0.01, 0, 0]], np.float32), S[0][::7]):
        o = fe.handle(cloud, synth.RLC, synth.TLC, SMALL_K, SMALL_WH, [10000, -1, 10000, -1]); g = G.handle([cloud], synth.RLC, synth.TLC, SMALL_K, [[10000, -1, 10000, -1]])[0]
        assert np.array_equal(o["cloud_px"], g["cloud_px"]) and np.array_equal(o["lrud"], g["lrud"]) and o["addFeaturePoint"] == g["addFeaturePoint"] and o["n_segmented"] == g["n_segmented"]
    # S-STRESS sensor: 128 rings (the sensor constants are parameters of sdv_lidar_init)
    world = synth.World(3000); Rt = synth.trajectory(1, 3000); sw128 = synth.lidar_sweep(world, Rt[0][0], Rt[1][0], beams=128, seed=3)
    args = dict(n_scan=128, horizon=1800, ang_res_x=0.2, ang_res_y=0.427 / 2, ang_bottom=24.9, groundScanInd=100)
    o = orc.LidarFrontEnd(**args).handle(sw128, synth.RLC, synth.TLC, SMALL_K, SMALL_WH, [10000, -1, 10000, -1]); g = api.LidarFrontEnd(ctx, **args).handle([sw128, sw128[::3]], synth.RLC, synth.TLC, SMALL_K, [[10000, -1, 10000, -1]] * 2)[0]
    assert np.array_equal(o["cloud_px"], g["cloud_px"]) and len(o["cloud_px"]) > 6000 and np.array_equal(o["lrud"], g["lrud"]) and o["n_segmented"] == g["n_segmented"]
    G = api.LidarFrontEnd(ctx)
    # front-end -> selector: the device-produced pixel rows drive makeNewTraces of the same keyframe
    w, h = SMALL_WH; res = G.handle([S[0]], synth.RLC, synth.TLC, SMALL_K, [[10000, -1, 10000, -1]])[0]
    ps = api.PixelSelector(ctx, 1, rp); osel = orc.Selector(w, h, rp); dl = api.lidar_density(res["lrud"], SMALL_WH, 600.0)
    (T, I), num = ps.makeNewTracesBatch([0], [10], [res["cloud_px"]], dl, 600.0, res["addFeaturePoint"])[0][0], None
    To, _, _ = osel.makeNewTraces(of[0], res["cloud_px"], dl, 600.0, res["addFeaturePoint"], np.zeros((h, w), np.float32))
    assert To.tobytes() == T.tobytes() and len(T) > 100
    ctx.close()


def test_error_behaviour():
    """the reference's asserts become error codes: calls before init, bad slots, a slot twice in one batch, an output capacity that is too small"""
    api, seq, of, ctx, rp = _scene(SMALL_WH, SMALL_K, 3000, nfr=1); cloud = seq.clouds[0]
    with pytest.raises(api.SdvError): api.activateSelectBatch(ctx, [dict(pt_begin=[0], KRKi=np.zeros((0, 9)), Kt=np.zeros((0, 3)), uvid=np.zeros((0, 3)))])        # no selector yet
    with pytest.raises(api.SdvError): api.LidarFrontEnd(ctx, n_scan=200)                                                                                          # more rings than the row mask holds
    ps = api.PixelSelector(ctx, 2, rp)
    with pytest.raises(api.SdvError): ps.potential(5)
    with pytest.raises(api.SdvError): ps.makeNewTracesBatch([0, 0], [10, 10], [cloud, cloud], 300.0, 600.0, 1)                                                  # one selector cannot serve two keyframes at once
    with pytest.raises(api.SdvError): ps.makeNewTracesBatch([0], [999], [cloud], 300.0, 600.0, 1)                                                               # unknown frame
    with pytest.raises(api.SdvError): ps.makeNewTracesBatch([1], [10], [cloud], 300.0, 600.0, 1, cap=16)                                                        # SDV_ERR_CAPACITY
    res, num = ps.makeNewTracesBatch([1], [10], [cloud], 300.0, 600.0, 1); assert len(res[0][0]) > 100                                                           # the context is still usable
    ctx.close()
