"""Pins the candidate-management restatement (oracle/orc_select.cpp: PixelSelector, shiTomasiScore, makeNewTraces, CoarseDistanceMap and the candidate walk of
activatePointsMT — SURVEY §8f rank 4 and the caller half of rank 2) on the reference's own compiled PixelSelector2.cpp / FullSystem.cpp / CoarseTracker.cpp (oracle/_ref).
Everything here is discrete (selection maps, counts, potentials, distance maps, decisions) or a short float expression (thresholds, Shi-Tomasi score): BIT FOR BIT."""
import numpy as np
import pytest
import orc
import ref
from conftest import cached_sequence, SMALL_K, SMALL_WH

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference absent")


def _setup(wh, K, seed, n=3):
    from sdv_loam_b200 import synth
    K = K or synth.KITTI_K
    seq = cached_sequence(n, seed, K, wh); L = ref.set_calib(wh[0], wh[1], K)
    rs = ref.Selector(wh); rp = rs.randomPattern()
    return seq, L, rs, orc.Selector(wh[0], wh[1], rp), rp


def test_random_pattern_is_glibc_rand():
    """the host mirror regenerates PixelSelector::randomPattern (srand(3141592), rand() & 0xFF) with the C library's rand()"""
    ref.set_calib(SMALL_WH[0], SMALL_WH[1], SMALL_K)
    assert np.array_equal(ref.Selector(SMALL_WH).randomPattern(), orc.libc_random_pattern(*SMALL_WH))


@pytest.mark.parametrize("wh,K,seed", [(SMALL_WH, SMALL_K, 3000), ((1200, 360), None, 1000)])
def test_hists_and_select_passes_bit_exact(wh, K, seed):
    seq, L, rs, os_, _ = _setup(wh, K, seed)
    for k in range(2):
        of, rf = orc.Frame(seq.images[k], L), ref.Frame(seq.images[k], wh, L)
        a, b = os_.makeHists(of); ra, rb = rs.makeHists(rf)
        assert np.array_equal(a, ra) and np.array_equal(b, rb) and a.min() >= 3
        cloud = seq.clouds[k]
        for pot in (1, 2, 3, 4, 7):
            for th in (1.0, 2.0):
                m, n3 = os_.select(of, pot, th, cloud); rm, rn3 = rs.select(rf, pot, th, cloud)
                assert np.array_equal(n3, rn3) and np.array_equal(m, rm), ("lidar", pot, th, n3, rn3)
                m, n3 = os_.select(of, pot, th); rm, rn3 = rs.select(rf, pot, th)
                assert np.array_equal(n3, rn3) and np.array_equal(m, rm), ("dense", pot, th, n3, rn3)
        assert n3.sum() > 0


@pytest.mark.parametrize("wh,K,seed", [(SMALL_WH, SMALL_K, 3000), ((1200, 360), None, 1000)])
def test_make_maps_bit_exact(wh, K, seed):
    """makeMaps / makeMapsFromLidar incl. the potential recursion in both directions and the random sub-selection; the selector state (currentPotential) carries over"""
    seq, L, rs, os_, _ = _setup(wh, K, seed)
    of, rf = orc.Frame(seq.images[0], L), ref.Frame(seq.images[0], wh, L); cloud = seq.clouds[0]
    seen = set()
    for start_pot, density in ((3, 500.0), (3, 60.0), (1, 3000.0), (8, 1500.0), (2, 1e5), (5, 333.0)):
        for lidar in (True, False):
            os_.currentPotential = start_pot; rs.currentPotential = start_pot
            for rec in (1, 0):
                m, n = os_.makeMaps(of, density, rec, 1.0, cloud if lidar else None); rm, rn = rs.makeMaps(rf, density, rec, 1.0, cloud if lidar else None)
                assert n == rn and np.array_equal(m, rm) and os_.currentPotential == rs.currentPotential, (start_pot, density, lidar, rec, n, rn)
                assert n == np.count_nonzero(m)
                seen.add(os_.currentPotential)
    assert len(seen) >= 4, seen


def test_shi_tomasi_bit_exact():
    from test_ref_pin_ba import _window
    win, ob, rb, (of, rf) = _window((0, 1), 5)
    w, h = SMALL_WH; rng = np.random.default_rng(0)
    uv = np.concatenate([rng.integers(0, [w, h], (400, 2)), [[4, 4], [5, 5], [w - 6, h - 6], [w - 5, h - 5], [5, h - 6]]])
    so = np.array([orc.shi_tomasi(of[0], u, v) for u, v in uv], np.float32); sr = np.array([ref.shi_tomasi(rb, rf[0], u, v) for u, v in uv], np.float32)
    assert np.array_equal(so, sr) and (so != 0).sum() > 300 and (so == 0).sum() >= 2


@pytest.mark.parametrize("add_feature", [False, True])
def test_make_new_traces_bit_exact(add_feature):
    """FullSystem::makeNewTraces: LiDAR selection, Shi-Tomasi typing, monocular selection with the occupancy mask, and the stale monocular map walked when addFeaturePoint is off"""
    from test_ref_pin_ba import _window
    win, ob, rb, (of, rf) = _window((0, 1, 2), 5)
    seq = cached_sequence(3, 3000, SMALL_K, SMALL_WH); w, h = SMALL_WH
    rs = ref.Selector(SMALL_WH, owner=rb); os_ = orc.Selector(w, h, rs.randomPattern())
    sel_o = np.zeros((h, w), np.float32); sel_r = np.zeros((h, w), np.float32)
    for k, (start_pot, dens) in enumerate(((3, 600.0), (2, 1500.0), (4, 200.0))):
        cloud = seq.clouds[k]; lrud = [int(cloud[:, 0].min()), int(cloud[:, 0].max()), int(cloud[:, 1].min()), int(cloud[:, 1].max())]
        os_.currentPotential = start_pot; rs.currentPotential = start_pot
        add = add_feature and k != 1                                        # k == 1 with addFeaturePoint off: the map of keyframe 0 is walked again
        T, num, passes = os_.makeNewTraces(of[k], cloud, orc.lidar_density(lrud, SMALL_WH, dens), dens, add, sel_o)
        R = ref.make_new_traces(rb, rf[k], cloud, lrud, add, dens, sel_r)
        assert len(T) == len(R) and os_.currentPotential == rs.currentPotential, (k, len(T), len(R))
        assert np.array_equal(sel_o, sel_r)
        for f, col in (("u", 0), ("v", 1), ("my_type", 2), ("score", 3), ("idepth_fromSensor", 4)):
            assert np.array_equal(T[f], R[:, col]), (k, f)
        assert np.array_equal(T["isFromSensor"], R[:, 5].astype(np.int32)) and np.array_equal(T["type"], R[:, 6].astype(np.int32))
        assert T["isFromSensor"].sum() == num[0] and (T["type"] == 1).sum() > 0 and (T["type"] == 0).sum() > 0
        if add_feature:
            assert (T["isFromSensor"] == 0).sum() > 0 and (T["isFromSensor"] == 0).sum() < max(num[1], np.count_nonzero(sel_o))   # some monocular points exist, some were masked out


def _activation_inputs(win, nF, newest, rng, n_cand=600):
    """candidates per host: integer host pixels with a depth interval around plausible inverse depths, a few that project outside the newest frame"""
    w, h = SMALL_WH; cands, begin = [], [0]
    for hI in range(nF):
        n = n_cand if hI != newest else n_cand // 3
        u = rng.integers(4, w - 5, n); v = rng.integers(4, h - 5, n); idm = rng.uniform(0.01, 0.4, n).astype(np.float32)
        idm[:10] = 5.0                                                      # very close points: project far outside
        typ = rng.choice([1.0, 2.0, 4.0], n)
        cands.append(np.stack([u, v, idm, typ], 1).astype(np.float32)); begin.append(begin[-1] + n)
    return np.concatenate(cands), np.array(begin, np.int32)


@pytest.mark.parametrize("kf,newest", [((0, 1, 2, 3, 4), 4), ((0, 2, 3), 2)])
def test_distance_map_and_activation_walk_bit_exact(kf, newest):
    """makeDistanceMap (window's ACTIVE points forward-warped into the newest keyframe, 39 BFS rings alternating 4/8-connectivity), addIntoDistFinal, and the
    greedy candidate walk of activatePointsMT for four values of currentMinActDist"""
    from test_ref_pin_ba import _window
    win, ob, rb, (of, rf) = _window(kf, 5)
    nF = win["nF"]; w, h = SMALL_WH; rd = ref.DistMap(rb, newest, SMALL_WH); od = orc.DistMap(w >> 1, h >> 1)
    rd.make()
    hosts = [i for i in range(nF) if i != newest]
    geo = [rd.geometry(i) for i in hosts]; KRKi = np.stack([g[0] for g in geo]); Kt = np.stack([g[1] for g in geo])
    uvid = []; pt_begin = [0]
    for i in hosts:
        sel = win["host"] == i; uvid.append(np.concatenate([win["uv"][sel], win["idepth"][sel, None]], 1)); pt_begin.append(pt_begin[-1] + int(sel.sum()))
    od.make(pt_begin, KRKi, Kt, np.concatenate(uvid).astype(np.float32))
    a, b = od.get(), rd.get()
    assert np.array_equal(a, b) and (a == 0).sum() > 50 and a.max() == 1000 or a.max() <= 39
    for (u, v) in ((5, 5), (w // 4, h // 4), (1, 1), ((w >> 1) - 2, (h >> 1) - 2)):
        od.add(u, v); rd.add(u, v)
    assert np.array_equal(od.get(), rd.get())
    rng = np.random.default_rng(7); cand, begin = _activation_inputs(win, nF, newest, rng)
    all_hosts = list(range(nF)); geo = [rd.geometry(i) for i in all_hosts]; KRKi = np.stack([g[0] for g in geo]); Kt = np.stack([g[1] for g in geo])
    for minDist in (0.0, 1.0, 2.0, 4.0):
        rd.make(); od.make(pt_begin, np.stack([rd.geometry(i)[0] for i in hosts]), np.stack([rd.geometry(i)[1] for i in hosts]), np.concatenate(uvid).astype(np.float32))
        do = od.activateSelect(begin, KRKi, Kt, cand, minDist); dr = rd.activateSelect(all_hosts, begin, cand, minDist)
        assert np.array_equal(do, dr), minDist
        assert np.array_equal(od.get(), rd.get())
        assert (do == 1).sum() > 20 and (do == -1).sum() >= 10
    assert (do == 0).sum() > 100
