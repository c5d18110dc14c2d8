"""f2 on the GPU (-m gpu): ImmaturePoint constructor + traceOn through the C-ABI vs the oracle (which tests/test_ref_pin_trace.py pins bit for bit on the
reference's own compiled ImmaturePoint.cpp).  Every field of every candidate must be IDENTICAL after construction and after three consecutive traced frames,
for several host keyframes in one launch (the loop of FullSystem::traceNewCoarse) and at two image sizes."""
import numpy as np
import pytest
import orc
from conftest import cached_sequence, SMALL_K, SMALL_WH
from test_ref_pin_trace import candidates, flat

pytestmark = pytest.mark.gpu


def _mods():
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import api, synth
    return api, synth


@pytest.mark.parametrize("wh,K,seed,n", [(SMALL_WH, SMALL_K, 3000, 300), ((1200, 360), None, 2000, 1500)])
def test_init_and_trace_bit_exact(wh, K, seed, n):
    api, synth = _mods(); K = K or synth.KITTI_K; w, h = wh; L = api.pyr_levels(w, h)
    seq = cached_sequence(5 if wh == SMALL_WH else 8, seed, K, wh)
    of = [orc.Frame(im, L) for im in seq.images[:5]]
    ctx = api.Context(K, w, h, max_frames=8)
    for i in range(5):
        ctx.makeImages(10 + i, seq.images[i])
    poses = [np.concatenate([synth._quat_from_R(seq.R[i]), seq.t[i]]) for i in range(5)]
    hosts = [0, 1]                                                                                  # two host keyframes, their candidates traced into frames 2, 3, 4
    Po = []; Pg = []
    for hidx in hosts:
        uv = candidates(seq, n, seed + hidx)
        po = orc.immature_init(of[hidx], uv); pg = api.immatureInit(ctx, 10 + hidx, uv)
        assert api.IMMATURE_PT_DTYPE.itemsize == orc.IMM_DTYPE.itemsize == 120
        assert np.array_equal(pg.view(np.uint8), po.view(np.uint8)), hidx                          # byte for byte (NaN idepth_max included)
        Po.append(po); Pg.append(pg)
    pg_all = np.concatenate(Pg); pb = np.cumsum([0] + [len(p) for p in Pg]).astype(np.int32); seen = set()
    for k, (ab, ex) in zip((2, 3, 4), (((0.0, 0.0), 1.0), ((0.02, -1.5), 1.1), ((-0.01, 2.0), 0.9))):
        geo = [orc.trace_geometry(K, poses[hidx], poses[k], 1.0, ex, (0.0, 0.0), ab) for hidx in hosts]
        so = [orc.immature_trace(of[k], Po[j], *geo[j]) for j in range(len(hosts))]
        sg = api.traceOnBatch(ctx, [10 + k] * len(hosts), pb, np.stack([g[0] for g in geo]), np.stack([g[1] for g in geo]), np.stack([g[2] for g in geo]), pg_all)
        assert np.array_equal(sg, np.concatenate(so)), k
        for j in range(len(hosts)):
            a = pg_all[pb[j]:pb[j + 1]]; bad = [i for i in range(len(a)) if not np.array_equal(flat(a[i]), flat(Po[j][i]), equal_nan=True)]
            assert not bad, (k, j, bad[:5], flat(a[bad[0]]) - flat(Po[j][bad[0]]))
            assert np.array_equal(a["lastTraceStatus"], Po[j]["lastTraceStatus"])
        seen |= set(int(s) for s in sg)
    assert {api.IPS_GOOD, api.IPS_OOB, api.IPS_OUTLIER, api.IPS_SKIPPED} <= seen, seen
    ctx.close()


def test_trace_argument_errors_and_empty_groups():
    api, synth = _mods(); w, h = SMALL_WH
    seq = cached_sequence(3, 1000, SMALL_K, SMALL_WH); ctx = api.Context(SMALL_K, w, h, max_frames=4)
    ctx.makeImages(1, seq.images[0]); ctx.makeImages(2, seq.images[1])
    with pytest.raises(api.SdvError):
        api.immatureInit(ctx, 1, [[1, 50]])                                                          # too close to the border
    with pytest.raises(api.SdvError):
        api.immatureInit(ctx, 99, [[50, 50]])                                                        # unknown keyframe
    P = api.immatureInit(ctx, 1, [[50, 50], [100, 80]])
    I3 = np.eye(3, dtype=np.float32); z3 = np.zeros(3, np.float32); a2 = np.array([1, 0], np.float32)
    with pytest.raises(api.SdvError):
        api.traceOnBatch(ctx, [7], [0, 2], I3[None], z3[None], a2[None], P)                          # unknown traced frame
    st = api.traceOnBatch(ctx, [2, 2], [0, 0, 2], np.stack([I3, I3]), np.stack([z3, z3]), np.stack([a2, a2]), P)   # an empty group in front
    assert len(st) == 2
    assert api.traceOnBatch(ctx, [], [0], np.zeros((0, 9)), np.zeros((0, 3)), np.zeros((0, 2)), P[:0]).size == 0
    ctx.close()


def test_optimize_immature_point_bit_exact():
    """FullSystem::optimizeImmaturePoint on the device vs the oracle (pinned on the reference by tests/test_ref_pin_trace.py): two host keyframes of a 5-keyframe window
    in ONE launch; status, activated inverse depth and every temporary-residual state identical."""
    api, synth = _mods(); w, h = SMALL_WH; nF = 5
    from test_ref_pin_ba import _window
    import ref
    if not ref.available():
        pytest.skip("needs oracle/_ref for the window's precalc (prebuilt library travels to the GPU box)")
    win, ob, rb, (of, rf) = _window((0, 1, 2, 3, 4), 5)
    seq = cached_sequence(5, 3000, SMALL_K, SMALL_WH)
    ctx = api.Context(SMALL_K, w, h, max_frames=8)
    for i in range(nF):
        ctx.makeImages(40 + i, seq.images[i])
    poses = [np.concatenate([synth._quat_from_R(seq.R[i]), seq.t[i]]) for i in range(nF)]
    groups = []
    for host, tgt in ((0, 1), (2, 3)):
        pre, cal = rb.immature_pre(host, nF)
        sh = type(seq).__new__(type(seq)); sh.wh = seq.wh; sh.images = [seq.images[host]]; uv = candidates(sh, 400, 70 + host)
        P = orc.immature_init(of[host], uv); orc.immature_trace(of[tgt], P, *orc.trace_geometry(SMALL_K, poses[host], poses[tgt]))
        rng = np.random.default_rng(host); sensor = rng.uniform(size=len(P)) < 0.15
        bad = ~np.isfinite(P["idepth_max"]); P["idepth_max"][bad] = 0.2; P["idepth_min"][bad] = 0.0
        wide = rng.uniform(size=len(P)) < 0.2; P["idepth_min"][wide] *= np.float32(0.5); P["idepth_max"][wide] *= np.float32(1.7)
        groups.append(dict(host=host, pre=pre, cal=cal, P=P, sensor=sensor, targets=[t for t in range(nF) if t != host]))
    for min_obs in (1, 3):
        want = [orc.immature_optimize(g["P"], g["sensor"], [of[t] for t in g["targets"]], g["pre"], g["cal"], min_obs) for g in groups]
        pts = np.concatenate([g["P"] for g in groups]).view(api.IMMATURE_PT_DTYPE)
        pb = np.cumsum([0] + [len(g["P"]) for g in groups]); tb = np.cumsum([0] + [len(g["targets"]) for g in groups])
        st, idp, rs = api.optimizeImmaturePointBatch(ctx, pb, tb, [40 + t for g in groups for t in g["targets"]], np.concatenate([g["pre"] for g in groups]),
                                                     np.stack([g["cal"] for g in groups]), pts, np.concatenate([g["sensor"] for g in groups]), min_obs)
        so = np.concatenate([x[0] for x in want]); io = np.concatenate([x[1] for x in want]); ro = np.concatenate([x[2] for x in want])
        assert np.array_equal(st, so), np.nonzero(st != so)[0][:10]
        act = so == 1
        assert np.array_equal(idp[act], io[act]) and np.array_equal(rs[act], ro[act])
        assert {-1, 0, 1} <= set(int(s) for s in so) or min_obs == 1
    assert (so == 1).sum() > 50
    ctx.close()
