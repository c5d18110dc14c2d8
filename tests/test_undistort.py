"""f3 — the ingest in front of FullSystem::addActiveFrame: Undistort::undistort<unsigned char> (util/Undistort.cpp:341-435) + PhotometricUndistorter::processFrame (:177-214).

  * CPU: sdv_loam_b200.undistort (host mirror of the calibration-time set-up) reproduces the reference's K and remap tables BIT FOR BIT for every calibration file the
    reference ships (texts embedded below; oracle/_ref runs the reference's own Undistort), and its numpy restatement of undistort<> equals the reference's output.
  * CPU without the reference: the same against tests/golden/undistort_small.npz (made by tests/golden/make_undistort_golden.py from the reference).
  * GPU: raw mono8 -> sdv_set_undistort + sdv_frame_upload_batch_raw_u8 -> level-0 plane and every pyramid level equal the reference's undistorted image pushed through
    the oracle's makeImages, bit for bit; with a response function / vignette the photometric branch (:193-203) as well.
"""
import os
import numpy as np
import pytest
import orc
import ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "undistort_small.npz")
CALIBS = {   # verbatim calibration texts of the reference (calib/KITTI/00.txt, 03.txt, calib/kitti_360.txt, calib/kitti_carla.txt) — 4-line data files
    "kitti00": "Pinhole 718.856 718.856 607.1928 185.2157 0\n1241 376\ncrop\n1200 360\n",
    "kitti03": "Pinhole 721.5377 721.5377 609.5593 172.854 0\n1242 375\ncrop\n1200 360\n",
    "kitti360": "Pinhole 552.554261 552.554261 682.049453 238.769549 0\n1408 376\ncrop\n1400 360\n",
    "carla": "Pinhole 957.962 957.962 696.0 512.0 0\n1392 1024\ncrop\n1392 1024\n",
    "explicit": "Pinhole 0.58 1.9 0.49 0.5 0\n1241 376\n0.6 2.0 0.5 0.5 0\n1200 360\n",        # relative input format + explicit output calibration
    "none": "Pinhole 718.856 718.856 607.1928 185.2157 0\n640 192\nnone\n640 192\n",
}
SMALL = "Pinhole 130.5 131.0 70.3 43.9 0\n142 90\ncrop\n128 80\n"

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference absent")


def _mirror(text):
    from sdv_loam_b200 import undistort
    return undistort.Undistort.from_text(text)


@needs_ref
@pytest.mark.parametrize("name", sorted(CALIBS))
def test_host_mirror_matches_reference_tables(name):
    r = ref.Undistort(CALIBS[name]); u = _mirror(CALIBS[name])
    assert (u.wOrg, u.hOrg, u.w, u.h) == (r.wOrg, r.hOrg, r.w, r.h) and u.passthrough == r.passthrough
    assert np.array_equal(np.array([u.K[0, 0], u.K[1, 1], u.K[0, 2], u.K[1, 2]]), r.K4d)          # doubles, bit for bit
    assert np.array_equal(u.remapX, r.remapX) and np.array_equal(u.remapY, r.remapY)
    raw = np.random.default_rng(3).integers(0, 256, (r.hOrg, r.wOrg)).astype(np.uint8)
    assert np.array_equal(u.undistort_host(raw), r.undistort(raw))


def test_host_mirror_matches_golden():
    g = np.load(GOLD); u = _mirror(str(g["text"]))
    assert np.array_equal(np.array([u.K[0, 0], u.K[1, 1], u.K[0, 2], u.K[1, 2]]), g["K4d"])
    assert np.array_equal(u.remapX, g["remapX"]) and np.array_equal(u.remapY, g["remapY"])
    assert np.array_equal(u.undistort_host(g["raw"]), g["image"])
    for name in ("kitti00", "kitti360"):                                                             # full-size tables pinned by K + checksums
        u = _mirror(CALIBS[name])
        assert np.array_equal(np.array([u.K[0, 0], u.K[1, 1], u.K[0, 2], u.K[1, 2]]), g[name + "_K4d"])
        assert u.remapX.astype(np.float64).sum() == g[name + "_sum"][0] and u.remapY.astype(np.float64).sum() == g[name + "_sum"][1]


def test_mirror_rejects_what_it_does_not_cover():
    from sdv_loam_b200 import undistort
    with pytest.raises(NotImplementedError):
        undistort.Undistort.from_text("RadTan 0.5 0.8 0.5 0.5 0.1 0.0 0.0 0.0\n640 480\ncrop\n640 480\n")
    with pytest.raises(NotImplementedError):
        undistort.Undistort.from_text("Pinhole 500 500 320 240 0\n640 480\nfull\n640 480\n")


def _check_levels(api, ctx, fid, image, L):
    f = orc.Frame(image, L); h = image.shape[0]
    for l in range(L):
        dI, ab = ctx.frameLevel(fid, l); oI = f.dI(l); oa = f.absSquaredGrad(l); inner = slice(1, (h >> l) - 1)
        assert np.array_equal(dI[..., 0], oI[..., 0]), l
        assert np.array_equal(dI[inner], oI[inner]) and np.array_equal(ab[inner], oa[inner]), l


@pytest.mark.gpu
def test_raw_ingest_matches_golden():
    from sdv_loam_b200 import api
    g = np.load(GOLD); u = _mirror(str(g["text"])); L = api.pyr_levels(u.w, u.h)
    ctx = api.Context(u.K4, u.w, u.h, max_frames=4)
    with pytest.raises(api.SdvError):
        ctx.makeImagesBatch([1], [np.ascontiguousarray(g["raw"]).ctypes.data], raw=True)             # no tables yet
    ctx.setUndistort(u); ctx.makeImagesRaw(1, g["raw"])
    _check_levels(api, ctx, 1, g["image"], L)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["kitti00", "kitti360", "carla"])
def test_raw_ingest_bit_exact_full_size(name):
    """BASELINE sizes: 1241x376 -> 1200x360 (KITTI), 1408x376 -> 1400x360 (KITTI-360), 1392x1024 (CARLA); batch of separate and adjacent host buffers"""
    from sdv_loam_b200 import api
    u = _mirror(CALIBS[name]); L = api.pyr_levels(u.w, u.h); rng = np.random.default_rng(11)
    raws = np.ascontiguousarray(rng.integers(0, 256, (3, u.hOrg, u.wOrg)).astype(np.uint8))
    yy, xx = np.mgrid[0:u.hOrg, 0:u.wOrg]; raws[2] = ((xx * 3 + yy * 5) % 256).astype(np.uint8)    # smooth ramp: interpolation weights matter everywhere
    ctx = api.Context(u.K4, u.w, u.h, max_frames=6); ctx.setUndistort(u)
    ctx.makeImagesBatch([10, 11, 12], [raws[i].ctypes.data for i in range(3)], raw=True); ctx.sync()
    want = [ref.Undistort(CALIBS[name]).undistort(raws[i]) if ref.available() else u.undistort_host(raws[i]) for i in range(3)]
    for i in range(3):
        _check_levels(api, ctx, 10 + i, want[i], L)
    ctx.close()


@pytest.mark.gpu
def test_raw_ingest_photometric_branch():
    """response function G and inverse vignette (PhotometricUndistorter::processFrame :193-203): applied for exposure > 0, skipped (factor path) for exposure <= 0"""
    from sdv_loam_b200 import api
    u = _mirror(SMALL); rng = np.random.default_rng(5)
    u.G = (255.0 * (np.arange(256) / 255.0) ** 1.8).astype(np.float32); u.vignetteMapInv = rng.uniform(0.8, 1.6, (u.hOrg, u.wOrg)).astype(np.float32)
    raw = rng.integers(0, 256, (u.hOrg, u.wOrg)).astype(np.uint8)
    ctx = api.Context(u.K4, u.w, u.h, max_frames=4); ctx.setUndistort(u)
    ctx.makeImagesBatch([1, 2], [raw.ctypes.data, raw.ctypes.data], exposures=[0.02, 0.0], raw=True); ctx.sync()
    src = (u.G[raw] * u.vignetteMapInv).astype(np.float32)

    def remap(src):
        xx = u.remapX.reshape(-1); yy = u.remapY.reshape(-1); xi = xx.astype(np.int32); yi = yy.astype(np.int32); s = src.reshape(-1)
        fx = xx - xi.astype(np.float32); fy = yy - yi.astype(np.float32); xy = fx * fy; o = xi + yi * u.wOrg
        return (xy * s[o + 1 + u.wOrg] + (fy - xy) * s[o + u.wOrg] + (fx - xy) * s[o + 1] + (np.float32(1) - fx - fy + xy) * s[o]).reshape(u.h, u.w)
    assert (u.remapX >= 0).all()
    L = api.pyr_levels(u.w, u.h)
    _check_levels(api, ctx, 1, remap(src), L); _check_levels(api, ctx, 2, remap(raw.astype(np.float32)), L)
    bad = _mirror(SMALL); bad.remapY = bad.remapY.copy(); bad.remapY[0, 0] = bad.hOrg - 0.5        # Undistort.cpp:871 lets such an entry through (iy tested against wOrg)
    with pytest.raises(api.SdvError):
        ctx.setUndistort(bad)
    ctx.close()
