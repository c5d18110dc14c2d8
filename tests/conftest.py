import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu() -> bool:
    try:
        import ctypes
        cu = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        return cu.cuInit(0) == 0 and cu.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


_CACHE = os.path.join(ROOT, "tests", ".cache")


def cached_sequence(n, seed, K, wh, **kw):
    """Synthetic sequences are deterministic but slow to render; cache them on disk (git-ignored)."""
    import sdv_loam_b200  # noqa: F401
    from sdv_loam_b200 import synth
    os.makedirs(_CACHE, exist_ok=True)
    key = f"seq_n{n}_s{seed}_{wh[0]}x{wh[1]}_" + "_".join(f"{k}{v}" for k, v in sorted(kw.items())) + ".npz"
    path = os.path.join(_CACHE, key)
    if os.path.exists(path):
        d = np.load(path, allow_pickle=True)
        seq = synth.Sequence.__new__(synth.Sequence)
        seq.K, seq.wh, seq.n, seq.seed = K, wh, n, seed
        seq.R, seq.t = d["R"], d["t"]
        seq.images = [im.astype(np.float32) for im in d["images"]]; seq.clouds = list(d["clouds"])     # stored mono8 when the frames are mono8-exact (they are: synth.render rounds)
        return seq
    seq = render_sequence(n, seed, K, wh, **kw)
    clouds = np.empty(n, dtype=object)
    for i in range(n):
        clouds[i] = seq.clouds[i]
    imgs = np.stack(seq.images); u8 = imgs.astype(np.uint8)
    tmp = path + ".%d.tmp.npz" % os.getpid()                                # atomic publish: several ranks of a multi-GPU bench may render the same sequence at once
    np.savez_compressed(tmp, R=seq.R, t=seq.t, images=(u8 if np.array_equal(u8.astype(np.float32), imgs) else imgs), clouds=clouds)
    os.replace(tmp, path)
    return seq


def _render_one(args):
    from sdv_loam_b200 import synth
    seed, K, wh, R, t, gain, bias, noise, i, beams = args
    world = synth.World(seed)
    img, _ = synth.render(world, R, t, K, wh, gain=gain, bias=bias, noise=noise, seed=seed * 1000 + i)
    return img, synth.lidar_pixels(world, R, t, K, wh, beams=beams)


def render_sequence(n, seed, K, wh, beams=64, step=1.0, gain_jitter=0.0, bias_jitter=0.0, noise=0.0):
    """synth.Sequence(...) with the frames rendered by a process pool (3 s per KITTI-size frame on one core): identical output, frame by frame."""
    import sdv_loam_b200  # noqa: F401
    from sdv_loam_b200 import synth
    if n * wh[0] * wh[1] < 6 * 1200 * 360:                                # a few KITTI-size frames: not worth a pool
        return synth.Sequence(n, seed=seed, K=K, wh=wh, beams=beams, step=step, gain_jitter=gain_jitter, bias_jitter=bias_jitter, noise=noise)
    import multiprocessing as mp
    seq = synth.Sequence.__new__(synth.Sequence)
    seq.K, seq.wh, seq.n, seq.seed = K, wh, n, seed
    seq.world = synth.World(seed); seq.R, seq.t = synth.trajectory(n, seed, step)
    rng = np.random.default_rng(seed + 13); jobs = []
    for i in range(n):
        g = 1.0 + (rng.normal(0, gain_jitter) if gain_jitter > 0 else 0.0); b = rng.normal(0, bias_jitter) if bias_jitter > 0 else 0.0
        jobs.append((seed, K, wh, seq.R[i], seq.t[i], g, b, noise, i, beams))
    with mp.get_context("fork").Pool(min(os.cpu_count() or 1, 16)) as pool:
        out = pool.map(_render_one, jobs, chunksize=2)
    seq.images = [o[0] for o in out]; seq.clouds = [o[1] for o in out]
    return seq


SMALL_WH = (640, 192)
SMALL_K = (383.4, 383.4, 312.0, 97.0)


@pytest.fixture(scope="session")
def small_seq():
    return cached_sequence(3, 1000, SMALL_K, SMALL_WH)


@pytest.fixture(scope="session")
def kitti_seq():
    from sdv_loam_b200 import synth
    return cached_sequence(3, 1000, synth.KITTI_K, synth.KITTI_WH)
