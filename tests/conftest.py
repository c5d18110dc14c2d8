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
        seq.images = list(d["images"]); seq.clouds = list(d["clouds"])
        return seq
    seq = synth.Sequence(n, seed=seed, K=K, wh=wh, **kw)
    clouds = np.empty(n, dtype=object)
    for i in range(n):
        clouds[i] = seq.clouds[i]
    tmp = path + ".%d.tmp.npz" % os.getpid()                                # atomic publish: several ranks of a multi-GPU bench may render the same sequence at once
    np.savez_compressed(tmp, R=seq.R, t=seq.t, images=np.stack(seq.images), clouds=clouds)
    os.replace(tmp, path)
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
