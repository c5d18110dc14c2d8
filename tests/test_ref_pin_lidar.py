"""Pins the LiDAR front-end restatement (oracle/orc_lidar.cpp: projectPointCloud, groundRemoval, cloudSegmentation and the pixel projection of lidarCloudHandler —
SURVEY §8f rank 3, second half) on the reference's own src/main.cpp, compiled unmodified into oracle/_ref against PCL / ROS stand-ins (oracle/ref_stub/pcl*).
Range image, ground image, every label (the BFS numbering included), the pixel rows {Ku, Kv, depth}, the running pixel box and addFeaturePoint: BIT FOR BIT."""
import numpy as np
import pytest
import orc
import ref
from conftest import SMALL_K, SMALL_WH

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference absent")


def sweeps(n, seed0=0):
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import synth
    world = synth.World(3000); R, t = synth.trajectory(n, 3000)
    return synth, [synth.lidar_sweep(world, R[k], t[k], seed=seed0 + k) for k in range(n)]


def test_lidar_handler_bit_exact():
    from test_ref_pin_ba import _window
    win, ob, rb, _keep = _window((0, 1), 5)                               # a FullSystem to stand for the node's global `fullSystem`; calibration = SMALL
    synth, S = sweeps(3); fe = orc.LidarFrontEnd()
    lr_o = np.array([10000, -1, 10000, -1], np.int32); lr_r = lr_o.copy()
    for k, sw in enumerate(S):
        tlc = synth.TLC if k < 2 else np.array([0.0, -0.08, 0.35])           # k == 2: camera in FRONT of the sensor origin -> the empty ground cells (points at the origin) project too
        o = fe.handle(sw, synth.RLC, tlc, SMALL_K, SMALL_WH, lr_o, images=True); r = ref.lidar_handler(rb, sw, synth.RLC, tlc, SMALL_K, lr_r, images=True)
        assert np.array_equal(o["range"], r["range"]) and np.array_equal(o["ground"], r["ground"]) and np.array_equal(o["label"], r["label"]), k
        assert o["n_segmented"] == r["n_segmented"] and np.array_equal(o["cloud_px"], r["cloud_px"]) and len(o["cloud_px"]) > 3000, k
        assert np.array_equal(o["lrud"], r["lrud"]) and o["addFeaturePoint"] == r["addFeaturePoint"], k
        lr_o, lr_r = o["lrud"], r["lrud"]                                    # the pixel box keeps growing across sweeps (FullSystem members)
        lab = o["label"]; feasible = np.unique(lab[(lab > 0) & (lab != 999999)])
        assert len(feasible) > 20 and (lab == 999999).sum() > 100 and (o["ground"] == 1).sum() > 1000 and (o["range"] == np.finfo(np.float32).max).sum() > 500
    assert o["addFeaturePoint"] in (0, 1)


def test_sparse_and_degenerate_sweeps():
    """empty sweep, a sweep of only NaN / near returns, a single ring: the handler must agree there too"""
    from test_ref_pin_ba import _window
    win, ob, rb, _keep = _window((0, 1), 5)
    synth, S = sweeps(1, seed0=7); fe = orc.LidarFrontEnd(); sw = S[0]
    ring = sw[np.abs(np.degrees(np.arctan2(sw[:, 2], np.hypot(sw[:, 0], sw[:, 1]))) + 10.0) < 0.3]
    for cloud in (np.zeros((0, 4), np.float32), np.array([[np.nan, 0, 0, 0], [0.01, 0.01, 0, 0]], np.float32), ring, sw[::7]):
        lr = [10000, -1, 10000, -1]
        o = fe.handle(cloud, synth.RLC, synth.TLC, SMALL_K, SMALL_WH, lr, images=True); r = ref.lidar_handler(rb, cloud, synth.RLC, synth.TLC, SMALL_K, lr, images=True)
        assert np.array_equal(o["range"], r["range"]) and np.array_equal(o["ground"], r["ground"]) and np.array_equal(o["label"], r["label"])
        assert np.array_equal(o["cloud_px"], r["cloud_px"]) and np.array_equal(o["lrud"], r["lrud"]) and o["addFeaturePoint"] == r["addFeaturePoint"]
