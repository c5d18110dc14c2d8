"""Sequence-level parity of the sliding-window back-end (rows b1-b8) on LIVE windows of the reference's own running pipeline: the reference's FullSystem (oracle/_ref) runs the
synthetic KITTI drive (mode 1: affine brightness free); at three points of the run its current window — seven keyframes, ~1 200 active points, ~3 800 residuals with the
matchers backprojectMap gave them, the marginalisation prior (HM, bM) that real marginalisations accumulated, intrinsics that earlier bundle adjustments moved away from their
linearisation point — is flattened (EnergyFunctional order) and FullSystem::optimize is run on it by the reference itself and by the restatement.  Same final energy (rmse)
to the last printed digit of a float, inverse depths / frame states / intrinsics within 1e-6 (the windows pass through LDLT and the nullspace SVD, which the pin build takes
from stand-ins — DESIGN.md §0 — and the two sides walk a point's residuals in different container orders)."""
import ctypes as C
import numpy as np
import pytest
import orc
import ref
from conftest import cached_sequence
import seq_replay as sr

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference absent")


def test_oracle_optimizes_live_windows_like_the_reference():
    from sdv_loam_b200 import synth
    seq = cached_sequence(200, 1000, synth.KITTI_K, synth.KITTI_WH, step=0.5); w, h = seq.wh
    libc = C.CDLL(None); libc.mallopt(-6, 0xFF)                               # zero heap: a pipeline run independent of what earlier runs left behind (see test_sequence_trace.py)
    try:
        L = ref.set_calib(w, h, seq.K); S = ref.System(L, perfect_images=False); lrud = np.array([10000, -1, 10000, -1], np.int32); checked = 0
        for i in range(46):
            cloud = sr.frame_cloud(seq, i); ku, kv = cloud[:, 0].astype(np.float32), cloud[:, 1].astype(np.float32)
            lrud = np.array([min(lrud[0], int(ku.min())), max(lrud[1], int(ku.max())), min(lrud[2], int(kv.min())), max(lrud[3], int(kv.max()))], np.int32); S.set_lidar_state(lrud, 1)
            S.srand(1000 + i); assert S.addActiveFrame(seq.images[i], cloud, 0.1 * i) == 0, i
            if i not in (27, 36, 45): continue
            A = S.export_window((w, h)); assert A["nF"] >= 5 and len(A["uv"]) > 500 and len(A["r_point"]) > 1500 and np.abs(A["HM"]).max() > 1.0, (i, A["nF"], len(A["uv"]))
            assert np.abs(A["K"] / np.array([50, 50, 50, 50]) - A["K_zero"]).max() > 1e-9          # the intrinsics have left their linearisation point
            rmse_ref = S.optimize(6); B = S.export_window((w, h)); assert np.array_equal(A["shell_ids"], B["shell_ids"]) and len(A["uv"]) == len(B["uv"])
            ob = orc.BAWindow(A, [orc.Frame(seq.images[k], L) for k in A["kf_idx"]]); r = ob.optimize(6); po = ob.points(); fo = ob.frames(); co = ob.calib()
            assert abs(r["rmse"] - rmse_ref) <= 1e-6 * rmse_ref, (i, r["rmse"], rmse_ref)
            assert np.abs(po["idepth"] - B["idepth"]).max() < 1e-6 and np.abs(fo["state"] - B["state"]).max() < 1e-6, (i, np.abs(po["idepth"] - B["idepth"]).max(), np.abs(fo["state"] - B["state"]).max())
            assert np.abs(np.asarray(co[0]) * np.array([50, 50, 50, 50]) - B["K"]).max() < 1e-5 * 700, i
            assert np.abs(B["idepth"] - A["idepth"]).max() > 1e-5                                   # the optimisation did move the window
            checked += 1
        assert checked == 3
    finally:
        libc.mallopt(-6, 0)
