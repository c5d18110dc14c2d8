"""Sequence-level parity on S-KITTI-200 (BASELINE.json config #1: "first 200 frames, single sequence, CPU reference path — pose + energy dump for comparison").

Reference arm = the reference's own FullSystem (oracle/_ref) running the whole pipeline frame by frame; see tests/seq_replay.py.
  * test_reference_pipeline_tracks_the_sequence  (CPU)  the unmodified pipeline runs 200 frames of the synthetic drive and stays on the ground truth
  * test_oracle_replays_reference_sequence       (CPU)  orc.track_new_coarse from the reference's per-frame state == the reference's own result (pins a4, a10, a11)
  * test_gpu_replays_reference_sequence          (GPU)  the product through the C-ABI, same comparison, north_star tolerances: pose 1e-3 m / 1e-3 rad, energy 1e-4 rel
"""
import os
import numpy as np
import pytest
import ref
from conftest import cached_sequence
import seq_replay as sr

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference absent")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "s_kitti_200_reference_dump.npz")


def _seq():
    from sdv_loam_b200 import synth
    return cached_sequence(200, 1000, synth.KITTI_K, synth.KITTI_WH, step=0.5), synth


def _gt(seq, synth, i):
    return np.concatenate([synth._quat_from_R(seq.R[0].T @ seq.R[i]), seq.R[0].T @ (seq.t[i] - seq.t[0])])


def test_reference_pipeline_tracks_the_sequence():
    seq, synth = _seq()
    L = ref.set_calib(seq.wh[0], seq.wh[1], seq.K)
    run = sr.ReferenceRun(seq, L)
    errs = []; kfs = 0
    for i in range(200):
        _, _, res = run.step()
        assert res["rc"] == 0, f"reference pipeline lost track at frame {i}"
        errs.append(sr.pose_err(res["camToWorld"], _gt(seq, synth, i))); kfs += res["isKeyframe"]
    et, er = np.array(errs).T
    assert et.max() < 0.5 and er.max() < 0.02, (et.max(), er.max())        # < 0.5 % drift over the 100 m drive
    assert 30 <= run.S.num_keyframes() <= 150


@pytest.mark.parametrize("n_frames", [200])
def test_oracle_replays_reference_sequence(n_frames):
    """Every frame: the restatement, started from the reference's state, reproduces the reference's trackNewCoarse (hypotheses, LM, reprojectMap, structPoseEstimation)."""
    import orc
    seq, synth = _seq()
    L = ref.set_calib(seq.wh[0], seq.wh[1], seq.K)
    run = sr.ReferenceRun(seq, L); cache = {}
    worst = [0.0, 0.0, 0.0]; exact = 0; n = 0
    for i in range(n_frames):
        snap, order, res = run.step()
        assert res["rc"] == 0
        if snap is None:
            continue
        r = sr.replay_orc(seq, snap, order, i, L, seq.K, cache)
        et, er = sr.pose_err(r["camToWorld"], res["tracked_camToWorld"])
        ee = abs(r["lastCoarseRMSE"][0] - res["lastCoarseRMSE"][0]) / res["lastCoarseRMSE"][0]
        worst = [max(worst[0], et), max(worst[1], er), max(worst[2], ee)]; n += 1
        exact += int(et < 1e-12 and er < 1e-12 and ee == 0.0)
        # The reference run itself is not bit-reproducible (it reads uninitialised heap in a few places — e.g. CoarseInitializer.cpp:864, gradient rows 0/h-1 of dIp,
        # the freed EFFrame of EnergyFunctional.cpp:428 — found with AddressSanitizer): an occasional frame differs in the 7th digit.  Everything else is identical.
        assert et < 1e-4 and er < 1e-5 and ee < 1e-5, (i, et, er, ee, r["tries"])
        assert np.allclose(r["aff_g2l"], res["aff_g2l"], atol=1e-9)
    assert n >= n_frames - 3 and exact >= 0.9 * n
    print(f"orc vs reference over {n} frames: max pose err {worst[0]:.2e} m {worst[1]:.2e} rad, energy {worst[2]:.2e} rel, bit-identical poses {exact}/{n}")


@pytest.mark.gpu
def test_gpu_replays_reference_sequence():
    """200 frames, teacher-forced from the reference's state: pose within 1e-3 m / 1e-3 rad and photometric energy (achievedRes[0]) within 1e-4 relative (north_star)
    on every frame whose discrete decisions agree.  trackNewCoarse is full of them — the LM accepts/rejects steps on an energy comparison, the Reprojector keeps/drops
    candidate matches by thresholds, structPoseEstimation accepts/rejects damped steps — so a last-bit difference (float sums in another order than the SSE code) can
    flip one decision on a frame that sits on a boundary: one more LM iteration moves the energy in the 4th digit, one match more moves the refined pose by millimetres
    along the weakly constrained direction.  The reference run itself is not bit-reproducible (uninitialised reads, see test_oracle_replays_reference_sequence), so WHICH
    frames sit on a boundary changes from run to run (observed: none, or one in 197).  Such frames are listed with the oracle's own numbers next to the GPU's (same
    inputs), must stay under 1 cm / 1e-3 rad / 1e-3 relative energy, and may not exceed 3 % of the sequence."""
    import orc
    seq, synth = _seq()
    L = ref.set_calib(seq.wh[0], seq.wh[1], seq.K)
    run = sr.ReferenceRun(seq, L); gpu = sr.GpuReplay(seq, L, seq.K); cache = {}
    dump = []; worst = [0.0, 0.0, 0.0]; n = 0; flipped = []
    for i in range(200):
        snap, order, res = run.step()
        assert res["rc"] == 0
        dump.append(sr.dump_line(res))
        if snap is None:
            continue
        assert np.array_equal(np.float32(snap["tracker_K"]), np.float32(snap["calib"]))  # tracker K (makeK) == current CalibHessian on every frame: one context calibration (sdv_set_calib) serves both
        g = gpu.track(snap, order, i)
        et, er = sr.pose_err(g["camToWorld"], res["tracked_camToWorld"])
        ee = abs(g["lastCoarseRMSE"][0] - res["lastCoarseRMSE"][0]) / res["lastCoarseRMSE"][0]
        n += 1
        assert g["tries"] >= 1
        if et < 1e-3 and er < 1e-3 and ee < 1e-4:
            worst = [max(worst[0], et), max(worst[1], er), max(worst[2], ee)]
            continue
        o = sr.replay_orc(seq, snap, order, i, L, seq.K, cache)                           # same call on the oracle: which stage moved?
        flipped.append((i, et, er, ee, g["n_matches"], o["n_matches"], (g["refine_iterations"], g["refine_accepts"]), (o["refine_iterations"], o["refine_accepts"])))
        assert et < 1e-2 and er < 1e-3 and ee < 1e-3, flipped[-1]
    assert n == 197
    print(f"GPU vs reference CPU path over {n} frames: {n - len(flipped)} frames within 1e-3 m / 1e-3 rad (max {worst[0]:.2e} m {worst[1]:.2e} rad), energy max {worst[2]:.2e} rel")
    for f in flipped:
        print("  frame %d: %.2e m %.2e rad %.2e energy; matches gpu/orc %d/%d; refine (its,acc) gpu %s orc %s" % f)
    assert len(flipped) <= 0.03 * n
    if os.path.exists(GOLD):                                                # the committed dump of the reference run (drift pin of the reference arm itself; it has known run-to-run jitter)
        g = np.load(GOLD)["dump"]; d = np.array(dump)
        assert g.shape == d.shape and np.abs(g[:, 4:7] - d[:, 4:7]).max() < 0.05
