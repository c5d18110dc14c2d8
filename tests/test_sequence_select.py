"""Sequence-level parity of the keyframe-rate candidate management on the reference's OWN running pipeline: the reference's FullSystem (oracle/_ref) runs
the synthetic KITTI drive frame by frame (addActiveFrame -> trackNewCoarse -> makeKeyFrame -> ... -> makeNewTraces); at every new keyframe the immature points it created
(FullSystem::makeNewTraces: makeHists, makeMapsFromLidar, makeMaps, Shi-Tomasi typing, occupancy mask) are compared with the restatement / the CUDA path started from the
SAME selector state — and the selector state (currentPotential, the persistent monocular map) is carried along on both sides, never re-synchronised, so one wrong decision
anywhere would show up on every later keyframe.
  * test_oracle_follows_reference_keyframes  (CPU)  orc.Selector.makeNewTraces == the reference's immature points, every keyframe, bit for bit
  * test_gpu_follows_reference_keyframes     (GPU)  sdv_make_new_traces_batch through the C-ABI, same comparison
The reference reads the never-written first / last rows of absSquaredGrad[] (PixelSelector2.cpp:306) and the tail of thsSmoothed: glibc's M_PERTURB makes every malloc'ed
block read as zero for the duration of the test (what the restatement and the device define those reads to be)."""
import ctypes as C
import numpy as np
import pytest
import orc
import ref
from conftest import cached_sequence
import seq_replay as sr

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference absent")
N_FRAMES = 70
DENSITY = 500.0                                                             # setting_desiredImmatureDensity, util/settings.cpp:46 (ref_sys_create keeps the default)


def _drive(on_keyframe):
    """runs the reference pipeline; calls on_keyframe(i, cloud, lrud, add, rows) for every frame that became a keyframe after the bootstrap; returns the number of keyframes seen"""
    from sdv_loam_b200 import synth
    seq = cached_sequence(200, 1000, synth.KITTI_K, synth.KITTI_WH, step=0.5); w, h = seq.wh
    libc = C.CDLL(None); libc.mallopt(-6, 0xFF)                              # M_PERTURB: malloc'ed memory reads as 0x00
    try:
        L = ref.set_calib(w, h, seq.K); run = sr.ReferenceRun(seq, L); S = run.S; S.selector_zero(); S.set_selection_map(np.zeros((h, w), np.float32))
        lrud = np.array([10000, -1, 10000, -1], np.int32); seen = 0; last_kf = -1
        for i in range(N_FRAMES):
            cloud = sr.frame_cloud(seq, i); ku, kv = cloud[:, 0].astype(np.float32), cloud[:, 1].astype(np.float32)     # the running pixel box of main.cpp:834-837
            lrud = np.array([min(lrud[0], int(ku.min())), max(lrud[1], int(ku.max())), min(lrud[2], int(kv.min())), max(lrud[3], int(kv.max()))], np.int32)
            add = int(i % 3 != 1)                                            # both values of FullSystem::addFeaturePoint occur, incl. the stale-map walk
            S.set_lidar_state(lrud, add); pot_before, map_before = S.selector_state((w, h))
            # in this reference makeNewTraces runs BEFORE activatePointsMT (FullSystem.cpp:1080 / :1102): its complete output cannot be read off the keyframe afterwards, so it
            # is taken by running the reference's own function on a probe frame with the system's live selector state (restored afterwards) right before the frame is added
            rows = S.probe_new_traces(ref.Frame(seq.images[i], seq.wh, L), cloud) if i >= 3 else None
            _, _, res = run.step(); assert res["rc"] == 0, i
            _, kf = S.newest_kf_immature()
            if kf != i or kf == last_kf: continue                            # not a keyframe: the selector state did not move
            last_kf = kf
            on_keyframe(i, cloud, lrud, add, rows, pot_before, map_before, S, seq, L); seen += rows is not None
        return seen
    finally:
        libc.mallopt(-6, 0)


def _check_rows(T, rows, tag):
    assert len(T) == len(rows), (tag, len(T), len(rows))
    for f, col in (("u", 0), ("v", 1), ("my_type", 2), ("score", 3), ("idepth_fromSensor", 4)):
        assert np.array_equal(T[f], rows[:, col]), (tag, f)
    assert np.array_equal(T["isFromSensor"], rows[:, 5].astype(np.int32)) and np.array_equal(T["type"], rows[:, 6].astype(np.int32)), tag


def test_oracle_follows_reference_keyframes():
    state = {}
    def on_kf(i, cloud, lrud, add, rows, pot_before, map_before, S, seq, L):
        w, h = seq.wh
        if "sel" not in state: state["sel"] = orc.Selector(w, h, orc.libc_random_pattern(w, h)); state["map"] = np.zeros((h, w), np.float32); state["sel"].currentPotential = pot_before
        osel, omap = state["sel"], state["map"]
        assert osel.currentPotential == pot_before and np.array_equal(omap, map_before), i           # carried state == the reference's state before the call
        T, num, _ = osel.makeNewTraces(orc.Frame(seq.images[i], L), cloud, orc.lidar_density(lrud, seq.wh, DENSITY), DENSITY, add, omap)
        pot_after, map_after = S.selector_state((w, h))
        assert osel.currentPotential == pot_after and np.array_equal(omap, map_after), i
        if rows is not None: _check_rows(T, rows, i); state["n"] = state.get("n", 0) + len(T); state["mono"] = state.get("mono", 0) + int((T["isFromSensor"] == 0).sum())
    seen = _drive(on_kf)
    assert seen >= 8 and state["n"] > 3000 and state["mono"] > 300, (seen, state.get("n"), state.get("mono"))
    print(f"makeNewTraces: {seen} keyframes of the reference run reproduced, {state['n']} immature points ({state['mono']} monocular)")


@pytest.mark.gpu
def test_gpu_follows_reference_keyframes():
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import api
    state = {}
    def on_kf(i, cloud, lrud, add, rows, pot_before, map_before, S, seq, L):
        w, h = seq.wh
        if "ps" not in state:
            state["ctx"] = api.Context(seq.K, w, h, max_frames=4); state["ps"] = api.PixelSelector(state["ctx"], 1, api.random_pattern(w, h)); state["ps"].potential(0, pot_before)
        ctx, ps = state["ctx"], state["ps"]
        assert ps.potential(0) == pot_before and np.array_equal(ps.selectionMap(0), map_before.astype(np.uint8)), i
        ctx.makeImages(i, seq.images[i])
        (T, I), num = ps.makeNewTracesBatch([0], [i], [cloud], api.lidar_density(lrud, seq.wh, DENSITY), DENSITY, add, cap=1 << 15)[0][0], None
        ctx.releaseFrame(i)
        pot_after, map_after = S.selector_state((w, h))
        assert ps.potential(0) == pot_after and np.array_equal(ps.selectionMap(0), map_after.astype(np.uint8)), i
        if rows is not None: _check_rows(T, rows, i); state["n"] = state.get("n", 0) + len(T)
    seen = _drive(on_kf)
    assert seen >= 8 and state["n"] > 3000
    state["ctx"].close()
