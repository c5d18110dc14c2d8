"""Sequence-level parity harness (BASELINE.json config #1).

The reference's OWN pipeline (oracle/_ref: FullSystem::addActiveFrame per frame — trackNewCoarse, makeKeyFrame, optimize, marginalisation, new traces, ...) runs a
synthetic KITTI-shape sequence.  Before every frame the state FullSystem::trackNewCoarse is about to see is read back (reference clouds of the tracker in use, pose
history, active map, the Reprojector's cell order) and the SAME call is replayed on another arm:
  * replay_orc : the oracle restatement (orc.track_new_coarse)                     — CPU, pins a4/a10/a11 of the restatement on the reference at sequence level
  * replay_gpu : the product (sdv_track_new_coarse_batch through the C-ABI)        — the north_star parity statement: pose 1e-3 m / 1e-3 rad, energy 1e-4 relative
Teacher-forced: every frame starts from the reference's state, so one flipped accept cannot hide behind a diverged trajectory — it shows up at that frame.
Test infrastructure only (imports oracle/).
"""
from __future__ import annotations
import numpy as np

MARGIN = 20          # LiDAR pixels closer than this to the image border are not fed: CoarseInitializer.cpp:864 reads the 8-pixel pattern around every pixel of every level
                     # without a bounds check (heap over-read near the border in the reference)


def frame_cloud(seq, i):
    w, h = seq.wh; cl = seq.clouds[i]
    return cl[(cl[:, 0] > MARGIN) & (cl[:, 0] < w - MARGIN) & (cl[:, 1] > MARGIN) & (cl[:, 1] < h - MARGIN)]


def n_cells(wh, cell=25):
    return int(np.ceil(wh[0] / float(cell))) * int(np.ceil(wh[1] / float(cell)))


class ReferenceRun:
    """Drives ref.System over `seq` and yields, per tracked frame, (snapshot-before, cell_order, result-after)."""

    def __init__(self, seq, levels, seed_base=1000):
        import ref
        self.ref = ref; self.seq = seq; self.S = ref.System(levels, perfect_images=True); self.seed_base = seed_base; self.i = 0

    def step(self):
        ref, S, seq, i = self.ref, self.S, self.seq, self.i
        snap = S.tracker_snapshot() if i >= 3 else None     # frames 0-2: initialiser bootstrap (first-frame reference, two-frame window), outside the running-system branch
        S.srand(self.seed_base + i); order = ref.libc_rand_shuffle(n_cells(seq.wh)); S.srand(self.seed_base + i)   # the Reprojector of trackNewCoarse is the first rand() user of a frame
        rc = S.addActiveFrame(seq.images[i], frame_cloud(seq, i), 0.1 * i)
        res = S.frame(i); res["rc"] = rc; res["lastCoarseRMSE"] = S.lastCoarseRMSE(); res["index"] = i
        # what trackNewCoarse produced: shell->camToWorld is overwritten by the bundle adjustment when the frame becomes a keyframe (makeKeyFrame), camToTrackingRef is not
        import orc
        res["tracked_camToWorld"] = orc.se3_mul(snap["lastF"], res["camToTrackingRef"]) if snap is not None else res["camToWorld"]
        self.i += 1
        return snap, order, res


def map_struct(snap, dtype):
    p5 = snap["map_pts"]; m = np.zeros(len(p5), dtype)
    m["u"], m["v"], m["idepth"], m["host"], m["type"] = p5[:, 0], p5[:, 1], p5[:, 2], p5[:, 3].astype(np.int32), p5[:, 4].astype(np.int32)
    return m


def replay_orc(seq, snap, order, i, levels, K, frames_cache):
    """FullSystem::trackNewCoarse for frame i on the oracle restatement, from the reference's state `snap`."""
    import orc
    w, h = seq.wh
    def F(k):
        if k not in frames_cache:
            frames_cache[k] = orc.Frame(seq.images[k], levels)
        return frames_cache[k]
    # the bundle adjustment optimises the intrinsics too: each tracker carries the K of its keyframe (CoarseTracker::makeK), the Reprojector reads the current CalibHessian
    Kt = tuple(float(x) for x in snap["tracker_K"]); Kc = tuple(float(np.float32(x)) for x in snap["calib"])
    tr = orc.CoarseTracker(w, h, levels, Kt); tr.settings(6.0, 20.0, -1.0, -1.0)                    # mode 2 of main.cpp:461-468 (perfect images: affine fixed)
    for l in range(levels):
        u, v, idp, col = snap["clouds"][l]; tr.setCloud(F(snap["ref_frame"]), l, u, v, idp, col, *snap["ref_ab"])
    m = map_struct(snap, orc_map_dtype())
    r = orc.track_new_coarse(tr, F(i), Kc, [F(int(k)) for k in snap["kf_ids"]], snap["kf_T7"], snap["kf_ab"], m, snap["sprelast"], snap["slast"], snap["lastF"], snap["aff_last"],
                             True, snap["lastCoarseRMSE"], cell_order=order, max_matches=400)
    for k in [k for k in frames_cache if k < i - 40]:
        del frames_cache[k]
    return r


def orc_map_dtype():
    return np.dtype([("u", np.float32), ("v", np.float32), ("idepth", np.float32), ("host", np.int32), ("type", np.int32)])


class GpuReplay:
    """The same call through the product's C-ABI (sdv_frame_upload, sdv_tracker_set_cloud, sdv_map_set, sdv_track_new_coarse_batch)."""

    def __init__(self, seq, levels, K):
        from sdv_loam_b200 import api
        self.api = api; self.seq = seq; w, h = seq.wh
        self.ctx = api.Context(K, w, h, levels=levels, n_tracker_slots=2, max_frames=24, affineOptModeA=-1.0, affineOptModeB=-1.0)
        self.tr = api.CoarseTracker(self.ctx, 0); self.rp = api.Reprojector(self.ctx); self.resident = set(); self.K = None

    def _need(self, ids):
        for k in ids:
            if k not in self.resident:
                self.ctx.makeImages(int(k), self.seq.images[int(k)]); self.resident.add(int(k))

    def track(self, snap, order, i):
        api = self.api; kf = [int(k) for k in snap["kf_ids"]]; need = set(kf) | {int(snap["ref_frame"]), i}
        Kt = tuple(float(x) for x in snap["tracker_K"])                                              # CoarseTracker::makeK of the tracker in use (== float(CalibHessian) on every frame)
        if Kt != self.K:
            self.ctx.setCalib(Kt); self.K = Kt
        for k in list(self.resident - need):
            try:
                self.ctx.releaseFrame(k); self.resident.discard(k)
            except api.SdvError:
                pass                                                                                 # still referenced by the resident map of the previous frame: released after setMap below
        self._need(sorted(need))
        for l in range(self.ctx.levels):
            u, v, idp, col = snap["clouds"][l]; self.tr.setCloud(int(snap["ref_frame"]), l, u, v, idp, col, *snap["ref_ab"])
        self.rp.setMap(0, kf, snap["kf_T7"], snap["kf_ab"], map_struct(snap, api.MAP_PT_DTYPE))
        for k in list(self.resident - need):
            self.ctx.releaseFrame(k); self.resident.discard(k)
        job = dict(slot=0, frame=i, sprelast_c2w=snap["sprelast"], slast_c2w=snap["slast"], lastF_c2w=snap["lastF"], aff_last=snap["aff_last"], poses_valid=1, lastCoarseRMSE=snap["lastCoarseRMSE"])
        return api.trackNewCoarseBatch(self.ctx, [job], cell_order=order, max_matches=400)[0]


def pose_err(Ta, Tb):
    """(translation m, rotation rad) between two camToWorld poses T7 = {qw,qx,qy,qz,tx,ty,tz}"""
    import orc
    d = orc.se3_log(orc.se3_mul(orc.se3_inv(np.asarray(Ta, np.float64)), np.asarray(Tb, np.float64)))
    return float(np.linalg.norm(d[:3])), float(np.linalg.norm(d[3:]))


def dump_line(res):
    """the fields of the reference's coarseTrackingLog line (FullSystem.cpp:502-513) that exist without ROS time: id, camToWorld.log(), a, b, achievedRes[0]"""
    import orc
    return np.concatenate([[res["index"]], orc.se3_log(res["camToWorld"]), res["aff_g2l"], [res["lastCoarseRMSE"][0]]])
