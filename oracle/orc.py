"""ctypes binding of the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs — never by the product package.
Parity: pinned on the reference's own compiled code (oracle/_ref, oracle/ref.py; see orc_math.hpp and DESIGN.md §0)."""
from __future__ import annotations
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liborc.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_frame_create.restype = C.c_void_p
        L.orc_frame_create.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_float]
        L.orc_frame_destroy.argtypes = [C.c_void_p]
        L.orc_frame_dI.restype = C.POINTER(C.c_float); L.orc_frame_dI.argtypes = [C.c_void_p, C.c_int]
        L.orc_frame_abs.restype = C.POINTER(C.c_float); L.orc_frame_abs.argtypes = [C.c_void_p, C.c_int]
        L.orc_tracker_create.restype = C.c_void_p
        L.orc_tracker_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
        L.orc_tracker_destroy.argtypes = [C.c_void_p]
        L.orc_tracker_settings.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
        L.orc_tracker_get_K.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.orc_tracker_get_Ki.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.orc_tracker_set_ref.argtypes = [C.c_void_p, C.c_void_p, _f32p, _i32p, C.c_int, C.c_double, C.c_double]
        L.orc_tracker_set_cloud.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, C.c_double, C.c_double]
        L.orc_tracker_cloud_n.argtypes = [C.c_void_p, C.c_int]
        L.orc_tracker_get_cloud.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, _f32p, _f32p]
        L.orc_tracker_calc_res.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _f64p, C.c_double, C.c_double, C.c_float, _f64p]
        L.orc_tracker_warped_n.argtypes = [C.c_void_p]
        L.orc_tracker_get_warped.argtypes = [C.c_void_p, _f32p]
        L.orc_tracker_calc_gs.argtypes = [C.c_void_p, C.c_int, _f64p, C.c_double, C.c_double, _f64p, _f64p]
        L.orc_tracker_track.argtypes = [C.c_void_p, C.c_void_p, _f64p, _f64p, C.c_int, _f64p, _f64p, _f64p, _i64p, _i32p, _i32p]
        for name, n_in in (("orc_se3_exp", 1), ("orc_se3_log", 1), ("orc_se3_inv", 1), ("orc_se3_rot", 1), ("orc_se3_adj", 1), ("orc_se3_mul", 2), ("orc_se3_from_rt", 2)):
            getattr(L, name).argtypes = [_f64p] * (n_in + 1)
        L.orc_ldlt_solve.argtypes = [C.c_int, _f64p, _f64p, _f64p]
        L.orc_aff_from_to.argtypes = [C.c_float, C.c_float, C.c_double, C.c_double, C.c_double, C.c_double, _f64p]
        _LIB = L
    return _LIB


# ------------------------------------------------------------------ SE3 helpers (T7 = {qw,qx,qy,qz,tx,ty,tz})
def se3_exp(a):
    T = np.zeros(7); lib().orc_se3_exp(np.ascontiguousarray(a, np.float64), T); return T

def se3_log(T):
    a = np.zeros(6); lib().orc_se3_log(np.ascontiguousarray(T, np.float64), a); return a

def se3_mul(A, B):
    Cc = np.zeros(7); lib().orc_se3_mul(np.ascontiguousarray(A, np.float64), np.ascontiguousarray(B, np.float64), Cc); return Cc

def se3_inv(A):
    Cc = np.zeros(7); lib().orc_se3_inv(np.ascontiguousarray(A, np.float64), Cc); return Cc

def se3_rot(A):
    R = np.zeros(9); lib().orc_se3_rot(np.ascontiguousarray(A, np.float64), R); return R.reshape(3, 3)

def se3_adj(A):
    M = np.zeros(36); lib().orc_se3_adj(np.ascontiguousarray(A, np.float64), M); return M.reshape(6, 6)

def se3_from_rt(R, t):
    T = np.zeros(7); lib().orc_se3_from_rt(np.ascontiguousarray(R, np.float64).reshape(-1), np.ascontiguousarray(t, np.float64), T); return T

def ldlt_solve(A, b):
    n = len(b); x = np.zeros(n)
    lib().orc_ldlt_solve(n, np.ascontiguousarray(A, np.float64).reshape(-1), np.ascontiguousarray(b, np.float64), x); return x


class Frame:
    def __init__(self, color, levels: int, exposure: float = 1.0):
        color = np.ascontiguousarray(color, np.float32)
        self.h, self.w = color.shape
        self.levels = levels
        self.p = lib().orc_frame_create(color, self.w, self.h, levels, exposure)

    def dI(self, lvl):
        w, h = self.w >> lvl, self.h >> lvl
        return np.ctypeslib.as_array(lib().orc_frame_dI(self.p, lvl), shape=(h, w, 3)).copy()

    def absSquaredGrad(self, lvl):
        w, h = self.w >> lvl, self.h >> lvl
        return np.ctypeslib.as_array(lib().orc_frame_abs(self.p, lvl), shape=(h, w)).copy()

    def __del__(self):
        if getattr(self, "p", None):
            lib().orc_frame_destroy(self.p); self.p = None


class CoarseTracker:
    """Mirror of the reference class (FullSystem/CoarseTracker.h:16-133) over the oracle."""

    def __init__(self, w, h, levels, K):
        self.w, self.h, self.levels = w, h, levels
        self.p = lib().orc_tracker_create(w, h, levels, *[float(k) for k in K])
        self._ref = None

    def settings(self, huberTH=6.0, coarseCutoffTH=20.0, affA=0.0, affB=0.0):
        lib().orc_tracker_settings(self.p, huberTH, coarseCutoffTH, affA, affB)

    def K(self, lvl):
        o = np.zeros(4, np.float32); lib().orc_tracker_get_K(self.p, lvl, o); return o

    def Ki(self, lvl):
        o = np.zeros(9, np.float32); lib().orc_tracker_get_Ki(self.p, lvl, o); return o.reshape(3, 3)

    def setCoarseTrackingRef(self, ref: Frame, pts, round_half, ref_a=0.0, ref_b=0.0):
        """pts (n,4) float32 {u,v,idepth,HdiF}; round_half (n,) int32."""
        self._ref = ref
        pts = np.ascontiguousarray(pts, np.float32); rh = np.ascontiguousarray(round_half, np.int32)
        lib().orc_tracker_set_ref(self.p, ref.p, pts, rh, len(pts), ref_a, ref_b)

    def setCloud(self, ref: Frame, lvl, u, v, idepth, color, ref_a=0.0, ref_b=0.0):
        self._ref = ref
        a = [np.ascontiguousarray(x, np.float32) for x in (u, v, idepth, color)]
        lib().orc_tracker_set_cloud(self.p, ref.p, lvl, len(a[0]), *a, ref_a, ref_b)

    def cloud(self, lvl):
        n = lib().orc_tracker_cloud_n(self.p, lvl)
        a = [np.zeros(n, np.float32) for _ in range(4)]
        lib().orc_tracker_get_cloud(self.p, lvl, *a)
        return a

    def calcRes(self, new: Frame, lvl, T7, a, b, cutoff):
        rs = np.zeros(6); lib().orc_tracker_calc_res(self.p, new.p, lvl, np.ascontiguousarray(T7, np.float64), a, b, cutoff, rs); return rs

    def warped(self):
        n = lib().orc_tracker_warped_n(self.p); o = np.zeros((8, n), np.float32)
        if n: lib().orc_tracker_get_warped(self.p, o)
        return o

    def calcGSSSE(self, lvl, T7, a, b):
        H = np.zeros(64); bb = np.zeros(8)
        lib().orc_tracker_calc_gs(self.p, lvl, np.ascontiguousarray(T7, np.float64), a, b, H, bb); return H.reshape(8, 8), bb

    def trackNewestCoarse(self, new: Frame, T7, ab, coarsest, minRes=None):
        T = np.array(T7, np.float64); abv = np.array(ab, np.float64)
        minRes = np.full(5, np.nan) if minRes is None else np.ascontiguousarray(minRes, np.float64)
        lastRes = np.zeros(5); flow = np.zeros(3)
        ev = np.zeros(6, np.int64); its = np.zeros(6, np.int32); acc = np.zeros(6, np.int32)
        good = lib().orc_tracker_track(self.p, new.p, T, abv, coarsest, minRes, lastRes, flow, ev, its, acc)
        return dict(good=bool(good), T=T, ab=abv, lastResiduals=lastRes, flow=flow, evals=ev, iterations=its, accepts=acc)

    def __del__(self):
        if getattr(self, "p", None):
            lib().orc_tracker_destroy(self.p); self.p = None


# ------------------------------------------------------------------------------------------------ back-end window
def reproject_map(w, h, levels, K4, kf_frames, kf_T7, kf_ab, cur_frame, cur_T7, cur_ab, pts, cur_kf_index=-1, only_host=-1, backup=False,
                  cell_order=None, max_matches=400):
    """Reprojector::reprojectMap / backprojectMap restated (orc_reproject.cpp).  pts: structured array with u, v, idepth, host, type.
    Returns (pt_index[n], px[n,2]) in cell visiting order."""
    L = lib()
    L.orc_reproject_map.argtypes = [C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.POINTER(C.c_void_p), _f64p, _f64p, C.c_void_p, _f64p, _f64p, C.c_int,
                                    C.c_int, _f32p, C.c_int, C.c_int, C.c_void_p, C.c_int, _i32p, _f64p]
    nH = len(kf_frames); fr = (C.c_void_p * nH)(*[f.p for f in kf_frames])
    p5 = np.ascontiguousarray(np.stack([pts["u"], pts["v"], pts["idepth"], pts["host"].astype(np.float32), pts["type"].astype(np.float32)], 1), np.float32)
    ncells = int(np.ceil(w / 25.0)) * int(np.ceil(h / 25.0))
    out_pt = np.zeros(ncells, np.int32); out_px = np.zeros((ncells, 2))
    co = None if cell_order is None else np.ascontiguousarray(cell_order, np.int32)
    n = L.orc_reproject_map(w, h, levels, np.ascontiguousarray(K4, np.float32), nH, fr, np.ascontiguousarray(kf_T7, np.float64), np.ascontiguousarray(kf_ab, np.float64),
                            cur_frame.p, np.ascontiguousarray(cur_T7, np.float64), np.ascontiguousarray(cur_ab, np.float64), cur_kf_index, len(p5), p5,
                            only_host, 1 if backup else 0, None if co is None else co.ctypes.data, max_matches, out_pt, out_px)
    return out_pt[:n].copy(), out_px[:n].copy()


def track_hypotheses(sprelast_c2w, slast_c2w, lastF_c2w, poses_valid):
    """lastF_2_fh_tries of FullSystem::trackNewCoarse for a running system (FullSystem.cpp:334-394)."""
    if not poses_valid:
        return [np.array([1, 0, 0, 0, 0, 0, 0.0])]
    ROT = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (-1, 0, 0), (0, -1, 0), (0, 0, -1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (-1, 1, 0), (0, -1, 1), (-1, 0, 1), (1, -1, 0),
           (0, 1, -1), (1, 0, -1), (-1, -1, 0), (0, -1, -1), (-1, 0, -1), (-1, -1, -1), (-1, -1, 1), (-1, 1, -1), (-1, 1, 1), (1, -1, -1), (1, -1, 1), (1, 1, -1), (1, 1, 1)]
    if poses_valid == 2:                                                          # second frame of a sequence (FullSystem.cpp:299-331): identity + 26 rotations x {0.02f, 0.04f}
        tries = [np.array([1, 0, 0, 0, 0, 0, 0.0])]
        for rd in (np.float32(0.02), np.float32(0.02) + np.float32(0.02)):
            r = float(rd)
            for q in ROT:
                qq = np.array([1.0, q[0] * r, q[1] * r, q[2] * r]); qq = qq / np.sqrt(qq[1] * qq[1] + qq[2] * qq[2] + qq[3] * qq[3] + qq[0] * qq[0])
                tries.append(np.concatenate([qq, np.zeros(3)]))
        return tries
    slast_2_sprelast = se3_mul(se3_inv(sprelast_c2w), slast_c2w); lastF_2_slast = se3_mul(se3_inv(slast_c2w), lastF_c2w)
    fh_2_slast = slast_2_sprelast; inv = se3_inv(fh_2_slast); cm = se3_mul(inv, lastF_2_slast)
    tries = [cm, se3_mul(se3_mul(inv, inv), lastF_2_slast), se3_mul(se3_inv(se3_exp(se3_log(fh_2_slast) * 0.5)), lastF_2_slast), lastF_2_slast,
             np.array([1, 0, 0, 0, 0, 0, 0.0])]
    r = float(np.float32(0.02))
    for q in [(r, 0, 0), (0, r, 0), (0, 0, r), (-r, 0, 0), (0, -r, 0), (0, 0, -r), (r, r, 0), (0, r, r), (r, 0, r), (-r, r, 0), (0, -r, r), (-r, 0, r), (r, -r, 0),
              (0, r, -r), (r, 0, -r), (-r, -r, 0), (0, -r, -r), (-r, 0, -r), (-r, -r, -r), (-r, -r, r), (-r, r, -r), (-r, r, r), (r, -r, -r), (r, -r, r), (r, r, -r), (r, r, r)]:
        qq = np.array([1.0, q[0], q[1], q[2]]); qq = qq / np.sqrt(qq[1] * qq[1] + qq[2] * qq[2] + qq[3] * qq[3] + qq[0] * qq[0])
        tries.append(se3_mul(cm, np.concatenate([qq, np.zeros(3)])))
    return tries


def track_new_coarse(tracker, new_frame, K4, kf_frames, kf_T7, kf_ab, map_pts, sprelast_c2w, slast_c2w, lastF_c2w, aff_last, poses_valid, lastCoarseRMSE,
                     cell_order=None, max_matches=400):
    """FullSystem::trackNewCoarse restated for a running system (FullSystem.cpp:283-500, branch :334-395) over the oracle pieces:
    hypotheses :346-388, re-track loop :410-462, fallback :464-470, pose composition :474-479, reprojectMap + structPoseEstimation :481-491.
    `tracker` is an orc.CoarseTracker whose reference is lastF.  Pure-Python control flow (<= 31 tries), all numerics in liborc."""
    w, h, L = tracker.w, tracker.h, tracker.levels
    tries = track_hypotheses(sprelast_c2w, slast_c2w, lastF_c2w, poses_valid)
    achieved = np.full(5, np.nan); have = False; flow = np.array([100.0, 100.0, 100.0]); lastF_2_fh = np.array([1, 0, 0, 0, 0, 0, 0.0]); aff = np.zeros(2); n_tries = 0
    for T in tries:
        r = tracker.trackNewestCoarse(new_frame, T, aff_last, L - 1, achieved.copy()); n_tries += 1
        lr = r["lastResiduals"]
        if r["good"] and np.isfinite(np.float32(lr[0])) and not (lr[0] >= achieved[0]):
            flow = r["flow"].copy(); aff = r["ab"].copy(); lastF_2_fh = r["T"].copy(); have = True
        if have:
            for i in range(5):
                if (not np.isfinite(np.float32(achieved[i]))) or achieved[i] > lr[i]:
                    achieved[i] = lr[i]
        if have and achieved[0] < lastCoarseRMSE[0] * float(np.float32(1.5)):
            break
    if not have:
        flow = np.zeros(3); aff = np.array(aff_last, np.float64); lastF_2_fh = tries[0]
    camToWorld = se3_mul(lastF_c2w, se3_inv(lastF_2_fh))
    idx, px = reproject_map(w, h, L, K4, kf_frames, kf_T7, kf_ab, new_frame, camToWorld, aff, map_pts, cell_order=cell_order, max_matches=max_matches)
    p6 = np.stack([map_pts["u"][idx], map_pts["v"][idx], map_pts["idepth"][idx], map_pts["host"][idx].astype(np.float32), px[:, 0].astype(np.float32), px[:, 1].astype(np.float32)], 1).astype(np.float32) \
        if len(idx) else np.zeros((0, 6), np.float32)
    sp = struct_pose(w, h, np.asarray(K4, np.float32), kf_T7, p6, camToWorld)
    return dict(camToWorld=sp["T"], camToTrackingRef=se3_mul(se3_inv(lastF_c2w), sp["T"]), aff_g2l=aff, flow=flow, lastCoarseRMSE=achieved, have_one_good=have, tries=n_tries,
                n_matches=len(idx), refine_iterations=sp["iterations"], refine_accepts=sp["accepts"], camToWorld_tracked=camToWorld)


def struct_pose_hb(w, h, K4, host_T7, pts6, curToWorld7):
    """calcHandb + calculateRes at one pose: returns (H 6x6, b 6, mean squared pixel error, number of in-frame points)."""
    L = lib()
    L.orc_struct_pose_hb.argtypes = [C.c_int, C.c_int, _f32p, C.c_int, _f64p, C.c_int, _f32p, _f64p, _f64p, _f64p, _i32p]; L.orc_struct_pose_hb.restype = C.c_float
    hT = np.ascontiguousarray(host_T7, np.float64).reshape(-1, 7); p = np.ascontiguousarray(pts6, np.float32).reshape(-1, 6)
    H = np.zeros(36); b = np.zeros(6); num = np.zeros(1, np.int32)
    e = L.orc_struct_pose_hb(w, h, np.ascontiguousarray(K4, np.float32), len(hT), hT, len(p), p, np.ascontiguousarray(curToWorld7, np.float64), H, b, num)
    return H.reshape(6, 6), b, float(e), int(num[0])


def struct_pose(w, h, K4, host_T7, pts6, curToWorld7):
    """CoarseTracker::structPoseEstimation restated (orc_refine.cpp).  pts6: (n,6) float32 {u,v,idepth,host,obs_x,obs_y}."""
    L = lib()
    L.orc_struct_pose.argtypes = [C.c_int, C.c_int, _f32p, C.c_int, _f64p, C.c_int, _f32p, _f64p, _i32p]; L.orc_struct_pose.restype = C.c_float
    hT = np.ascontiguousarray(host_T7, np.float64).reshape(-1, 7); p = np.ascontiguousarray(pts6, np.float32).reshape(-1, 6)
    T = np.array(curToWorld7, np.float64).copy(); st = np.zeros(2, np.int32)
    if len(p) == 0:
        p = np.zeros((1, 6), np.float32); n = 0
    else:
        n = len(p)
    res = L.orc_struct_pose(w, h, np.ascontiguousarray(K4, np.float32), len(hT), hT, n, p, T, st)
    return dict(T=T, res=float(res), iterations=int(st[0]), accepts=int(st[1]))


def _ba_protos():
    L = lib()
    if getattr(L, "_ba_done", False):
        return L
    L.orc_ba_create.restype = C.c_void_p; L.orc_ba_create.argtypes = [C.c_int, C.c_int]
    L.orc_ba_destroy.argtypes = [C.c_void_p]
    L.orc_ba_set_calib.argtypes = [C.c_void_p, _f64p]
    L.orc_ba_add_frame.argtypes = [C.c_void_p, C.c_void_p, _f64p, _f64p, _f64p, C.c_float, C.c_int, C.c_float]
    L.orc_ba_set_points.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _i32p, _i32p, _i32p, _i32p]
    L.orc_ba_set_residuals.argtypes = [C.c_void_p, C.c_int, _i32p, _i32p, _i32p, _i32p, _f32p, _i32p]
    L.orc_ba_set_prior.argtypes = [C.c_void_p, _f64p, _f64p]
    for nm in ("orc_ba_init", "orc_ba_reset_oob", "orc_ba_apply_res", "orc_ba_backup", "orc_ba_load_backup"):
        getattr(L, nm).argtypes = [C.c_void_p]
    L.orc_ba_linearize_all.argtypes = [C.c_void_p, C.c_int]; L.orc_ba_linearize_all.restype = C.c_double
    L.orc_ba_energy_L.argtypes = [C.c_void_p]; L.orc_ba_energy_L.restype = C.c_double
    L.orc_ba_energy_M.argtypes = [C.c_void_p]; L.orc_ba_energy_M.restype = C.c_double
    L.orc_ba_get_residuals.argtypes = [C.c_void_p, _i32p, _i32p, _f64p, _i32p, _f32p, _f32p, _f32p, _f32p, _i32p]
    L.orc_ba_accumulate.argtypes = [C.c_void_p, _f64p, _f64p, _f64p, _f64p]
    L.orc_ba_solve.argtypes = [C.c_void_p, C.c_int, C.c_double, _f64p, _f64p, _f64p]
    L.orc_ba_do_step.argtypes = [C.c_void_p, C.c_float]
    L.orc_ba_get_points.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, _f32p, _f32p, _i32p, _f32p]
    L.orc_ba_get_frames.argtypes = [C.c_void_p, _f64p, _f64p, _f64p, _f32p, _f64p]
    L.orc_ba_get_calib.argtypes = [C.c_void_p, _f64p, _f64p]
    L.orc_ba_get_precalc.argtypes = [C.c_void_p, C.c_int, C.c_int, _f32p, _f64p, _f64p, _f32p]
    L.orc_ba_optimize.argtypes = [C.c_void_p, C.c_int, _i32p]; L.orc_ba_optimize.restype = C.c_float
    L.orc_ba_linearize_calls.argtypes = [C.c_void_p]; L.orc_ba_linearize_calls.restype = C.c_longlong
    L.orc_ba_flag_points.argtypes = [C.c_void_p, _i32p, _i32p]
    L.orc_ba_marginalize_points.argtypes = [C.c_void_p, _i32p, _f64p, _f64p, _f64p, _f64p]
    L.orc_ba_marginalize_frame.argtypes = [C.c_void_p, C.c_int]
    L.orc_ba_dim.argtypes = [C.c_void_p]
    L.orc_ba_get_prior.argtypes = [C.c_void_p, _f64p, _f64p]
    L.orc_ba_get_res_to_zero.argtypes = [C.c_void_p, _f32p, _i32p]
    L._ba_done = True
    return L


class BAWindow:
    """The flattened sliding window of synth.make_ba_window() loaded into the oracle back-end (orc_ba.hpp)."""

    def __init__(self, win: dict, frames):
        L = _ba_protos(); self.L = L; self.win = win; self._frames = frames
        w, h = win["wh"]; self.nF = win["nF"]; self.nP = len(win["uv"]); self.nR = len(win["r_point"]); self.n = 4 + 6 * self.nF
        self.p = L.orc_ba_create(w, h)
        L.orc_ba_set_calib(self.p, np.ascontiguousarray(win["K"], np.float64))
        if "K_zero" in win: L.orc_ba_set_calib_zero.argtypes = [C.c_void_p, _f64p]; L.orc_ba_set_calib_zero(self.p, np.ascontiguousarray(win["K_zero"], np.float64))   # live window: value_zero = initial intrinsics
        for i in range(self.nF):
            L.orc_ba_add_frame(self.p, frames[i].p, np.ascontiguousarray(win["T_eval"][i]), np.ascontiguousarray(win["state"][i]),
                               np.ascontiguousarray(win["state_zero"][i]), float(win["ab_exposure"][i]), int(win["frameID"][i]), float(win["frameEnergyTH"][i]))
        c = lambda k, t: np.ascontiguousarray(win[k], t)
        L.orc_ba_set_points(self.p, self.nP, c("uv", np.float32), c("idepth", np.float32), c("idepth_zero", np.float32), c("color", np.float32),
                            c("weights", np.float32), c("host", np.int32), c("hasDepthPrior", np.int32), c("isFromSensor", np.int32), c("res_begin", np.int32))
        L.orc_ba_set_residuals(self.p, self.nR, c("r_point", np.int32), c("r_host", np.int32), c("r_target", np.int32), c("r_hasMatcher", np.int32),
                               c("r_matcher", np.float32), c("r_isNew", np.int32))
        L.orc_ba_set_prior(self.p, c("HM", np.float64), c("bM", np.float64))
        L.orc_ba_init(self.p)

    def reset_oob(self): self.L.orc_ba_reset_oob(self.p)
    def linearizeAll(self, fix=False): return self.L.orc_ba_linearize_all(self.p, 1 if fix else 0)
    def applyRes(self): self.L.orc_ba_apply_res(self.p)
    def calcLEnergy(self): return self.L.orc_ba_energy_L(self.p)
    def calcMEnergy(self): return self.L.orc_ba_energy_M(self.p)
    def backupState(self): self.L.orc_ba_backup(self.p)
    def doStepFromBackup(self, f=1.0): return bool(self.L.orc_ba_do_step(self.p, f))
    def loadStateBackup(self): self.L.orc_ba_load_backup(self.p)

    def residuals(self):
        n = self.nR
        o = dict(state=np.zeros(n, np.int32), new_state=np.zeros(n, np.int32), energies=np.zeros((n, 3)), active=np.zeros(n, np.int32),
                 J=np.zeros((n, 24), np.float32), efJ=np.zeros((n, 24), np.float32), JpJdF=np.zeros((n, 8), np.float32),
                 center=np.zeros((n, 3), np.float32), toRemove=np.zeros(n, np.int32))
        self.L.orc_ba_get_residuals(self.p, o["state"], o["new_state"], o["energies"], o["active"], o["J"], o["efJ"], o["JpJdF"], o["center"], o["toRemove"])
        return o

    def accumulate(self):
        n = self.n; HA = np.zeros((n, n)); bA = np.zeros(n); Hsc = np.zeros((n, n)); bsc = np.zeros(n)
        self.L.orc_ba_accumulate(self.p, HA, bA, Hsc, bsc); return HA, bA, Hsc, bsc

    def solveSystem(self, iteration, lam):
        n = self.n; x = np.zeros(n); HS = np.zeros((n, n)); bS = np.zeros(n)
        self.L.orc_ba_solve(self.p, iteration, lam, x, HS, bS); return x, HS, bS

    def points(self):
        n = self.nP
        o = dict(idepth=np.zeros(n, np.float32), step=np.zeros(n, np.float32), HdiF=np.zeros(n, np.float32), bdSumF=np.zeros(n, np.float32),
                 maxRelBaseline=np.zeros(n, np.float32), numGood=np.zeros(n, np.int32), idepth_hessian=np.zeros(n, np.float32))
        self.L.orc_ba_get_points(self.p, o["idepth"], o["step"], o["HdiF"], o["bdSumF"], o["maxRelBaseline"], o["numGood"], o["idepth_hessian"]); return o

    def frames(self):
        n = self.nF
        o = dict(T_eval=np.zeros((n, 7)), state=np.zeros((n, 10)), step=np.zeros((n, 10)), frameEnergyTH=np.zeros(n, np.float32), PRE_worldToCam=np.zeros((n, 7)))
        self.L.orc_ba_get_frames(self.p, o["T_eval"], o["state"], o["step"], o["frameEnergyTH"], o["PRE_worldToCam"]); return o

    def calib(self):
        v = np.zeros(4); s = np.zeros(4); self.L.orc_ba_get_calib(self.p, v, s); return v, s

    def precalc(self, host, target):
        o = np.zeros(27, np.float32); aH = np.zeros(36); aT = np.zeros(36); d = np.zeros(6, np.float32)
        self.L.orc_ba_get_precalc(self.p, host, target, o, aH, aT, d)
        return dict(KRKi=o[:9].reshape(3, 3), Kt=o[9:12], R0=o[12:21].reshape(3, 3), t0=o[21:24], aff=o[24:26], b0=o[26], adHost=aH.reshape(6, 6), adTarget=aT.reshape(6, 6), adHTdelta=d)

    def optimize(self, its=6):
        st = np.zeros(2, np.int32); rmse = self.L.orc_ba_optimize(self.p, its, st)
        return dict(rmse=float(rmse), iterations=int(st[0]), accepts=int(st[1]), linearize_calls=int(self.L.orc_ba_linearize_calls(self.p)))

    # ---- keyframe hand-over
    def flagPointsForRemoval(self, selected):
        st = np.zeros(self.nP, np.int32); self.L.orc_ba_flag_points(self.p, np.ascontiguousarray(selected, np.int32), st); return st

    def marginalizePointsF(self, status):
        n = self.L.orc_ba_dim(self.p); M = np.zeros((n, n)); Mb = np.zeros(n); S = np.zeros((n, n)); Sb = np.zeros(n)
        self.L.orc_ba_marginalize_points(self.p, np.ascontiguousarray(status, np.int32), M, Mb, S, Sb)
        return dict(M=M, Mb=Mb, Msc=S, Mbsc=Sb)

    def marginalizeFrame(self, idx): self.L.orc_ba_marginalize_frame(self.p, int(idx))

    def prior(self):
        n = self.L.orc_ba_dim(self.p); HM = np.zeros((n, n)); bM = np.zeros(n); self.L.orc_ba_get_prior(self.p, HM, bM); return HM, bM

    def res_to_zero(self):
        r = np.zeros((self.nR, 2), np.float32); l = np.zeros(self.nR, np.int32); self.L.orc_ba_get_res_to_zero(self.p, r, l); return r, l

    def __del__(self):
        if getattr(self, "p", None):
            self.L.orc_ba_destroy(self.p); self.p = None


# ---------------------------------------------------------------------------------------------- immature points (orc_trace.cpp): ImmaturePoint ctor + traceOn
IMM_DTYPE = np.dtype([("u", np.float32), ("v", np.float32), ("idepth_min", np.float32), ("idepth_max", np.float32), ("color", np.float32, 8), ("weights", np.float32, 8),
                      ("gradH", np.float32, 4), ("energyTH", np.float32), ("quality", np.float32), ("lastTraceUV", np.float32, 2), ("lastTracePixelInterval", np.float32),
                      ("lastTraceStatus", np.int32)])
IPS_GOOD, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED = range(6)


def immature_init(host: "Frame", uv):
    """ImmaturePoint::ImmaturePoint (ImmaturePoint.cpp:8-36) for integer pixels uv (n,2) of the host keyframe"""
    L = lib(); L.orc_immature_init.argtypes = [C.c_void_p, C.c_int, _i32p, C.c_void_p]; L.orc_immature_bytes.restype = C.c_int
    assert L.orc_immature_bytes() == IMM_DTYPE.itemsize
    uv = np.ascontiguousarray(uv, np.int32).reshape(-1, 2); P = np.zeros(len(uv), IMM_DTYPE)
    L.orc_immature_init(host.p, len(uv), uv, P.ctypes.data)
    return P


def immature_trace(frame: "Frame", pts, KRKi, Kt, aff):
    """ImmaturePoint::traceOn (ImmaturePoint.cpp:50-352) of every candidate in pts (IMM_DTYPE, one host) against `frame`; updates pts in place, returns the statuses"""
    L = lib(); L.orc_immature_trace.argtypes = [C.c_void_p, C.c_int, C.c_void_p, _f32p, _f32p, _f32p, _i32p]
    assert pts.dtype == IMM_DTYPE and pts.flags.c_contiguous
    st = np.zeros(len(pts), np.int32)
    L.orc_immature_trace(frame.p, len(pts), pts.ctypes.data, np.ascontiguousarray(KRKi, np.float32).reshape(-1), np.ascontiguousarray(Kt, np.float32), np.ascontiguousarray(aff, np.float32), st)
    return st


def trace_geometry(K4, host_c2w7, new_c2w7, host_exposure=1.0, new_exposure=1.0, host_ab=(0.0, 0.0), new_ab=(0.0, 0.0)):
    """KRKi, Kt, aff of FullSystem::traceNewCoarse (FullSystem.cpp:525-538): hostToNew = new_worldToCam * host_camToWorld, floats as the reference casts them"""
    fx, fy, cx, cy = [np.float32(x) for x in K4]
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
    T = se3_mul(se3_inv(np.asarray(new_c2w7, np.float64)), np.asarray(host_c2w7, np.float64))
    qw, qx, qy, qz = T[:4]
    R = np.array([[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)], [2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)],
                  [2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)]]).astype(np.float32)
    Ki = np.linalg.inv(K.astype(np.float64)).astype(np.float32)
    KRKi = (K @ R @ Ki).astype(np.float32); Kt = (K @ T[4:].astype(np.float32)).astype(np.float32)
    a = np.exp(new_ab[0] - host_ab[0]) * new_exposure / host_exposure
    return KRKi, Kt, np.array([a, new_ab[1] - a * host_ab[1]], np.float32)


def immature_optimize(pts, is_from_sensor, target_frames, pre14, calib6, min_obs=1):
    """FullSystem::optimizeImmaturePoint (FullSystemOptPoint.cpp:18-183) for candidates of ONE host against the target frames (window order without the host).
    pre14 (nres,14): PRE_RTll, PRE_tTll, PRE_aff_mode of (host,target); calib6: fxl fyl cxl cyl fxli fyli.  Returns status (0 stay, -1 drop, 1 activate), idepth, res states."""
    L = lib(); n = len(pts); nres = len(target_frames)
    L.orc_immature_optimize.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p), _f32p, _f32p, C.c_int, _i32p, _f32p, _i32p]
    assert pts.dtype == IMM_DTYPE and pts.flags.c_contiguous
    fs = np.ascontiguousarray(is_from_sensor, np.uint8); tf = (C.c_void_p * nres)(*[f.p for f in target_frames])
    st = np.zeros(n, np.int32); idp = np.zeros(n, np.float32); rs = np.zeros((n, nres), np.int32)
    L.orc_immature_optimize(n, pts.ctypes.data, fs.ctypes.data, nres, tf, np.ascontiguousarray(pre14, np.float32).reshape(-1), np.ascontiguousarray(calib6, np.float32), min_obs, st, idp, rs.reshape(-1))
    return st, idp, rs


# ---------------------------------------------------------------------------------------------- candidate management (orc_select.cpp): PixelSelector, makeNewTraces, CoarseDistanceMap
NEW_TRACE_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("my_type", "<f4"), ("score", "<f4"), ("idepth_fromSensor", "<f4"), ("isFromSensor", "<i4"), ("type", "<i4")])
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def libc_random_pattern(w: int, h: int):
    """PixelSelector::PixelSelector (PixelSelector2.cpp:14-16): srand(3141592); randomPattern[i] = rand() & 0xFF — glibc's rand(), through ctypes"""
    libc = C.CDLL(None); libc.srand(3141592)
    return np.array([libc.rand() & 0xFF for _ in range(w * h)], np.uint8)


def _cloud(cloud3):
    return None if cloud3 is None else np.ascontiguousarray(cloud3, np.float64).reshape(-1, 3)


class Selector:
    """PixelSelector (FullSystem/PixelSelector2.cpp) on orc Frames; state = currentPotential + the histogram thresholds of the last frame."""

    def __init__(self, w, h, random_pattern):
        L = lib(); self.w, self.h = w, h
        L.orc_selector_create.restype = C.c_void_p; L.orc_selector_create.argtypes = [C.c_int, C.c_int, _u8p]
        L.orc_selector_destroy.argtypes = [C.c_void_p]; L.orc_selector_set_potential.argtypes = [C.c_void_p, C.c_int]; L.orc_selector_get_potential.argtypes = [C.c_void_p]
        L.orc_selector_make_hists.argtypes = [C.c_void_p, C.c_void_p, _f32p, _f32p]
        L.orc_selector_select.argtypes = [C.c_void_p, C.c_void_p, _f32p, C.c_int, C.c_float, C.c_void_p, C.c_int, _i32p]
        L.orc_selector_make_maps.argtypes = [C.c_void_p, C.c_void_p, _f32p, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_int]
        L.orc_make_new_traces.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, _f32p, C.c_void_p, C.c_int, _i32p, _i32p]
        L.orc_shi_tomasi.restype = C.c_float; L.orc_shi_tomasi.argtypes = [C.c_void_p, C.c_int, C.c_int]
        self.p = L.orc_selector_create(w, h, np.ascontiguousarray(random_pattern, np.uint8))

    @property
    def currentPotential(self): return lib().orc_selector_get_potential(self.p)
    @currentPotential.setter
    def currentPotential(self, v): lib().orc_selector_set_potential(self.p, int(v))

    def makeHists(self, frame):
        n = (self.w // 32) * (self.h // 32); a = np.zeros(n, np.float32); b = np.zeros(n, np.float32); lib().orc_selector_make_hists(self.p, frame.p, a, b); return a, b

    def select(self, frame, pot, thFactor=1.0, cloud3=None):
        """one pass of select (cloud3 None: map over pixels) / selectFromLidar (map over cloud rows); makeHists must have run.  -> map, (n2, n3, n4)"""
        c = _cloud(cloud3); m = np.zeros(self.w * self.h if c is None else max(len(c), 1), np.float32); n3 = np.zeros(3, np.int32)
        lib().orc_selector_select(self.p, frame.p, m, pot, thFactor, None if c is None else c.ctypes.data, 0 if c is None else len(c), n3)
        return (m.reshape(self.h, self.w) if c is None else m[:len(c)]), n3

    def makeMaps(self, frame, density, recursionsLeft=1, thFactor=1.0, cloud3=None):
        """makeMaps / makeMapsFromLidar (makeHists included) -> map, numHaveSub"""
        c = _cloud(cloud3); m = np.zeros(self.w * self.h if c is None else max(len(c), 1), np.float32)
        n = lib().orc_selector_make_maps(self.p, frame.p, m, density, recursionsLeft, thFactor, None if c is None else c.ctypes.data, 0 if c is None else len(c))
        return (m.reshape(self.h, self.w) if c is None else m[:len(c)]), n

    def makeNewTraces(self, frame, cloud3, densityLidar, densityDense, addFeaturePoint, selectionMap, cap=1 << 16):
        """FullSystem::makeNewTraces; selectionMap (h, w) float32 is updated in place when addFeaturePoint.  -> NEW_TRACE_DTYPE records, (numPointLidar, numPointMonocular), passes"""
        assert lib().orc_new_trace_bytes() == NEW_TRACE_DTYPE.itemsize and selectionMap.dtype == np.float32 and selectionMap.flags.c_contiguous
        c = _cloud(cloud3); out = np.zeros(cap, NEW_TRACE_DTYPE); num = np.zeros(2, np.int32); passes = np.zeros(2, np.int32)
        m = lib().orc_make_new_traces(self.p, frame.p, c.ctypes.data, len(c), densityLidar, densityDense, int(addFeaturePoint), selectionMap.reshape(-1), out.ctypes.data, cap, num, passes)
        assert m <= cap
        return out[:m], num, passes

    def __del__(self):
        if getattr(self, "p", None) and _LIB is not None:
            _LIB.orc_selector_destroy(self.p); self.p = None


def shi_tomasi(frame, u, v):
    L = lib(); L.orc_shi_tomasi.restype = C.c_float; L.orc_shi_tomasi.argtypes = [C.c_void_p, C.c_int, C.c_int]; return L.orc_shi_tomasi(frame.p, int(u), int(v))


def lidar_density(lrud, wh, desiredImmatureDensity):
    """((float)lidarArea/(float)imageArea) * setting_desiredImmatureDensity (FullSystem.cpp:1287-1290), in float like the reference"""
    area = (lrud[1] - lrud[0]) * (lrud[3] - lrud[2])
    return float(np.float32(np.float32(area) / np.float32(wh[0] * wh[1])) * np.float32(desiredImmatureDensity))


def distmap_geometry(K4, host_c2w7, new_c2w7):
    """KRKi = K[1] R Ki[0], Kt = K[1] t of host -> newest (FullSystem.cpp:606-608 / CoarseTracker.cpp:1156-1158), floats as the reference casts them; level-1 K per CoarseDistanceMap::makeK"""
    fx, fy, cx, cy = [np.float32(k) for k in K4]
    K0 = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
    fx1, fy1 = np.float32(np.float64(fx) * 0.5), np.float32(np.float64(fy) * 0.5)
    cx1, cy1 = np.float32((np.float64(cx) + 0.5) / 2 - 0.5), np.float32((np.float64(cy) + 0.5) / 2 - 0.5)
    K1 = np.array([[fx1, 0, cx1], [0, fy1, cy1], [0, 0, 1]], np.float32)
    return K0, K1


class DistMap:
    """CoarseDistanceMap (CoarseTracker.cpp:1139-1282) at level-1 resolution, and the candidate walk of activatePointsMT"""

    def __init__(self, w1, h1):
        L = lib(); self.w1, self.h1 = w1, h1
        L.orc_distmap_create.restype = C.c_void_p; L.orc_distmap_create.argtypes = [C.c_int, C.c_int]; L.orc_distmap_destroy.argtypes = [C.c_void_p]
        L.orc_distmap_make.argtypes = [C.c_void_p, C.c_int, _i32p, _f32p, _f32p, _f32p]; L.orc_distmap_add.argtypes = [C.c_void_p, C.c_int, C.c_int]; L.orc_distmap_get.argtypes = [C.c_void_p, _f32p]
        L.orc_activate_select.argtypes = [C.c_void_p, C.c_int, _i32p, _f32p, _f32p, _f32p, C.c_float, _i32p]
        self.p = L.orc_distmap_create(w1, h1)

    def make(self, pt_begin, KRKi, Kt, uvid):
        pt_begin = np.ascontiguousarray(pt_begin, np.int32)
        lib().orc_distmap_make(self.p, len(pt_begin) - 1, pt_begin, np.ascontiguousarray(KRKi, np.float32).reshape(-1), np.ascontiguousarray(Kt, np.float32).reshape(-1), np.ascontiguousarray(uvid, np.float32).reshape(-1))

    def add(self, u, v): lib().orc_distmap_add(self.p, int(u), int(v))

    def get(self):
        o = np.zeros(self.w1 * self.h1, np.float32); lib().orc_distmap_get(self.p, o); return o.reshape(self.h1, self.w1)

    def activateSelect(self, cand_begin, KRKi, Kt, cand4, currentMinActDist):
        cand_begin = np.ascontiguousarray(cand_begin, np.int32); dec = np.zeros(cand_begin[-1], np.int32)
        lib().orc_activate_select(self.p, len(cand_begin) - 1, cand_begin, np.ascontiguousarray(KRKi, np.float32).reshape(-1), np.ascontiguousarray(Kt, np.float32).reshape(-1),
                                  np.ascontiguousarray(cand4, np.float32).reshape(-1), currentMinActDist, dec)
        return dec

    def __del__(self):
        if getattr(self, "p", None) and _LIB is not None:
            _LIB.orc_distmap_destroy(self.p); self.p = None


# ---------------------------------------------------------------------------------------------- LiDAR front-end of the node (orc_lidar.cpp): src/main.cpp:563-858
class LidarFrontEnd:
    """projectPointCloud -> groundRemoval -> cloudSegmentation -> pixel projection of lidarCloudHandler; lrud = FullSystem::left/right/up/down (running box)"""

    def __init__(self, n_scan=64, horizon=1800, ang_res_x=0.2, ang_res_y=0.427, ang_bottom=24.9, groundScanInd=50):
        L = lib(); self.n_scan, self.horizon = n_scan, horizon
        L.orc_lidar_create.restype = C.c_void_p; L.orc_lidar_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]; L.orc_lidar_destroy.argtypes = [C.c_void_p]
        L.orc_lidar_handler.argtypes = [C.c_void_p, _f32p, C.c_int, _f64p, _f64p, _f32p, C.c_int, C.c_int, _i32p, _f64p, C.c_int, _i32p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.p = L.orc_lidar_create(n_scan, horizon, ang_res_x, ang_res_y, ang_bottom, groundScanInd)

    def handle(self, xyzi, Rlc, tlc, K4, wh, lrud, images=False):
        xyzi = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4); cap = self.n_scan * self.horizon; out = np.zeros((cap, 3)); flags = np.zeros(4, np.int32)
        lrud = np.ascontiguousarray(lrud, np.int32).copy(); m = self.n_scan * self.horizon
        rng, lab, gnd = (np.zeros(m, np.float32), np.zeros(m, np.int32), np.zeros(m, np.int8)) if images else (None, None, None)
        k = lib().orc_lidar_handler(self.p, xyzi.reshape(-1), len(xyzi), np.ascontiguousarray(Rlc, np.float64).reshape(-1), np.ascontiguousarray(tlc, np.float64), np.ascontiguousarray(K4, np.float32),
                                    wh[0], wh[1], lrud, out.reshape(-1), cap, flags, *(a.ctypes.data if a is not None else None for a in (rng, lab, gnd)))
        assert k >= 0
        r = dict(cloud_px=out[:k].copy(), lrud=lrud, addFeaturePoint=int(flags[0]), numGround=int(flags[1]), numAll=int(flags[2]), n_segmented=int(flags[3]))
        if images: r.update(range=rng.reshape(self.n_scan, self.horizon), label=lab.reshape(self.n_scan, self.horizon), ground=gnd.reshape(self.n_scan, self.horizon))
        return r

    def __del__(self):
        if getattr(self, "p", None) and _LIB is not None:
            _LIB.orc_lidar_destroy(self.p); self.p = None
