"""ctypes binding of the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs — never by the product package."""
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
