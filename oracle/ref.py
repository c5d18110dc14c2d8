"""ctypes binding of oracle/_ref/libsdvref.so — the reference's OWN hot-path sources (unmodified, /root/reference/src) compiled against the
stand-in headers of oracle/ref_stub/ (see oracle/Makefile, oracle/ref_shim.cpp).  TEST INFRASTRUCTURE ONLY.

What it is for: pinning the oracle restatement (oracle/orc_*.cpp) on reference-compiled code (tests/test_ref_pin.py) and, where the library
exists, serving as the "reference" CPU arm of bench.py.  It is built only where /root/reference exists; elsewhere the prebuilt .so is used.
The reference keeps its calibration and settings in process globals (util/globalCalib.cpp, util/settings.cpp): one image size at a time.
"""
from __future__ import annotations
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libsdvref.so")
REF_SRC = "/root/reference/src"
_LIB = None

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_vp = C.c_void_p


def available() -> bool:
    return os.path.exists(SO) or os.path.isdir(REF_SRC)


def build(force: bool = False) -> str | None:
    """Compile the reference translation units (only possible where /root/reference exists); returns the .so path or None."""
    if os.path.isdir(REF_SRC):
        deps = [os.path.join(_HERE, "ref_shim.cpp"), os.path.join(_HERE, "orc_math.hpp"), os.path.join(_HERE, "Makefile")]
        for d, _, fs in os.walk(os.path.join(_HERE, "ref_stub")):
            deps += [os.path.join(d, f) for f in fs]
        stale = force or not os.path.exists(SO) or any(os.path.getmtime(p) > os.path.getmtime(SO) for p in deps)
        if stale:
            subprocess.check_call(["make", "-C", _HERE, "-s", "-j8", "ref"])
    return SO if os.path.exists(SO) else None


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        if so is None:
            raise RuntimeError("oracle/_ref/libsdvref.so is not built and /root/reference is absent")
        devnull = os.open(os.devnull, os.O_WRONLY)
        L = C.CDLL(so, mode=os.RTLD_NOW)
        os.close(devnull)
        L.ref_set_calib.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
        L.ref_get_global_K.argtypes = [C.c_int, _f32p, _f32p]
        L.ref_settings.argtypes = [C.c_float] * 4
        L.ref_frame_create.restype = _vp; L.ref_frame_create.argtypes = [_f32p, C.c_float]
        L.ref_frame_destroy.argtypes = [_vp]
        L.ref_frame_dI.restype = C.POINTER(C.c_float); L.ref_frame_dI.argtypes = [_vp, C.c_int]
        L.ref_frame_abs.restype = C.POINTER(C.c_float); L.ref_frame_abs.argtypes = [_vp, C.c_int]
        L.ref_tracker_create.restype = _vp; L.ref_tracker_create.argtypes = []
        L.ref_tracker_destroy.argtypes = [_vp]
        L.ref_tracker_get_K.argtypes = [_vp, C.c_int, _f32p]; L.ref_tracker_get_Ki.argtypes = [_vp, C.c_int, _f32p]
        L.ref_tracker_set_ref.argtypes = [_vp, _vp, _vp, _f32p, _i32p, C.c_int, C.c_double, C.c_double]
        L.ref_tracker_cloud_n.argtypes = [_vp, C.c_int]
        L.ref_tracker_get_cloud.argtypes = [_vp, C.c_int, _f32p, _f32p, _f32p, _f32p]
        L.ref_tracker_calc_res.argtypes = [_vp, _vp, C.c_int, _f64p, C.c_double, C.c_double, C.c_float, _f64p]
        L.ref_tracker_warped_n.argtypes = [_vp]; L.ref_tracker_get_warped.argtypes = [_vp, _f32p]
        L.ref_tracker_calc_gs.argtypes = [_vp, C.c_int, _f64p, C.c_double, C.c_double, _f64p, _f64p]
        L.ref_tracker_track.argtypes = [_vp, _vp, _f64p, _f64p, C.c_int, _f64p, _f64p, _f64p]
        L.ref_interp33.argtypes = [_f32p, C.c_int, C.c_float, C.c_float, _f32p]
        L.ref_interp33_bilin.argtypes = [_f32p, C.c_int, C.c_float, C.c_float, _f32p]
        L.ref_aff_from_to.argtypes = [C.c_float, C.c_float, C.c_double, C.c_double, C.c_double, C.c_double, _f64p]
        L.ref_ba_create.restype = _vp; L.ref_ba_create.argtypes = []
        L.ref_ba_destroy.argtypes = [_vp]
        L.ref_ba_set_calib.argtypes = [_vp, _f64p]
        L.ref_ba_add_frame.argtypes = [_vp, _vp, _f64p, _f64p, _f64p, C.c_float, C.c_int, C.c_float]
        L.ref_ba_set_points.argtypes = [_vp, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _i32p, _i32p, _i32p, _i32p]
        L.ref_ba_set_residuals.argtypes = [_vp, C.c_int, _i32p, _i32p, _i32p, _i32p, _f32p, _i32p]
        L.ref_ba_set_prior.argtypes = [_vp, _f64p, _f64p]
        for nm in ("ref_ba_init", "ref_ba_reset_oob", "ref_ba_apply_res", "ref_ba_backup", "ref_ba_load_backup", "ref_ba_marginalize_points", "ref_ba_drop_points"):
            getattr(L, nm).argtypes = [_vp]
        L.ref_ba_linearize_all.argtypes = [_vp, C.c_int]; L.ref_ba_linearize_all.restype = C.c_double
        L.ref_ba_energy_L.argtypes = [_vp]; L.ref_ba_energy_L.restype = C.c_double
        L.ref_ba_energy_M.argtypes = [_vp]; L.ref_ba_energy_M.restype = C.c_double
        L.ref_ba_get_residuals.argtypes = [_vp, _i32p, _i32p, _f64p, _i32p, _f32p, _f32p, _f32p, _f32p, _i32p]
        L.ref_ba_accumulate.argtypes = [_vp, _f64p, _f64p, _f64p, _f64p]
        L.ref_ba_solve.argtypes = [_vp, C.c_int, C.c_double, _f64p, _f64p, _f64p]
        L.ref_ba_do_step.argtypes = [_vp, C.c_float]
        L.ref_ba_optimize.argtypes = [_vp, C.c_int]; L.ref_ba_optimize.restype = C.c_float
        L.ref_ba_get_points.argtypes = [_vp, _f32p, _f32p, _f32p, _f32p, _f32p, _i32p, _f32p]
        L.ref_ba_get_frames.argtypes = [_vp, _f64p, _f64p, _f64p, _f32p, _f64p]
        L.ref_ba_get_calib.argtypes = [_vp, _f64p, _f64p]
        L.ref_ba_get_precalc.argtypes = [_vp, C.c_int, C.c_int, _f32p, _f64p, _f64p, _f32p]
        L.ref_ba_flag_points.argtypes = [_vp, _i32p, _i32p]
        L.ref_ba_marginalize_frame.argtypes = [_vp, C.c_int]
        L.ref_ba_dim.argtypes = [_vp]
        L.ref_ba_get_prior.argtypes = [_vp, _f64p, _f64p]
        L.ref_ba_get_res_to_zero.argtypes = [_vp, _f32p, _i32p]
        L.ref_sys_create.restype = _vp; L.ref_sys_create.argtypes = [C.c_int]
        L.ref_sys_destroy.argtypes = [_vp]
        L.ref_srand.argtypes = [C.c_uint]
        L.ref_sys_add_frame.argtypes = [_vp, _f32p, C.c_float, C.c_double, _f64p, C.c_int]
        L.ref_sys_num_frames.argtypes = [_vp]; L.ref_sys_num_keyframes.argtypes = [_vp]
        L.ref_sys_get_frame.argtypes = [_vp, C.c_int, _f64p, _f64p, _i32p]
        L.ref_sys_get_last_rmse.argtypes = [_vp, _f64p]
        L.ref_sys_window.argtypes = [_vp, _i32p, _i32p]
        L.ref_sys_tracker_info.argtypes = [_vp, C.POINTER(C.c_int), _f64p, C.POINTER(C.c_float), _i32p, C.POINTER(C.c_double)]
        L.ref_sys_tracker_cloud.argtypes = [_vp, C.c_int, _f32p, _f32p, _f32p, _f32p]
        L.ref_sys_tracker_K.argtypes = [_vp, _f32p, _f64p]
        L.ref_sys_history.argtypes = [_vp, _f64p, _f64p, _f64p, _f64p, _f64p, C.POINTER(C.c_int)]
        L.ref_sys_map_size.argtypes = [_vp, C.POINTER(C.c_int)]
        L.ref_sys_map.argtypes = [_vp, _i32p, _f64p, _f64p, _f32p, _f32p]
        L.ref_sys_get_track_result.argtypes = [_vp, C.c_int, _f64p]
        L.ref_reproject_map.argtypes = [C.c_int, C.POINTER(_vp), _f64p, _f64p, _vp, _f64p, _f64p, C.c_int, _f32p, C.c_uint, C.c_int, _i32p, _f64p]
        L.ref_ba_immature_pre.argtypes = [_vp, C.c_int, _f32p, _f32p]
        L.ref_ba_optimize_immature.argtypes = [_vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_float), _i32p]
        L.ref_immature_create.argtypes = [_vp, C.c_int, C.c_int, C.c_float]; L.ref_immature_create.restype = _vp
        L.ref_immature_destroy.argtypes = [_vp]; L.ref_immature_destroy.restype = None
        L.ref_immature_get.argtypes = [_vp, _f32p, C.POINTER(C.c_int)]; L.ref_immature_get.restype = None
        L.ref_immature_set_range.argtypes = [_vp, C.c_float, C.c_float, C.c_int]; L.ref_immature_set_range.restype = None
        L.ref_immature_trace.argtypes = [_vp, _vp, _f32p, _f32p, _f32p]
        L.ref_undistort_create.argtypes = [C.c_char_p, _i32p, _i32p, _f64p]; L.ref_undistort_create.restype = _vp
        L.ref_undistort_destroy.argtypes = [_vp]; L.ref_undistort_destroy.restype = None
        L.ref_undistort_maps.argtypes = [_vp, _f32p, _f32p]; L.ref_undistort_maps.restype = None
        L.ref_undistort_passthrough.argtypes = [_vp]
        L.ref_undistort_apply_u8.argtypes = [_vp, _vp, C.c_float, _f32p]; L.ref_undistort_apply_u8.restype = None
        _LIB = L
    return _LIB


class _Quiet:
    """setGlobalCalib printf()s: keep the reference's stdout chatter out of test output"""
    def __enter__(self):
        import sys
        sys.stdout.flush(); self.fd = os.dup(1); dn = os.open(os.devnull, os.O_WRONLY); os.dup2(dn, 1); os.close(dn)
    def __exit__(self, *a):
        C.CDLL(None).fflush(None); os.dup2(self.fd, 1); os.close(self.fd)


def set_calib(w, h, K) -> int:
    """setGlobalCalib + CalibHessian: returns pyrLevelsUsed.  Process-global, like the reference."""
    with _Quiet():
        return lib().ref_set_calib(w, h, *[float(k) for k in K])


def settings(huberTH=6.0, coarseCutoffTH=20.0, affA=0.0, affB=0.0):
    lib().ref_settings(huberTH, coarseCutoffTH, affA, affB)


def global_K(lvl):
    a = np.zeros(4, np.float32); b = np.zeros(4, np.float32); lib().ref_get_global_K(lvl, a, b); return a, b


class Frame:
    """FrameHessian after makeImages (HessianBlocks.cpp:107-167)."""

    def __init__(self, color, wh, levels, exposure: float = 1.0):
        self.w, self.h = wh; self.levels = levels
        color = np.ascontiguousarray(color, np.float32); assert color.shape == (self.h, self.w)
        self.p = lib().ref_frame_create(color, exposure)

    def dI(self, lvl):
        w, h = self.w >> lvl, self.h >> lvl
        return np.ctypeslib.as_array(lib().ref_frame_dI(self.p, lvl), shape=(h, w, 3)).copy()

    def absSquaredGrad(self, lvl):
        w, h = self.w >> lvl, self.h >> lvl
        return np.ctypeslib.as_array(lib().ref_frame_abs(self.p, lvl), shape=(h, w)).copy()

    def __del__(self):
        if getattr(self, "p", None) and _LIB is not None:
            _LIB.ref_frame_destroy(self.p); self.p = None


class CoarseTracker:
    """The reference's CoarseTracker (FullSystem/CoarseTracker.cpp) behind the flat signatures of orc.CoarseTracker."""

    def __init__(self):
        self.p = lib().ref_tracker_create(); self._keep = None

    def K(self, lvl):
        o = np.zeros(4, np.float32); lib().ref_tracker_get_K(self.p, lvl, o); return o

    def Ki(self, lvl):
        o = np.zeros(9, np.float32); lib().ref_tracker_get_Ki(self.p, lvl, o); return o.reshape(3, 3)

    def setCoarseTrackingRef(self, ref: Frame, pts, round_half, ref_a=0.0, ref_b=0.0, old: Frame | None = None):
        """pts (n,4) {u,v,idepth,HdiF}; rows with round_half != 0 (points of older keyframes, need `old`) must come first."""
        self._keep = (ref, old)
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 4); rh = np.ascontiguousarray(round_half, np.int32)
        rc = lib().ref_tracker_set_ref(self.p, ref.p, None if old is None else old.p, pts, rh, len(pts), ref_a, ref_b)
        if rc != 0:
            raise ValueError("round_half rows must precede the direct rows and need an `old` frame")

    def cloud(self, lvl):
        n = lib().ref_tracker_cloud_n(self.p, lvl); a = [np.zeros(max(n, 1), np.float32) for _ in range(4)]
        lib().ref_tracker_get_cloud(self.p, lvl, *a); return [x[:n] for x in a]

    def calcRes(self, new: Frame, lvl, T7, a, b, cutoff):
        rs = np.zeros(6); lib().ref_tracker_calc_res(self.p, new.p, lvl, np.ascontiguousarray(T7, np.float64), a, b, cutoff, rs); return rs

    def warped(self):
        n = lib().ref_tracker_warped_n(self.p); o = np.zeros((8, max(n, 1)), np.float32)
        if n:
            o = np.zeros((8, n), np.float32); lib().ref_tracker_get_warped(self.p, o)
            return o
        return o[:, :0]

    def calcGSSSE(self, lvl, T7, a, b):
        H = np.zeros(64); bb = np.zeros(8); lib().ref_tracker_calc_gs(self.p, lvl, np.ascontiguousarray(T7, np.float64), a, b, H, bb); return H.reshape(8, 8), bb

    def trackNewestCoarse(self, new: Frame, T7, ab, coarsest, minRes=None):
        T = np.array(T7, np.float64); abv = np.array(ab, np.float64)
        minRes = np.full(5, np.nan) if minRes is None else np.ascontiguousarray(minRes, np.float64)
        lastRes = np.zeros(5); flow = np.zeros(3)
        with _Quiet():
            good = lib().ref_tracker_track(self.p, new.p, T, abv, coarsest, minRes, lastRes, flow)
        return dict(good=bool(good), T=T, ab=abv, lastResiduals=lastRes, flow=flow)

    def __del__(self):
        if getattr(self, "p", None) and _LIB is not None:
            _LIB.ref_tracker_destroy(self.p); self.p = None


def interp33(dI3, x, y, bilin=False):
    dI3 = np.ascontiguousarray(dI3, np.float32); o = np.zeros(3, np.float32)
    (lib().ref_interp33_bilin if bilin else lib().ref_interp33)(dI3.reshape(-1), dI3.shape[1], x, y, o); return o


def aff_from_to(eF, eT, aF, bF, aT, bT):
    o = np.zeros(2); lib().ref_aff_from_to(eF, eT, aF, bF, aT, bT, o); return o


class BAWindow:
    """The flattened sliding window of synth.make_ba_window() loaded into the reference's own FullSystem / EnergyFunctional (ref_shim.cpp ref_ba_*).
    Same methods as orc.BAWindow.  `frames`: ref.Frame objects of the window's keyframes (set_calib must match their size)."""

    def __init__(self, win: dict, frames):
        L = lib(); self.L = L; self.win = win; self._frames = frames
        self.nF = win["nF"]; self.nP = len(win["uv"]); self.nR = len(win["r_point"]); self.n = 4 + 6 * self.nF
        with _Quiet():
            self.p = L.ref_ba_create()
        L.ref_ba_set_calib(self.p, np.ascontiguousarray(win["K"], np.float64))
        for i in range(self.nF):
            L.ref_ba_add_frame(self.p, frames[i].p, np.ascontiguousarray(win["T_eval"][i]), np.ascontiguousarray(win["state"][i]), np.ascontiguousarray(win["state_zero"][i]),
                               float(win["ab_exposure"][i]), int(win["frameID"][i]), float(win["frameEnergyTH"][i]))
        c = lambda k, t: np.ascontiguousarray(win[k], t)
        L.ref_ba_set_points(self.p, self.nP, c("uv", np.float32), c("idepth", np.float32), c("idepth_zero", np.float32), c("color", np.float32), c("weights", np.float32),
                            c("host", np.int32), c("hasDepthPrior", np.int32), c("isFromSensor", np.int32), c("res_begin", np.int32))
        L.ref_ba_set_residuals(self.p, self.nR, c("r_point", np.int32), c("r_host", np.int32), c("r_target", np.int32), c("r_hasMatcher", np.int32), c("r_matcher", np.float32), c("r_isNew", np.int32))
        L.ref_ba_set_prior(self.p, c("HM", np.float64), c("bM", np.float64))
        L.ref_ba_init(self.p)

    def reset_oob(self): self.L.ref_ba_reset_oob(self.p)
    def linearizeAll(self, fix=False):
        with _Quiet():
            return self.L.ref_ba_linearize_all(self.p, 1 if fix else 0)
    def applyRes(self): self.L.ref_ba_apply_res(self.p)
    def calcLEnergy(self): return self.L.ref_ba_energy_L(self.p)
    def calcMEnergy(self): return self.L.ref_ba_energy_M(self.p)
    def backupState(self): self.L.ref_ba_backup(self.p)
    def doStepFromBackup(self, f=1.0): return bool(self.L.ref_ba_do_step(self.p, f))
    def loadStateBackup(self): self.L.ref_ba_load_backup(self.p)

    def residuals(self):
        n = self.nR
        o = dict(state=np.zeros(n, np.int32), new_state=np.zeros(n, np.int32), energies=np.zeros((n, 3)), active=np.zeros(n, np.int32), J=np.zeros((n, 24), np.float32),
                 efJ=np.zeros((n, 24), np.float32), JpJdF=np.zeros((n, 8), np.float32), center=np.zeros((n, 3), np.float32), isLinearized=np.zeros(n, np.int32))
        self.L.ref_ba_get_residuals(self.p, o["state"], o["new_state"], o["energies"], o["active"], o["J"], o["efJ"], o["JpJdF"], o["center"], o["isLinearized"]); return o

    def accumulate(self):
        n = self.L.ref_ba_dim(self.p); HA = np.zeros((n, n)); bA = np.zeros(n); Hsc = np.zeros((n, n)); bsc = np.zeros(n)
        self.L.ref_ba_accumulate(self.p, HA, bA, Hsc, bsc); return HA, bA, Hsc, bsc

    def solveSystem(self, iteration, lam):
        n = self.L.ref_ba_dim(self.p); x = np.zeros(n); HS = np.zeros((n, n)); bS = np.zeros(n)
        with _Quiet():
            self.L.ref_ba_solve(self.p, iteration, lam, x, HS, bS)
        return x, HS, bS

    def points(self):
        n = self.nP
        o = dict(idepth=np.zeros(n, np.float32), step=np.zeros(n, np.float32), HdiF=np.zeros(n, np.float32), bdSumF=np.zeros(n, np.float32), maxRelBaseline=np.zeros(n, np.float32),
                 numGood=np.zeros(n, np.int32), idepth_hessian=np.zeros(n, np.float32))
        self.L.ref_ba_get_points(self.p, o["idepth"], o["step"], o["HdiF"], o["bdSumF"], o["maxRelBaseline"], o["numGood"], o["idepth_hessian"]); return o

    def frames(self):
        n = (self.L.ref_ba_dim(self.p) - 4) // 6
        o = dict(T_eval=np.zeros((n, 7)), state=np.zeros((n, 10)), step=np.zeros((n, 10)), frameEnergyTH=np.zeros(n, np.float32), PRE_worldToCam=np.zeros((n, 7)))
        self.L.ref_ba_get_frames(self.p, o["T_eval"], o["state"], o["step"], o["frameEnergyTH"], o["PRE_worldToCam"]); return o

    def calib(self):
        v = np.zeros(4); s = np.zeros(4); self.L.ref_ba_get_calib(self.p, v, s); return v, s

    def precalc(self, host, target):
        o = np.zeros(27, np.float32); aH = np.zeros(36); aT = np.zeros(36); d = np.zeros(6, np.float32)
        self.L.ref_ba_get_precalc(self.p, host, target, o, aH, aT, d)
        return dict(KRKi=o[:9].reshape(3, 3), Kt=o[9:12], R0=o[12:21].reshape(3, 3), t0=o[21:24], aff=o[24:26], b0=o[26], adHost=aH.reshape(6, 6), adTarget=aT.reshape(6, 6), adHTdelta=d)

    def immature_pre(self, host, nF):
        """(pre14 (nF-1,14): PRE_RTll, PRE_tTll, PRE_aff_mode of (host, every other frame in window order), calib6: fxl fyl cxl cyl fxli fyli)"""
        pre = np.zeros((nF - 1, 14), np.float32); cal = np.zeros(6, np.float32); k = self.L.ref_ba_immature_pre(self.p, host, pre.reshape(-1), cal); assert k == nF - 1
        return pre, cal

    def optimizeImmaturePoint(self, host, u, v, idepth_min, idepth_max, is_from_sensor, min_obs, nF):
        """FullSystem::optimizeImmaturePoint on a candidate constructed at (u,v) of keyframe `host` -> (status 0/-1/1, idepth, final temporary-residual states)"""
        idp = C.c_float(0); rs = np.zeros(nF - 1, np.int32)
        with _Quiet():
            st = self.L.ref_ba_optimize_immature(self.p, host, int(u), int(v), idepth_min, idepth_max, int(is_from_sensor), min_obs, C.byref(idp), rs)
        return st, idp.value, rs

    def optimize(self, its=6):
        with _Quiet():
            rmse = self.L.ref_ba_optimize(self.p, its)
        return dict(rmse=float(rmse))

    def flagPointsForRemoval(self, selected):
        st = np.zeros(self.nP, np.int32)
        with _Quiet():
            self.L.ref_ba_flag_points(self.p, np.ascontiguousarray(selected, np.int32), st)
        return st

    def marginalizePointsF(self):
        with _Quiet():
            self.L.ref_ba_marginalize_points(self.p)

    def marginalizeFrame(self, idx):
        with _Quiet():
            self.L.ref_ba_marginalize_frame(self.p, int(idx))

    def prior(self):
        n = self.L.ref_ba_dim(self.p); HM = np.zeros((n, n)); bM = np.zeros(n); self.L.ref_ba_get_prior(self.p, HM, bM); return HM, bM

    def res_to_zero(self):
        r = np.zeros((self.nR, 2), np.float32); l = np.zeros(self.nR, np.int32); self.L.ref_ba_get_res_to_zero(self.p, r, l); return r, l

    def __del__(self):
        if getattr(self, "p", None) and _LIB is not None:
            with _Quiet():
                _LIB.ref_ba_destroy(self.p)
            self.p = None


class System:
    """The reference's whole vision pipeline: FullSystem::addActiveFrame per frame (FullSystem.cpp:822-900), fed like main.cpp:466-509 feeds it.
    BASELINE.json config #1 (single sequence, CPU reference path, pose + energy dump) runs through this class.  set_calib() first."""

    def __init__(self, levels: int, perfect_images: bool = True, seed: int = 3141592):
        self.levels = levels
        with _Quiet():
            self.p = lib().ref_sys_create(1 if perfect_images else 0)
        lib().ref_srand(seed)                                  # PixelSelector2.cpp:15 does srand(3141592); rand() is then consumed by the Reprojector shuffle and makeNewTraces

    def srand(self, seed: int): lib().ref_srand(seed)

    def addActiveFrame(self, image, cloud_px, timestamp: float, exposure: float = 1.0) -> int:
        """image (h,w) float32 0..255; cloud_px (n,3) {Ku, Kv, depth} of the LiDAR sweep in the cropped image (main.cpp:810-855).  Returns 0, -1 lost, -2 init failed."""
        img = np.ascontiguousarray(image, np.float32); cl = np.ascontiguousarray(cloud_px, np.float64).reshape(-1, 3)
        with _Quiet():
            return lib().ref_sys_add_frame(self.p, img, exposure, timestamp, cl, len(cl))

    # ---- PixelSelector / makeNewTraces state of the running system (tests/test_sequence_select.py)
    def selector_zero(self): lib().ref_sys_selector_zero.argtypes = [_vp]; lib().ref_sys_selector_zero(self.p)
    def selector_state(self, wh, with_map=True):
        L = lib(); L.ref_sys_selector_state.argtypes = [_vp, _vp]; m = np.zeros((wh[1], wh[0]), np.float32) if with_map else None
        return L.ref_sys_selector_state(self.p, m.ctypes.data if with_map else None), m
    def set_selection_map(self, m): lib().ref_sys_set_selection_map.argtypes = [_vp, _f32p]; lib().ref_sys_set_selection_map(self.p, np.ascontiguousarray(m, np.float32).reshape(-1))
    def set_lidar_state(self, lrud, addFeaturePoint): lib().ref_sys_set_lidar_state.argtypes = [_vp, _i32p, C.c_int]; lib().ref_sys_set_lidar_state(self.p, np.ascontiguousarray(lrud, np.int32), int(addFeaturePoint))
    def newest_kf_immature(self, cap=1 << 16):
        L = lib(); L.ref_sys_newest_kf_immature.argtypes = [_vp, _f32p, C.c_int, C.POINTER(C.c_int)]; o = np.zeros((cap, 7), np.float32); k = C.c_int(-1)
        m = L.ref_sys_newest_kf_immature(self.p, o.reshape(-1), cap, C.byref(k)); assert m <= cap
        return (o[:m], k.value) if m >= 0 else (None, -1)

    def probe_new_traces(self, frame: "Frame", cloud3, cap=1 << 16):
        """FullSystem::makeNewTraces of the running system on `frame` with its live selector state, state restored afterwards -> rows {u,v,my_type,score,idepth_fromSensor,isFromSensor,type}"""
        L = lib(); L.ref_sys_probe_new_traces.argtypes = [_vp, _vp, _vp, C.c_int, _f32p, C.c_int, C.c_int]; c = np.ascontiguousarray(cloud3, np.float64).reshape(-1, 3); o = np.zeros((cap, 7), np.float32)
        with _Quiet():
            m = L.ref_sys_probe_new_traces(self.p, frame.p, c.ctypes.data, len(c), o.reshape(-1), cap, 0)
        assert m <= cap
        return o[:m]

    def immature_dump(self, cap=1 << 16):
        """every ImmaturePoint of every keyframe of the window -> list of (shell id, records (n,29) float32, status (n,) int32), window order"""
        L = lib(); L.ref_sys_immature_dump.argtypes = [_vp, _i32p, _i32p, _f32p, _i32p, C.c_int]; ids = np.zeros(16, np.int32); cnt = np.zeros(16, np.int32)
        rec = np.zeros((cap, 29), np.float32); st = np.zeros(cap, np.int32); nk = L.ref_sys_immature_dump(self.p, ids, cnt, rec.reshape(-1), st, cap); assert nk >= 0
        out = []; o = 0
        for k in range(nk): out.append((int(ids[k]), rec[o:o + cnt[k]].copy(), st[o:o + cnt[k]].copy())); o += int(cnt[k])
        return out
    def trace_geometry(self, host_idx, new_c2w7):
        L = lib(); L.ref_sys_trace_geometry.argtypes = [_vp, C.c_int, _f64p, _f32p, _f32p]; a = np.zeros(9, np.float32); b = np.zeros(3, np.float32)
        L.ref_sys_trace_geometry(self.p, host_idx, np.ascontiguousarray(new_c2w7, np.float64), a, b); return a.reshape(3, 3), b

    def export_window(self, wh):
        """the LIVE sliding window flattened into the layout of synth.make_ba_window (frames / points / residuals in EnergyFunctional order) + shell ids"""
        L = lib(); sz = np.zeros(3, np.int32); L.ref_sys_window_sizes.argtypes = [_vp, _i32p]; L.ref_sys_window_sizes(self.p, sz); nF, nP, nR = [int(x) for x in sz]; n = 4 + 6 * nF
        f32 = lambda *sh: np.zeros(sh, np.float32); i32 = lambda *sh: np.zeros(sh, np.int32); f64 = lambda *sh: np.zeros(sh, np.float64)
        a = dict(shell_ids=i32(nF), T_eval=f64(nF, 7), state=f64(nF, 10), state_zero=f64(nF, 10), ab_exposure=f32(nF), frameID=i32(nF), frameEnergyTH=f32(nF), K=f64(4), K_zero=f64(4),
                 uv=f32(nP, 2), idepth=f32(nP), idepth_zero=f32(nP), color=f32(nP, 8), weights=f32(nP, 8), host=i32(nP), hasDepthPrior=i32(nP), isFromSensor=i32(nP), res_begin=i32(nP + 1),
                 r_point=i32(nR), r_host=i32(nR), r_target=i32(nR), r_hasMatcher=i32(nR), r_matcher=f32(nR, 2), r_isNew=i32(nR), HM=f64(n, n), bM=f64(n))
        L.ref_sys_export_window.argtypes = [_vp] + [C.c_void_p] * 26
        L.ref_sys_export_window(self.p, *[a[k].ctypes.data for k in ("shell_ids", "T_eval", "state", "state_zero", "ab_exposure", "frameID", "frameEnergyTH", "K", "K_zero", "uv", "idepth", "idepth_zero",
                                                                     "color", "weights", "host", "hasDepthPrior", "isFromSensor", "res_begin", "r_point", "r_host", "r_target", "r_hasMatcher", "r_matcher",
                                                                     "r_isNew", "HM", "bM")])
        a.update(nF=nF, wh=tuple(wh), kf_idx=[int(x) for x in a["shell_ids"]]); return a
    def optimize(self, its=6):
        L = lib(); L.ref_sys_optimize.restype = C.c_float; L.ref_sys_optimize.argtypes = [_vp, C.c_int]
        with _Quiet():
            return L.ref_sys_optimize(self.p, its)

    def num_frames(self): return lib().ref_sys_num_frames(self.p)
    def num_keyframes(self): return lib().ref_sys_num_keyframes(self.p)

    def frame(self, i):
        T = np.zeros(7); ab = np.zeros(2); fl = np.zeros(3, np.int32); lib().ref_sys_get_frame(self.p, i, T, ab, fl)
        c2r = np.zeros(7); lib().ref_sys_get_track_result(self.p, i, c2r)
        return dict(camToWorld=T, aff_g2l=ab, poseValid=bool(fl[0]), isKeyframe=bool(fl[1]), trackingRef=int(fl[2]), camToTrackingRef=c2r)

    def lastCoarseRMSE(self):
        o = np.zeros(5); lib().ref_sys_get_last_rmse(self.p, o); return o

    def window(self):
        ids = np.zeros(16, np.int32); n = np.zeros(16, np.int32); k = lib().ref_sys_window(self.p, ids, n); return ids[:k].copy(), n[:k].copy()

    def tracker_snapshot(self):
        """State trackNewCoarse will see for the NEXT frame: reference clouds of the tracker that will be used, pose history, active map."""
        rid = C.c_int(0); ab = np.zeros(2); ex = C.c_float(0); n = np.zeros(6, np.int32); fr = C.c_double(0)
        if lib().ref_sys_tracker_info(self.p, C.byref(rid), ab, C.byref(ex), n, C.byref(fr)) != 0:
            return None
        clouds = []
        for l in range(self.levels):
            a = [np.zeros(max(int(n[l]), 1), np.float32) for _ in range(4)]; lib().ref_sys_tracker_cloud(self.p, l, *a); clouds.append([x[:n[l]].copy() for x in a])
        sp = np.zeros(7); sl = np.zeros(7); lf = np.zeros(7); al = np.zeros(2); rm = np.zeros(5); nh = C.c_int(0)
        lib().ref_sys_history(self.p, sp, sl, lf, al, rm, C.byref(nh))
        nkf = C.c_int(0); npts = lib().ref_sys_map_size(self.p, C.byref(nkf)); k = nkf.value
        ids = np.zeros(max(k, 1), np.int32); kT = np.zeros((max(k, 1), 7)); kab = np.zeros((max(k, 1), 2)); kex = np.zeros(max(k, 1), np.float32); p5 = np.zeros((max(npts, 1), 5), np.float32)
        lib().ref_sys_map(self.p, ids, kT, kab, kex, p5)
        K4 = np.zeros(4, np.float32); cal = np.zeros(4); lib().ref_sys_tracker_K(self.p, K4, cal)
        return dict(tracker_K=K4, calib=cal, ref_frame=rid.value, ref_ab=ab, ref_exposure=ex.value, clouds=clouds, firstCoarseRMSE=fr.value, sprelast=sp, slast=sl, lastF=lf, aff_last=al, lastCoarseRMSE=rm,
                    n_history=nh.value, kf_ids=ids[:k], kf_T7=kT[:k], kf_ab=kab[:k], kf_exposure=kex[:k], map_pts=p5[:npts])

    def __del__(self):
        try:
            if getattr(self, "p", None) and _LIB is not None:
                with _Quiet():
                    _LIB.ref_sys_destroy(self.p)
                self.p = None
        except Exception:                                                     # interpreter shutdown: the mapping thread dies with the process
            pass


def libc_rand_shuffle(n: int):
    """std::random_shuffle(first, last) of libstdc++ on 0..n-1, driven by the C library's rand() in THIS process (the reference's Reprojector grid, Reprojector.cpp:107).
    Call right after srand(seed) to learn the cell order a Reprojector constructed next will use; then srand(seed) again."""
    libc = C.CDLL(None); a = list(range(n))
    for i in range(1, n):
        j = libc.rand() % (i + 1)
        if i != j:
            a[i], a[j] = a[j], a[i]
    return np.array(a, np.int32)


def reproject_map(wh, kf_frames, kf_T7, kf_ab, cur_frame, cur_T7, cur_ab, pts, seed=1, refine=False):
    """Reprojector::reprojectMap (Reprojector.cpp:117-156) [+ structPoseEstimation] of the reference on flat inputs; pts: structured array u, v, idepth, host, type.
    Returns (pt_index, px, cell_order used, refined camToWorld or None)."""
    L = lib(); nH = len(kf_frames); fr = (_vp * nH)(*[f.p for f in kf_frames])
    p5 = np.ascontiguousarray(np.stack([pts["u"], pts["v"], pts["idepth"], pts["host"].astype(np.float32), pts["type"].astype(np.float32)], 1), np.float32)
    ncells = int(np.ceil(wh[0] / 25.0)) * int(np.ceil(wh[1] / 25.0))
    L.ref_srand(seed); order = libc_rand_shuffle(ncells)
    out_pt = np.zeros(ncells, np.int32); out_px = np.zeros((ncells, 2)); T = np.array(cur_T7, np.float64).copy()
    with _Quiet():
        n = L.ref_reproject_map(nH, fr, np.ascontiguousarray(kf_T7, np.float64), np.ascontiguousarray(kf_ab, np.float64), cur_frame.p, T, np.ascontiguousarray(cur_ab, np.float64),
                                len(p5), p5, seed, 1 if refine else 0, out_pt, out_px)
    return out_pt[:n].copy(), out_px[:n].copy(), order, (T if refine else None)


class Undistort:
    """The reference's Undistort object for a calibration text (Undistort::getUndistorterForFile, util/Undistort.cpp:232-334): K, remap tables, undistort<unsigned char>."""

    def __init__(self, config_text: str):
        L = lib(); who = np.zeros(2, np.int32); wh = np.zeros(2, np.int32); K4 = np.zeros(4)
        with _Quiet():
            self.p = L.ref_undistort_create(config_text.encode(), who, wh, K4)
        if not self.p:
            raise RuntimeError("the reference rejected the calibration text")
        self.wOrg, self.hOrg = int(who[0]), int(who[1]); self.w, self.h = int(wh[0]), int(wh[1]); self.K4d = K4
        self.remapX = np.zeros((self.h, self.w), np.float32); self.remapY = np.zeros((self.h, self.w), np.float32)
        L.ref_undistort_maps(self.p, self.remapX.reshape(-1), self.remapY.reshape(-1))
        self.passthrough = bool(L.ref_undistort_passthrough(self.p))

    def undistort(self, raw_u8, exposure: float = 1.0):
        raw = np.ascontiguousarray(raw_u8, np.uint8); assert raw.shape == (self.hOrg, self.wOrg)
        out = np.zeros((self.h, self.w), np.float32)
        with _Quiet():
            lib().ref_undistort_apply_u8(self.p, raw.ctypes.data, exposure, out.reshape(-1))
        return out

    def __del__(self):
        try:
            if getattr(self, "p", None):
                lib().ref_undistort_destroy(self.p); self.p = None
        except Exception:
            pass


class ImmaturePoint:
    """The reference's ImmaturePoint (FullSystem/ImmaturePoint.cpp): constructor on a host Frame, traceOn against another Frame."""

    def __init__(self, host: "Frame", u: int, v: int, my_type: float = 1.0):
        self.host = host; self.p = lib().ref_immature_create(host.p, int(u), int(v), my_type)

    def record(self):
        """(29 floats in the layout of orc.IMM_DTYPE's float fields, status)"""
        o = np.zeros(29, np.float32); st = C.c_int(0); lib().ref_immature_get(self.p, o, C.byref(st)); return o, st.value

    def set_range(self, idepth_min, idepth_max, status):
        lib().ref_immature_set_range(self.p, idepth_min, idepth_max, status)

    def traceOn(self, frame: "Frame", KRKi, Kt, aff) -> int:
        return lib().ref_immature_trace(self.p, frame.p, np.ascontiguousarray(KRKi, np.float32).reshape(-1), np.ascontiguousarray(Kt, np.float32), np.ascontiguousarray(aff, np.float32))

    def __del__(self):
        try:
            if getattr(self, "p", None):
                lib().ref_immature_destroy(self.p); self.p = None
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------- candidate management: PixelSelector, makeNewTraces, CoarseDistanceMap (the reference's own)
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def _sel_protos(L):
    L.ref_selector_create.restype = _vp; L.ref_selector_create.argtypes = []; L.ref_selector_destroy.argtypes = [_vp]; L.ref_selector_destroy.restype = None
    L.ref_selector_random_pattern.argtypes = [_vp, _u8p]; L.ref_selector_set_potential.argtypes = [_vp, C.c_int]; L.ref_selector_get_potential.argtypes = [_vp]
    L.ref_selector_make_hists.argtypes = [_vp, _vp, _f32p, _f32p]
    L.ref_selector_select.argtypes = [_vp, _vp, _f32p, C.c_int, C.c_float, _vp, C.c_int, _i32p]
    L.ref_selector_make_maps.argtypes = [_vp, _vp, _f32p, C.c_float, C.c_int, C.c_float, _vp, C.c_int]
    L.ref_ba_selector.restype = _vp; L.ref_ba_selector.argtypes = [_vp]
    L.ref_shi_tomasi.restype = C.c_float; L.ref_shi_tomasi.argtypes = [_vp, _vp, C.c_int, C.c_int]
    L.ref_make_new_traces.argtypes = [_vp, _vp, _vp, C.c_int, _i32p, C.c_int, C.c_float, _f32p, _f32p, C.c_int]
    L.ref_distmap_make.argtypes = [_vp, C.c_int]; L.ref_distmap_add.argtypes = [_vp, C.c_int, C.c_int]; L.ref_distmap_get.argtypes = [_vp, _f32p]
    L.ref_distmap_geometry.argtypes = [_vp, C.c_int, C.c_int, _f32p, _f32p]
    L.ref_activate_select.argtypes = [_vp, C.c_int, C.c_int, _i32p, _i32p, _f32p, C.c_float, _i32p]
    return L


class Selector:
    """The reference's PixelSelector (own object, or the one inside a BAWindow's FullSystem when `owner` is given)."""

    def __init__(self, wh, owner=None):
        L = _sel_protos(lib()); self.w, self.h = wh; self.owner = owner
        self.p = L.ref_ba_selector(owner.p) if owner is not None else L.ref_selector_create()

    def randomPattern(self):
        o = np.zeros(self.w * self.h, np.uint8); lib().ref_selector_random_pattern(self.p, o); return o

    @property
    def currentPotential(self): return lib().ref_selector_get_potential(self.p)
    @currentPotential.setter
    def currentPotential(self, v): lib().ref_selector_set_potential(self.p, int(v))

    def makeHists(self, frame):
        n = (self.w // 32) * (self.h // 32); a = np.zeros(n, np.float32); b = np.zeros(n, np.float32); lib().ref_selector_make_hists(self.p, frame.p, a, b); return a, b

    def select(self, frame, pot, thFactor=1.0, cloud3=None):
        c = None if cloud3 is None else np.ascontiguousarray(cloud3, np.float64).reshape(-1, 3)
        m = np.zeros(self.w * self.h if c is None else max(len(c), 1), np.float32); n3 = np.zeros(3, np.int32)
        lib().ref_selector_select(self.p, frame.p, m, pot, thFactor, None if c is None else c.ctypes.data, 0 if c is None else len(c), n3)
        return (m.reshape(self.h, self.w) if c is None else m[:len(c)]), n3

    def makeMaps(self, frame, density, recursionsLeft=1, thFactor=1.0, cloud3=None):
        c = None if cloud3 is None else np.ascontiguousarray(cloud3, np.float64).reshape(-1, 3)
        m = np.zeros(self.w * self.h if c is None else max(len(c), 1), np.float32)
        n = lib().ref_selector_make_maps(self.p, frame.p, m, density, recursionsLeft, thFactor, None if c is None else c.ctypes.data, 0 if c is None else len(c))
        return (m.reshape(self.h, self.w) if c is None else m[:len(c)]), n

    def __del__(self):
        if self.owner is None and getattr(self, "p", None) and _LIB is not None:
            _LIB.ref_selector_destroy(self.p); self.p = None


def shi_tomasi(ba: "BAWindow", frame: "Frame", u, v):
    return _sel_protos(lib()).ref_shi_tomasi(ba.p, frame.p, int(u), int(v))


def make_new_traces(ba: "BAWindow", frame: "Frame", cloud3, lrud, addFeaturePoint, desiredImmatureDensity, selectionMap, cap=1 << 16):
    """FullSystem::makeNewTraces of the BAWindow's FullSystem on `frame` -> rows {u, v, my_type, score, idepth_fromSensor, isFromSensor, type}; selectionMap updated in place"""
    L = _sel_protos(lib()); c = np.ascontiguousarray(cloud3, np.float64).reshape(-1, 3); out = np.zeros((cap, 7), np.float32)
    with _Quiet():
        m = L.ref_make_new_traces(ba.p, frame.p, c.ctypes.data, len(c), np.ascontiguousarray(lrud, np.int32), int(addFeaturePoint), desiredImmatureDensity, selectionMap.reshape(-1), out.reshape(-1), cap)
    assert m <= cap
    return out[:m]


class DistMap:
    """CoarseDistanceMap of a BAWindow's FullSystem (sources: the window's ACTIVE points), newest = frame index `frame_idx`"""

    def __init__(self, ba: "BAWindow", frame_idx: int, wh):
        self.ba, self.frame_idx = ba, frame_idx; self.w1, self.h1 = wh[0] >> 1, wh[1] >> 1; _sel_protos(lib())

    def make(self): lib().ref_distmap_make(self.ba.p, self.frame_idx)
    def add(self, u, v): lib().ref_distmap_add(self.ba.p, int(u), int(v))
    def get(self):
        o = np.zeros(self.w1 * self.h1, np.float32); lib().ref_distmap_get(self.ba.p, o); return o.reshape(self.h1, self.w1)
    def geometry(self, host_idx):
        a = np.zeros(9, np.float32); b = np.zeros(3, np.float32); lib().ref_distmap_geometry(self.ba.p, host_idx, self.frame_idx, a, b); return a, b
    def activateSelect(self, host_idx, cand_begin, cand4, currentMinActDist):
        cand_begin = np.ascontiguousarray(cand_begin, np.int32); dec = np.zeros(cand_begin[-1], np.int32)
        lib().ref_activate_select(self.ba.p, self.frame_idx, len(host_idx), np.ascontiguousarray(host_idx, np.int32), cand_begin, np.ascontiguousarray(cand4, np.float32).reshape(-1), currentMinActDist, dec)
        return dec


def lidar_handler(ba: "BAWindow", xyzi, Rlc, tlc, K4, lrud, images=False):
    """The reference's own lidarCloudHandler (src/main.cpp:785-858: projectPointCloud, groundRemoval, cloudSegmentation, pixel projection) on one decoded XYZI sweep, with the
    FullSystem of `ba` as the node's global system (extrinsics, intrinsics, running pixel box).  Image size = the global calibration (set_calib)."""
    L = lib(); L.ref_lidar_handler.argtypes = [_vp, _f32p, C.c_int, _f64p, _f64p, _f32p, _i32p, _f64p, C.c_int, _i32p, _vp, _vp, _vp]; L.ref_lidar_dims.argtypes = [_i32p, _i32p]
    a = np.zeros(1, np.int32); b = np.zeros(1, np.int32); L.ref_lidar_dims(a, b); n_scan, horizon = int(a[0]), int(b[0]); m = n_scan * horizon
    xyzi = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4); out = np.zeros((m, 3)); flags = np.zeros(4, np.int32); lrud = np.ascontiguousarray(lrud, np.int32).copy()
    rng, lab, gnd = (np.zeros(m, np.float32), np.zeros(m, np.int32), np.zeros(m, np.int8)) if images else (None, None, None)
    with _Quiet():
        k = L.ref_lidar_handler(ba.p, xyzi.reshape(-1), len(xyzi), np.ascontiguousarray(Rlc, np.float64).reshape(-1), np.ascontiguousarray(tlc, np.float64), np.ascontiguousarray(K4, np.float32), lrud,
                                out.reshape(-1), m, flags, *(x.ctypes.data if x is not None else None for x in (rng, lab, gnd)))
    assert k >= 0
    r = dict(cloud_px=out[:k].copy(), lrud=lrud, addFeaturePoint=int(flags[0]), n_segmented=int(flags[3]))
    if images: r.update(range=rng.reshape(n_scan, horizon), label=lab.reshape(n_scan, horizon), ground=gnd.reshape(n_scan, horizon))
    return r
