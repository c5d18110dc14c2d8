"""ctypes driver of tests/emu/libselect_emu.so — the REAL kernels + host engine of sdv-loam_b200/csrc/sdv_select_core.cuh compiled for the host (tests/emu/cuda_emu.hpp).
TEST INFRASTRUCTURE: lets the CPU suite run the candidate-management CUDA source against the oracle in a container without a GPU."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "emu", "libselect_emu.so")
_LIB = None
NEW_TRACE_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("my_type", "<f4"), ("score", "<f4"), ("idepth_fromSensor", "<f4"), ("isFromSensor", "<i4"), ("type", "<i4")])
IMM_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("idepth_min", "<f4"), ("idepth_max", "<f4"), ("color", "<f4", 8), ("weights", "<f4", 8), ("gradH", "<f4", 4), ("energyTH", "<f4"),
                      ("quality", "<f4"), ("lastTraceUV", "<f4", 2), ("lastTracePixelInterval", "<f4"), ("lastTraceStatus", "<i4")])


def lib():
    global _LIB
    if _LIB is None:
        srcs = [os.path.join(_HERE, "emu", "select_emu.cpp"), os.path.join(_HERE, "emu", "cuda_emu.hpp"), os.path.join(_HERE, "..", "sdv-loam_b200", "csrc", "sdv_select_core.cuh"), os.path.join(_HERE, "..", "sdv-loam_b200", "csrc", "sdv_lidar_core.cuh"), os.path.join(_HERE, "..", "sdv-loam_b200", "csrc", "sdv_core_common.cuh")]
        if not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
            subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-w", "-o", _SO, srcs[0]])
        L = C.CDLL(_SO)
        L.emu_engine_create.restype = C.c_void_p; L.emu_engine_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.emu_engine_destroy.argtypes = [C.c_void_p]; L.emu_engine_error.restype = C.c_char_p; L.emu_engine_error.argtypes = [C.c_void_p]
        L.emu_engine_fuse_map.argtypes = [C.c_void_p, C.c_int]; L.emu_engine_max_scratch.argtypes = [C.c_void_p, C.c_longlong]; L.emu_engine_launches.restype = C.c_longlong; L.emu_engine_launches.argtypes = [C.c_void_p]
        L.emu_make_hists.argtypes = [C.c_void_p] * 4
        L.emu_make_maps.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 7
        L.emu_slot_create.restype = C.c_void_p; L.emu_slot_destroy.argtypes = [C.c_void_p]; L.emu_slot_set_potential.argtypes = [C.c_void_p, C.c_int]; L.emu_slot_get_potential.argtypes = [C.c_void_p]
        L.emu_slot_get_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.emu_make_new_traces.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 11 + [C.c_int] + [C.c_void_p] * 3
        L.emu_activate.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4 + [C.c_float, C.c_void_p, C.c_void_p, C.c_int]
        L.emu_atan2f.restype = C.c_float; L.emu_atan2f.argtypes = [C.c_float, C.c_float]
        L.emu_lidar_create.restype = C.c_void_p; L.emu_lidar_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]; L.emu_lidar_destroy.argtypes = [C.c_void_p]
        L.emu_lidar_error.restype = C.c_char_p; L.emu_lidar_error.argtypes = [C.c_void_p]
        L.emu_lidar_handle.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _LIB = L
    return _LIB


class FrameImgs:
    """what the device holds of a frame for this path: level-0 plane + packed {I,dx,dy,|grad|^2} texels of levels 1 and 2 (from an oracle Frame)"""
    def __init__(self, oframe):
        self.I0 = np.ascontiguousarray(oframe.dI(0)[:, :, 0]); self.L = []
        for l in (1, 2):
            d = oframe.dI(l); self.L.append(np.ascontiguousarray(np.concatenate([d, oframe.absSquaredGrad(l)[:, :, None]], 2).astype(np.float32)))


class Engine:
    def __init__(self, w, h, rp, dirDist=1):
        self.w, self.h = w, h; self.rp = np.ascontiguousarray(rp, np.uint8); self.p = lib().emu_engine_create(w, h, self.rp.ctypes.data, dirDist)

    def _ck(self, rc):
        if rc: raise RuntimeError(lib().emu_engine_error(self.p).decode())

    def makeHists(self, F):
        n = (self.w // 32) * (self.h // 32); a = np.zeros(n, np.float32); b = np.zeros(n, np.float32); self._ck(lib().emu_make_hists(self.p, F.I0.ctypes.data, a.ctypes.data, b.ctypes.data)); return a, b

    def makeMaps(self, F, pots, densities, recs, thFactors, cloud3=None):
        """a batch of makeMaps / makeMapsFromLidar calls on one frame -> maps (nj, ...), numHaveSub, final potentials, passes"""
        nj = len(pots); c = None if cloud3 is None else np.ascontiguousarray(cloud3, np.float64).reshape(-1, 3); n = 0 if c is None else len(c)
        msz = self.w * self.h if c is None else max(n, 1); maps = np.zeros((nj, msz), np.uint8); pot = np.ascontiguousarray(pots, np.int32).copy()
        num = np.zeros(nj, np.int32); passes = np.zeros(nj, np.int32); d = np.ascontiguousarray(densities, np.float32); r = np.ascontiguousarray(recs, np.int32); t = np.ascontiguousarray(thFactors, np.float32)
        self._ck(lib().emu_make_maps(self.p, nj, F.I0.ctypes.data, F.L[0].ctypes.data, F.L[1].ctypes.data, None if c is None else c.ctypes.data, n, d.ctypes.data, r.ctypes.data, t.ctypes.data,
                                     pot.ctypes.data, maps.ctypes.data, num.ctypes.data, passes.ctypes.data))
        return (maps[:, :n] if c is not None else maps.reshape(nj, self.h, self.w)), num, pot, passes

    def makeNewTraces(self, slots, Fs, clouds, densL, densD, add, cap=1 << 14):
        nj = len(slots); cl = [np.ascontiguousarray(c, np.float64).reshape(-1, 3) for c in clouds]
        arr = lambda xs: (C.c_void_p * nj)(*xs)
        out = np.zeros((nj, cap), NEW_TRACE_DTYPE); imm = np.zeros((nj, cap), IMM_DTYPE); n_out = np.zeros(nj, np.int32); num = np.zeros((nj, 2), np.int32); passes = np.zeros((nj, 2), np.int32)
        allc = np.ascontiguousarray(np.concatenate(cl)) if nj else np.zeros((0, 3)); offs = np.concatenate([[0], np.cumsum([len(c) for c in cl])])      # back to back, like the C-ABI hands them over
        n = np.array([len(c) for c in cl], np.int32); dl = np.ascontiguousarray(densL, np.float32); dd = np.ascontiguousarray(densD, np.float32); ad = np.ascontiguousarray(add, np.int32)
        self._ck(lib().emu_make_new_traces(self.p, nj, arr([s.p for s in slots]), arr([F.I0.ctypes.data for F in Fs]), arr([F.L[0].ctypes.data for F in Fs]), arr([F.L[1].ctypes.data for F in Fs]),
                                           arr([allc.ctypes.data + 24 * int(offs[j]) for j in range(nj)]), n.ctypes.data, dl.ctypes.data, dd.ctypes.data, ad.ctypes.data, out.ctypes.data, imm.ctypes.data, cap,
                                           n_out.ctypes.data, num.ctypes.data, passes.ctypes.data))
        return [out[j, :n_out[j]] for j in range(nj)], [imm[j, :n_out[j]] for j in range(nj)], num, passes

    def activate(self, pt_begin, KRKi, Kt, uvid, cand_begin=None, cKRKi=None, cKt=None, cand4=None, minActDist=0.0, copies=1):
        f = lambda a, t: np.ascontiguousarray(a, t).reshape(-1)
        pb = f(pt_begin, np.int32); A, B, Cc = f(KRKi, np.float32), f(Kt, np.float32), f(uvid, np.float32)
        nch = 0 if cand_begin is None else len(cand_begin) - 1; cb = f([0] if cand_begin is None else cand_begin, np.int32); nc = int(cb[-1])
        cA, cB, c4 = (f(x if x is not None else [0], np.float32) for x in (cKRKi, cKt, cand4))
        dec = np.zeros((copies, max(nc, 1)), np.int32); m = np.zeros((copies, self.h >> 1, self.w >> 1), np.float32)
        self._ck(lib().emu_activate(self.p, len(pb) - 1, pb.ctypes.data, A.ctypes.data, B.ctypes.data, Cc.ctypes.data, nch, cb.ctypes.data, cA.ctypes.data, cB.ctypes.data, c4.ctypes.data, minActDist,
                                    dec.ctypes.data, m.ctypes.data, copies))
        return dec[:, :nc], m


class Slot:
    def __init__(self, pot=3): self.p = lib().emu_slot_create(); lib().emu_slot_set_potential(self.p, pot)
    @property
    def currentPotential(self): return lib().emu_slot_get_potential(self.p)
    @currentPotential.setter
    def currentPotential(self, v): lib().emu_slot_set_potential(self.p, int(v))
    def map(self, w, h):
        o = np.zeros(w * h, np.uint8); lib().emu_slot_get_map(self.p, o.ctypes.data, w * h); return o.reshape(h, w)


class LidarEngine:
    """sdv_lidar_core.cuh on the host: the whole lidarCloudHandler for a batch of raw XYZI sweeps"""
    def __init__(self, n_scan=64, horizon=1800, ang_res_x=0.2, ang_res_y=0.427, ang_bottom=24.9, groundScanInd=50):
        self.n_scan, self.horizon = n_scan, horizon; self.p = lib().emu_lidar_create(n_scan, horizon, ang_res_x, ang_res_y, ang_bottom, groundScanInd)

    def handle(self, sweeps, Rlc, tlc, K4, wh, lruds):
        nj = len(sweeps); sw = [np.ascontiguousarray(s, np.float32).reshape(-1, 4) for s in sweeps]; cap = self.n_scan * self.horizon
        allp = np.ascontiguousarray(np.concatenate(sw)) if nj else np.zeros((0, 4), np.float32); offs = np.concatenate([[0], np.cumsum([len(s) for s in sw])])   # back to back, like the C-ABI hands them over
        ptrs = (C.c_void_p * nj)(*[allp.ctypes.data + 16 * int(offs[j]) for j in range(nj)]); n = np.array([len(s) for s in sw], np.int32); lr = np.ascontiguousarray(lruds, np.int32).reshape(nj, 4).copy()
        out = np.zeros((nj, cap, 3)); res = np.zeros((nj, 4), np.int32); R = np.ascontiguousarray(Rlc, np.float64).reshape(-1); t = np.ascontiguousarray(tlc, np.float64); K = np.ascontiguousarray(K4, np.float32)
        rc = lib().emu_lidar_handle(self.p, nj, ptrs, n.ctypes.data, R.ctypes.data, t.ctypes.data, K.ctypes.data, wh[0], wh[1], lr.ctypes.data, out.ctypes.data, cap, res.ctypes.data)
        if rc: raise RuntimeError(lib().emu_lidar_error(self.p).decode())
        return [dict(cloud_px=out[j, :res[j, 0]].copy(), lrud=lr[j], numGround=int(res[j, 1]), n_segmented=int(res[j, 2]), addFeaturePoint=int(res[j, 3])) for j in range(nj)]
