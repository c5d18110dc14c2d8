#!/usr/bin/env python
"""bench.py — frames/sec of the SDV-LOAM tracking hot path on B200 (BASELINE.json metric), one JSON line on stdout.

A "step" = one pass of the hot path over one batch: for each of B resident sequences (batched mode of north_star:
independent sequences / Monte-Carlo re-runs sharded over GPUs, no cross-GPU dependency inside a frame) take one new
1200x360 frame -> FrameHessian::makeImages (pyramid + gradients) -> CoarseTracker::trackNewestCoarse (coarse-to-fine
photometric SE(3)+affine LM against the keyframe's LiDAR-depth reference cloud, device-resident) -> pose/residuals back.

  value   : frames/s with the rectified frames already resident in HBM (sdv_frame_build_batch_dev + sdv_tracker_track_batch)
  e2e     : frames/s through the reference-facing C-ABI with HOST buffers: pinned RAW 1241x376 mono8 images (the sensor_msgs/Image wire format the reference
            ingests) H2D every step, rectified on the device with the reference's own Undistort tables for calib/KITTI/00.txt (crop -> 1200x360, fused into the
            pyramid kernel: sdv_frame_upload_batch_raw_u8), pose/residual D2H every step, all inside the timed region;
            e2e_float32 = host-rectified float images (FrameHessian::makeImages(float*) signature, 4x the PCIe bytes);
            e2e_trackNewCoarse = raw mono8 upload + the whole FullSystem::trackNewCoarse (sdv_track_new_coarse_batch) per frame
  roofline: the device-resident LM kernel (track_cluster_kernel): algorithmic bytes = 64 B x point evaluations
  refine  : reprojectMap + structPoseEstimation (sdv_tracker_refine_batch) on resident data; ba: FullSystem::optimize on resident 7-keyframe windows
            (sdv_ba_optimize_batch); combined: the three legs folded into one frames/s figure (tracking + refinement every frame, BA every kf_every-th)
  cpu_baseline / --impl reference: the reference's own compiled code (oracle/_ref: Undistort::undistort + makeImages + trackNewestCoarse per frame) on the host
            cores; the oracle port only where oracle/_ref is missing.
"""
from __future__ import annotations
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

WORKLOAD = "S-KITTI tracker step: %d resident sequences/GPU x 1 frame (raw 1241x376 mono8 -> Undistort crop 1200x360 of calib/KITTI/00.txt, 4 pyramid levels, %d LiDAR-depth splats -> ~10k reference points at level 0), batched mode"
KITTI00_CALIB = "Pinhole 718.856 718.856 607.1928 185.2157 0\n1241 376\ncrop\n1200 360\n"      # calib/KITTI/00.txt of the reference, verbatim (4-line data file)
K_RAW, WH_RAW = (718.856, 718.856, 607.1928, 185.2157), (1241, 376)
N_INIT_POOL = 24                 # distinct batches of initial guesses, cycled (drawing 500 x 592 guesses in Python would dominate the set-up)


def common_config(args, world, B):
    """config keys shared by the B200 arm and the reference arm (the driver compares them)"""
    return {"workload": WORKLOAD % (B, args.points), "sequences_per_gpu": B, "global_batch_frames": world * B, "batches_per_step": args.batches,
            "parallelism": "seq-shard x%d (no data-path collective)" % world,
            "init": "ground truth perturbed N(4cm, 0.002rad) (constant-motion prediction error)"}
ALG_BYTES_PER_EVAL = 64          # 16 B point + 4 texels x 12 B  
N_FRAMES = 4                     # frame 0 = keyframe, frames 1..3 tracked against it in turn


def load_sequence():
    """The bench drive as the camera delivers it: raw 1241x376 mono8 frames rendered with the real KITTI pinhole, plus what the reference's ingest makes of them —
    Undistort for calib/KITTI/00.txt (host mirror, bit-identical tables: tests/test_undistort.py): rectified K, 1200x360 float frames (undistort_host == the
    reference's undistort<unsigned char>, bit for bit), and the LiDAR pixels of the keyframe in the rectified image."""
    import types
    import sdv_loam_b200  # noqa: F401
    from sdv_loam_b200 import synth, undistort
    from conftest import cached_sequence
    und = undistort.Undistort.from_text(KITTI00_CALIB)
    rawseq = cached_sequence(N_FRAMES, 1000, K_RAW, WH_RAW)
    raw = [np.ascontiguousarray(im.astype(np.uint8)) for im in rawseq.images]
    seq = types.SimpleNamespace(K=und.K4, wh=(und.w, und.h), n=N_FRAMES, R=rawseq.R, t=rawseq.t, raw=raw, und=und, images=[und.undistort_host(r) for r in raw])
    seq.clouds = [synth.lidar_pixels(synth.World(1000), seq.R[0], seq.t[0], seq.K, seq.wh)]            # only the keyframe's cloud is used
    return seq, synth


def se3_helpers():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    return orc


def gt_and_inits(seq, synth, B, steps_total, seed=7):
    """Initial guesses = ground-truth relative pose perturbed like a constant-motion prediction error (few cm / ~0.1 deg)."""
    class orc:                                                            # data generation uses the package's own numpy SE(3) helpers, not the oracle
        se3_mul = staticmethod(synth.se3_mul7); se3_exp = staticmethod(synth.se3_exp7)
        @staticmethod
        def se3_from_rt(R, t): return np.concatenate([synth._quat_from_R(R), t])
    gts = [orc.se3_from_rt(*synth.rel_pose(seq.R[0], seq.t[0], seq.R[k], seq.t[k])) for k in range(N_FRAMES)]
    rng = np.random.default_rng(seed)
    inits = np.zeros((steps_total, B, 7))
    for s in range(steps_total):
        k = 1 + s % (N_FRAMES - 1)
        for b in range(B):
            d = np.concatenate([rng.normal(0, 0.04, 3), rng.normal(0, 0.002, 3)])
            inits[s, b] = orc.se3_mul(orc.se3_exp(d), gts[k])
    return gts, inits


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu, self.p, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "track_kernel_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------- CPU arm (oracle port)
def ref_available():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref
        return ref.available() and ref.build() is not None
    except Exception:
        return False


class RefArm:
    """The REFERENCE's own Undistort::undistort<unsigned char> + FrameHessian::makeImages + CoarseTracker::trackNewestCoarse on the raw mono8 frames (oracle/_ref/libsdvref.so = /root/reference/src compiled unmodified against stand-in
    headers, see oracle/Makefile), one independent sequence per host thread, the frame loop inside one C call per thread (ref_bench_track_loop)."""
    kind = "reference"

    def __init__(self, seq, synth, p4, threads):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref
        self.ref = ref; self.orc = se3_helpers(); self.seq, self.synth, self.threads = seq, synth, threads
        try:                                                                  # the reference allocates a fresh FrameHessian pyramid (10 MB) per frame: keep those blocks on the heap, or glibc
            libc = ctypes.CDLL("libc.so.6"); libc.mallopt(-3, 1 << 30); libc.mallopt(-1, 1 << 30)   # mmaps/unmaps each one and the threads serialise in the kernel (M_MMAP_THRESHOLD, M_TRIM_THRESHOLD)
        except OSError:
            pass
        w, h = seq.wh; self.L = ref.set_calib(w, h, seq.K); ref.settings()
        self.gts = [self.orc.se3_from_rt(*synth.rel_pose(seq.R[0], seq.t[0], seq.R[k], seq.t[k])) for k in range(N_FRAMES)]
        self.trackers = []; self.unds = []
        for i in range(threads):
            u = ref.Undistort(KITTI00_CALIB); self.unds.append(u)                                     # private: undistort<> writes into the undistorter's own buffer
            f0 = ref.Frame(u.undistort(seq.raw[0]), (w, h), self.L); tr = ref.CoarseTracker(); tr.setCoarseTrackingRef(f0, p4, np.zeros(len(p4), np.int32)); self.trackers.append((f0, tr))
        assert tuple(float(np.float32(x)) for x in self.unds[0].K4d) == tuple(seq.K)
        self.rngs = [np.random.default_rng(100 + i) for i in range(threads)]; self.count = [0] * threads
        L = ref.lib(); L.ref_bench_ingest_track_loop.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                                  np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS"), ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
        self.ptrs = (ctypes.c_void_p * (N_FRAMES - 1))(*[seq.raw[k].ctypes.data for k in range(1, N_FRAMES)])

    def run(self, frames_per_thread, budget_s=None):
        return _arm_run(self, frames_per_thread, budget_s)

    def _work(self, i, n_frames, budget_s, inits):
        done = self.ref.lib().ref_bench_ingest_track_loop(self.trackers[i][1].p, self.unds[i].p, self.ptrs, N_FRAMES - 1, self.count[i] % (N_FRAMES - 1), len(inits), inits, float(budget_s or 0.0), None, None)
        self.count[i] += done
        return done


class CpuArm:
    """The oracle's makeImages + trackNewestCoarse, one independent sequence per host thread (ctypes releases the GIL).
    Keyframe pyramids / reference clouds are built once up front, like the GPU arm does before its timed region."""
    kind = "port"

    def __init__(self, seq, synth, p4, threads):
        self.orc = orc = se3_helpers()
        try:                                                                  # keep multi-MB frame buffers on the heap: without this glibc mmaps/unmaps every
            libc = ctypes.CDLL("libc.so.6"); libc.mallopt(-3, 1 << 30); libc.mallopt(-1, 1 << 30)   # pyramid and threads serialise in the kernel (M_MMAP_THRESHOLD, M_TRIM_THRESHOLD)
        except OSError:
            pass
        self.seq, self.synth, self.threads = seq, synth, threads
        w, h = seq.wh; self.L = 4
        self.gts = [orc.se3_from_rt(*synth.rel_pose(seq.R[0], seq.t[0], seq.R[k], seq.t[k])) for k in range(N_FRAMES)]
        self.imgs = [np.ascontiguousarray(im, np.float32) for im in seq.images]
        self.trackers = []
        for i in range(threads):
            f0 = orc.Frame(self.imgs[0], self.L)
            tr = orc.CoarseTracker(w, h, self.L, seq.K); tr.setCoarseTrackingRef(f0, p4, np.zeros(len(p4), np.int32))
            self.trackers.append((f0, tr))
        self.rngs = [np.random.default_rng(100 + i) for i in range(threads)]
        self.count = [0] * threads

    def _inits(self, i, n):
        """initial guesses for the next n frames of thread i (same distribution as the GPU arm's), drawn before the clock starts"""
        orc = self.orc; rng = self.rngs[i]; out = np.zeros((n, 7))
        for f in range(n):
            k = 1 + (self.count[i] + f) % (N_FRAMES - 1)
            out[f] = orc.se3_mul(orc.se3_exp(np.concatenate([rng.normal(0, 0.04, 3), rng.normal(0, 0.002, 3)])), self.gts[k])
        return out

    def _work(self, i, n_frames, budget_s, inits):
        """n_frames x (FrameHessian::makeImages + trackNewestCoarse) inside ONE C call (orc_bench_track_loop): the thread spends its time in the CPU path,
        not in the interpreter, and the GIL is released throughout — a per-frame Python loop throttled 128 threads to a few frames/s each."""
        L = self.orc.lib(); tr = self.trackers[i][1]
        L.orc_bench_track_loop.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS"), ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
        w, h = self.seq.wh
        ptrs = (ctypes.c_void_p * (N_FRAMES - 1))(*[self.imgs[k].ctypes.data for k in range(1, N_FRAMES)])
        done = L.orc_bench_track_loop(tr.p, w, h, self.L, ptrs, N_FRAMES - 1, self.count[i] % (N_FRAMES - 1), len(inits), inits, float(budget_s or 0.0), None, None)
        self.count[i] += done
        return done

    def run(self, frames_per_thread, budget_s=None):
        return _arm_run(self, frames_per_thread, budget_s)


def _arm_inits(arm, i, n):
    orc = arm.orc; rng = arm.rngs[i]; out = np.zeros((n, 7))
    for f in range(n):
        k = 1 + (arm.count[i] + f) % (N_FRAMES - 1)
        out[f] = orc.se3_mul(orc.se3_exp(np.concatenate([rng.normal(0, 0.04, 3), rng.normal(0, 0.002, 3)])), arm.gts[k])
    return out


def _arm_run(self, frames_per_thread, budget_s=None):
    if True:
        res = [0] * self.threads
        n = min(frames_per_thread, 4096)
        inits = [_arm_inits(self, i, n) for i in range(self.threads)]
        def job(i): res[i] = self._work(i, n, budget_s, inits[i])
        th = [threading.Thread(target=job, args=(i,)) for i in range(self.threads)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        return sum(res), time.perf_counter() - t0


def ba_leg(ctx, api, synth, local_rank, W, reps=4, warm=2):
    """Back-end leg: FullSystem::optimize (6 GN iterations max) on W resident 7-keyframe windows per batch, device-resident schedule.
    Each window owns private copies of its 7 keyframes (nothing shared in L2).  Windows are re-uploaded before every repetition
    (optimize mutates them); only sdv_ba_optimize_batch is timed (CUDA events inside the library)."""
    from conftest import cached_sequence
    seq8 = cached_sequence(8, 2000, synth.KITTI_K, synth.KITTI_WH)
    wins = [synth.make_ba_window(seq8, list(range(7)), n_per_frame=250, seed=3 + i, pose_noise=(0.005, 0.0003), match_noise=0.1, prior_scale=1e-3) for i in range(4)]
    base = 1 << 41
    for wi in range(W):
        for k in range(7):
            ctx.makeImages(base + wi * 8 + k, seq8.images[k])
    ids = [[base + wi * 8 + k for k in range(7)] for wi in range(W)]
    ms = []; its = None
    for rep in range(warm + reps):
        for wi in range(W):
            api.EnergyFunctional(ctx, wins[wi % 4], ids[wi], window=wi)
        r = api.optimize_batch(ctx, list(range(W)), 6)
        if rep >= warm:
            ms.append(r["ms"]); its = r
    ms = float(np.mean(ms))
    one = []; one_wall = []                                              # latency of ONE window (BASELINE single-sequence view of the back-end): same schedule, grid of one window
    for rep in range(warm + reps):
        api.EnergyFunctional(ctx, wins[0], ids[0], window=0); ctx.sync(); t0 = time.perf_counter()
        r1 = api.optimize_batch(ctx, [0], 6); t1 = time.perf_counter()
        if rep >= warm:
            one.append(r1["ms"]); one_wall.append(1e3 * (t1 - t0))
    nR = int(np.mean([len(w["r_point"]) for w in wins])); nP = int(np.mean([len(w["uv"]) for w in wins]))
    lin_calls = float(np.mean(1 + its["iterations"] + (its["iterations"] - its["accepts"]) + 1))   # initial + per iteration + reloads + final
    return {"windows": W, "keyframes": 7, "points_per_window": nP, "residuals_per_window": nR, "ms_per_batch": ms, "windows_per_s": W / (ms * 1e-3),
            "gn_iterations_mean": float(its["iterations"].mean()), "accepts_mean": float(its["accepts"].mean()),
            "single_window_ms_device": float(np.mean(one)), "single_window_ms_wall": float(np.mean(one_wall)),
            "linearize_GBps_algorithmic": W * nR * 576 * lin_calls / (ms * 1e-3) / 1e9,
            "note": "device time of sdv_ba_optimize_batch (FullSystem::optimize, device-resident GN schedule); bit-exact vs the CPU oracle (tests/test_gpu_ba.py)"}


def refine_leg(ctx, api, synth, B, reps=6, warm=2, seed=11, cpu=True):
    """Semi-direct refinement leg = the tail of FullSystem::trackNewCoarse (FullSystem.cpp:481-488): Reprojector::reprojectMap of the active map
    (7 keyframes, ~2000 active points) into the new frame + CoarseTracker::structPoseEstimation, fused on the device (sdv_tracker_refine_batch).
    One map slot and one private target frame per sequence; the 7 keyframe images are shared by the sequences (read footprint per frame is
    ~400 10x10 patches).  Reports device time (CUDA events in the library) and wall time of the C-ABI call (job H2D + result D2H inside)."""
    from conftest import cached_sequence
    seq8 = cached_sequence(8, 2000, synth.KITTI_K, synth.KITTI_WH)
    pts, hT, hab = synth.make_map(seq8, list(range(7)), n_per_frame=300, seed=2)
    base = 1 << 42; kf_ids = [base + k for k in range(7)]
    for k in range(7):
        ctx.makeImages(kf_ids[k], seq8.images[k])
    rp = api.Reprojector(ctx); slots = np.arange(B, dtype=np.int32); ids = np.arange(B, dtype=np.uint64) + np.uint64(base + 100)
    for b in range(B):
        rp.setMap(b, kf_ids, hT, hab, pts); ctx.makeImages(int(ids[b]), seq8.images[7])
    gt = np.concatenate([synth._quat_from_R(seq8.R[7]), seq8.t[7]]); rng = np.random.default_rng(seed)
    order = rng.permutation(rp.n_cells).astype(np.int32)
    def inits():
        T = np.tile(gt, (B, 1))
        for b in range(B):
            T[b] = synth.se3_mul7(synth.se3_exp7(np.concatenate([rng.normal(0, 0.02, 3), rng.normal(0, 0.001, 3)])), gt)
        return T
    ms = []; wall = []; last = None
    for rep in range(warm + reps):
        T0 = inits(); ctx.sync(); t0 = time.perf_counter()
        r = rp.refineBatch(slots, ids, T0, cell_order=order, max_matches=400)
        t1 = time.perf_counter()
        if rep >= warm:
            ms.append(r["ms"]); wall.append(t1 - t0); last = (r, T0)
    r, T0 = last
    e0 = float(np.median(np.linalg.norm(T0[:, 4:] - gt[4:], axis=1))); e1 = float(np.median(np.linalg.norm(r["T"][:, 4:] - gt[4:], axis=1)))
    ts = [float('nan')]
    if cpu:
        # CPU port of the same stage on one core (bounded sample) — the cpu_baseline side of this leg, the only place it touches oracle/
        orc = se3_helpers(); L = 4; kf_frames = [orc.Frame(seq8.images[k], L) for k in range(7)]; cur = orc.Frame(seq8.images[7], L); w, h = synth.KITTI_WH
        ts = []
        for i in range(3):
            t0 = time.perf_counter()
            idx, px = orc.reproject_map(w, h, L, synth.KITTI_K, kf_frames, hT, hab, cur, T0[i], [0.0, 0.0], pts, cell_order=order, max_matches=400)
            p6 = np.stack([pts["u"][idx], pts["v"][idx], pts["idepth"][idx], pts["host"][idx].astype(np.float32), px[:, 0].astype(np.float32), px[:, 1].astype(np.float32)], 1).astype(np.float32)
            orc.struct_pose(w, h, np.array(synth.KITTI_K, np.float32), hT, p6, T0[i]); ts.append(time.perf_counter() - t0)
    for b in range(B):
        ctx.releaseFrame(int(ids[b]))
    return {"frames": B, "map_points": int(len(pts)), "keyframes": 7, "ms_per_batch_device": float(np.mean(ms)), "ms_per_batch_wall": 1e3 * float(np.mean(wall)),
            "frames_per_s_device": B / (float(np.mean(ms)) * 1e-3), "frames_per_s_wall": B / float(np.mean(wall)),
            "matches_mean": float(r["n_matches"].mean()), "gn_iterations_mean": float(r["iterations"].mean()), "accepts_mean": float(r["accepts"].mean()),
            "median_translation_err_in_out_m": [e0, e1], "cpu_ms_per_frame_1core": (1e3 * float(np.median(ts)) if cpu else None),
            "note": "sdv_tracker_refine_batch: reprojectMap (grid 25 px, warp-per-cell direct alignment) + structPoseEstimation, device resident; exact parity vs the CPU oracle (tests/test_gpu_reproject.py)"}


def keyframe_leg(ctx, api, synth, B, reps=3, warm=1, cpu=True):
    """Keyframe-rate candidate management for B sequences per call (SURVEY §8f ranks 3b, 4 and the caller half of 2), each stage through the C-ABI with host buffers:
    lidarCloudHandler on a raw 64-beam XYZI sweep (sdv_lidar_handler_batch) -> FullSystem::makeNewTraces on the new keyframe (sdv_make_new_traces_batch: makeHists,
    makeMapsFromLidar, makeMaps, Shi-Tomasi typing, ImmaturePoint records) -> makeDistanceMap + the activatePointsMT candidate walk (sdv_activate_select_batch).
    Wall time of the calls (H2D of sweeps / clouds / candidates and D2H of the results inside); the CPU oracle on one core beside it (bounded sample)."""
    from conftest import cached_sequence
    w, h = synth.KITTI_WH; K = synth.KITTI_K
    seq8 = cached_sequence(8, 2000, K, synth.KITTI_WH); world = synth.World(2000)
    base = 1 << 43; kf = base + 1; ctx.makeImages(kf, seq8.images[7])
    sweep = synth.lidar_sweep(world, seq8.R[7], seq8.t[7], seed=4)
    fe = api.LidarFrontEnd(ctx); rp = api.random_pattern(w, h); ps = api.PixelSelector(ctx, B, rp)
    lr0 = [[10000, -1, 10000, -1]] * B; out = {"sequences": B, "sweep_points": int(len(sweep))}
    def timed(fn):
        ts = []
        for rep in range(warm + reps):
            ctx.sync(); t0 = time.perf_counter(); r = fn(); t1 = time.perf_counter()
            if rep >= warm: ts.append(t1 - t0)
        return r, float(np.mean(ts))
    import torch
    pinned = torch.empty((B * len(sweep), 4), dtype=torch.float32, pin_memory=torch.cuda.is_available()).numpy(); pinned[:] = np.tile(sweep, (B, 1)); sb = (np.arange(B + 1) * len(sweep)).astype(np.int32)
    res, t_l = timed(lambda: fe.handle_packed(pinned, sb, synth.RLC, synth.TLC, K, lr0, cap=1 << 14))      # B raw sweeps back to back in pinned host memory, as a driver would hand them over
    cloud = res[0]["cloud_px"]; dl = api.lidar_density(res[0]["lrud"], synth.KITTI_WH, 600.0)
    out["lidar_front_end"] = {"ms_per_batch_device_kernels": ctx.last_kernel_ms(), "ms_per_batch_wall": 1e3 * t_l, "sweeps_per_s": B / t_l, "pixels_out": int(len(cloud)), "h2d_bytes_per_batch": int(B * sweep.nbytes)}
    cloud_all = np.ascontiguousarray(np.tile(cloud, (B, 1))); cbeg = (np.arange(B + 1) * len(cloud)).astype(np.int32)    # the B keyframes' pixel rows back to back
    def traces():
        for j in range(B): ps.potential(j, 3)
        return ps.makeNewTracesPacked(list(range(B)), [kf] * B, cloud_all, cbeg, dl, 600.0, 1, cap=1 << 11)
    (tr, num), t_t = timed(traces)
    out["make_new_traces"] = {"ms_per_batch_device_incl_copies": ctx.last_kernel_ms(), "ms_per_batch_wall": 1e3 * t_t, "keyframes_per_s": B / t_t, "points_per_keyframe": int(len(tr[0][0])), "lidar_monocular": [int(num[0][0]), int(num[0][1])]}
    pts, hT, hab = synth.make_map(seq8, list(range(7)), n_per_frame=300, seed=2)
    K0 = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32); K1 = np.array([[K[0] / 2, 0, (K[2] + 0.5) / 2 - 0.5], [0, K[1] / 2, (K[3] + 0.5) / 2 - 0.5], [0, 0, 1]], np.float32)
    Ki0 = np.linalg.inv(K0.astype(np.float64)).astype(np.float32); KRKi = []; Kt = []; uvid = []; pb = [0]
    for k in range(7):
        R, t = synth.rel_pose(seq8.R[k], seq8.t[k], seq8.R[7], seq8.t[7]); KRKi.append(K1 @ R.astype(np.float32) @ Ki0); Kt.append(K1 @ t.astype(np.float32))
        m = pts["host"] == k; uvid.append(np.stack([pts["u"][m], pts["v"][m], pts["idepth"][m]], 1)); pb.append(pb[-1] + int(m.sum()))
    rng = np.random.default_rng(5); cand = []; cb = [0]
    for k in range(7):
        c = seq8.clouds[k]; pick = rng.choice(len(c), 400, replace=False)
        cand.append(np.stack([np.floor(c[pick, 0]), np.floor(c[pick, 1]), 1.0 / c[pick, 2], rng.choice([1.0, 2.0, 4.0], 400)], 1)); cb.append(cb[-1] + 400)
    q = dict(pt_begin=pb, KRKi=np.stack(KRKi), Kt=np.stack(Kt), uvid=np.concatenate(uvid).astype(np.float32), cand_begin=cb, cKRKi=np.stack(KRKi), cKt=np.stack(Kt), cand4=np.concatenate(cand).astype(np.float32), minActDist=2.0)
    packed = api.packActivation([q] * B)
    dec, t_a = timed(lambda: api.activateSelectPacked(ctx, packed))
    out["activate_select"] = {"ms_per_batch_device_kernels": ctx.last_kernel_ms(), "ms_per_batch_wall": 1e3 * t_a, "sequences_per_s": B / t_a, "candidates": int(cb[-1]), "accepted": int((dec[0] == 1).sum()), "map_points": int(pb[-1])}
    if cpu:
        orc = se3_helpers(); L = 4; fo = orc.Frame(seq8.images[7], L); fe_o = orc.LidarFrontEnd(); sel_o = orc.Selector(w, h, rp); dm = orc.DistMap(w >> 1, h >> 1)
        t0 = time.perf_counter(); o = fe_o.handle(sweep, synth.RLC, synth.TLC, K, synth.KITTI_WH, lr0[0]); t1 = time.perf_counter()
        To, _, _ = sel_o.makeNewTraces(fo, cloud, dl, 600.0, 1, np.zeros((h, w), np.float32)); t2 = time.perf_counter()
        dm.make(pb, q["KRKi"], q["Kt"], q["uvid"]); do = dm.activateSelect(cb, q["cKRKi"], q["cKt"], q["cand4"], 2.0); t3 = time.perf_counter()
        out["cpu_ms_1core"] = {"lidar_front_end": 1e3 * (t1 - t0), "make_new_traces": 1e3 * (t2 - t1), "activate_select": 1e3 * (t3 - t2)}
        out["identical_to_cpu"] = bool(np.array_equal(o["cloud_px"], cloud) and To.tobytes() == tr[0][0].tobytes() and np.array_equal(do, dec[0]))
    out["note"] = "B sequences per call; every stage bit-identical to the CPU path (tests/test_gpu_select.py)"
    return out


def ba_cpu_ms(synth):
    """oracle optimize() on one host core, ms per window"""
    orc = se3_helpers()
    from conftest import cached_sequence
    seq8 = cached_sequence(8, 2000, synth.KITTI_K, synth.KITTI_WH)
    win = synth.make_ba_window(seq8, list(range(7)), n_per_frame=250, seed=3, pose_noise=(0.005, 0.0003), match_noise=0.1, prior_scale=1e-3)
    frames = [orc.Frame(seq8.images[k], 4) for k in range(7)]
    ts = []
    for _ in range(8):
        ob = orc.BAWindow(win, frames); t0 = time.perf_counter(); ob.optimize(6); ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))


def host_cores():
    """Host threads the CPU arm can really use: the affinity mask, capped by the container's CPU quota (cgroup v2 cpu.max / v1 cfs quota).  The GPU boxes of this pool
    show 128 logical CPUs but grant 16 cores of CPU time; 128 runnable threads only get throttled (measured: 970 frames/s with 128 threads)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    quota = None
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = float(q) / float(p)
    except (OSError, ValueError):
        try:
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read()); p = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / p
        except (OSError, ValueError):
            pass
    if quota is not None:
        n = max(1, min(n, int(quota + 0.5)))
    return n


def make_cpu_arm(seq, synth, p4, threads):
    """the reference's own code when oracle/_ref can be had (it travels prebuilt to the GPU box), else the oracle port"""
    return RefArm(seq, synth, p4, threads) if ref_available() else CpuArm(seq, synth, p4, threads)


def cpu_note(arm):
    if arm.kind == "reference":
        return "the reference's own sources (/root/reference/src, unmodified) compiled against stand-in Eigen/Sophus headers (oracle/_ref, g++ -O3, SSE2, no FMA); per frame: Undistort::undistort + makeImages + trackNewestCoarse"
    return "oracle/ CPU restatement (g++ -O3, no FMA) on host-rectified frames (no undistort stage): oracle/_ref is not available on this box"


def single_sequence_leg(api, synth, seq, p4, rh, local_rank, frames=240, warm=12):
    """BASELINE.json configs[1]/[2] as written: ONE sequence on one GPU.  A sequence is a chain (frame n+1 needs frame n's pose; an LM iteration needs the previous one), so
    this is a latency figure: undistort + makeImages + trackNewestCoarse per frame with the low-latency launch (256 threads, thread-block cluster of 16 CTAs, DSMEM all-gather of the
    partial sums).  device = CUDA-event time of the kernels; e2e = wall clock per frame with the raw mono8 image uploaded from pinned host memory and the pose read back."""
    import torch
    w, h = seq.wh; wo, ho = WH_RAW
    ctx = api.Context(seq.K, w, h, device=local_rank, n_tracker_slots=1, max_frames=8, cluster_size=16, track_threads=256); ctx.setUndistort(seq.und)
    KF = 1 << 40; ctx.makeImagesRaw(KF, seq.raw[0]); tr = api.CoarseTracker(ctx, 0); tr.setCoarseTrackingRef(KF, p4, rh)
    gts, inits = gt_and_inits(seq, synth, 1, N_INIT_POOL, seed=99)
    host_u8 = torch.empty((N_FRAMES - 1, ho, wo), dtype=torch.uint8).pin_memory()
    for k in range(N_FRAMES - 1):
        host_u8[k].copy_(torch.from_numpy(seq.raw[1 + k]))
    ptrs = [np.array([host_u8[k].data_ptr()], np.uint64) for k in range(N_FRAMES - 1)]
    slots = np.zeros(1, np.int32); dev_ms = []; t0 = None
    for f in range(warm + frames):
        if f == warm:
            ctx.sync(); t0 = time.perf_counter()
        ids = np.array([f & 1], np.uint64)
        ctx.makeImagesBatch(ids, ptrs[f % (N_FRAMES - 1)], raw=True)
        T = inits[(3 * (f // 3) + f % (N_FRAMES - 1)) % len(inits)].copy(); ab = np.zeros((1, 2))
        r = ctx.trackBatch(slots, ids, T, ab)
        if f >= warm:
            dev_ms.append(ctx.last_kernel_ms())
    ctx.sync(); wall = time.perf_counter() - t0
    ok = bool(r["good"][0]); ctx.close()
    return {"frames": frames, "frames_per_s_e2e": frames / wall, "ms_per_frame_e2e": 1e3 * wall / frames, "track_kernel_ms": float(np.mean(dev_ms)), "tracked_ok": ok,
            "launch": "track_cluster_kernel<256,1>, cluster of 16 CTAs per job (DSMEM), 1 job",
            "note": "one sequence = a dependent chain: latency-bound by design (SURVEY 8e: replicas only within a sequence); the batched mode above is the throughput mode"}


def stress_leg(api, local_rank):
    """BASELINE.json configs[4] (S-STRESS): 1920x1200, 5 pyramid levels, ~160k reference points at level 0 (32k active points x ~5 from the dilation), 8-keyframe window with
    32k points.  Step-wise residual/Jacobian kernel (coarse_res_gs_kernel = calcRes + calcGSSSE fused, solve on host) and the device-resident tracker on ONE job."""
    w, h = 1920, 1200; K = (1100.0, 1100.0, 959.5, 599.5)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.rint(127 + 60 * np.sin(xx / 23.0) * np.cos(yy / 31.0) + 30 * np.sin((xx + yy) / 7.0)).astype(np.float32)
    ctx = api.Context(K, w, h, device=local_rank, max_frames=3, n_tracker_slots=1, cluster_size=16, track_threads=256)
    ctx.makeImages(0, img); ctx.makeImages(1, np.roll(img, 2, axis=1)); tr = api.CoarseTracker(ctx, 0)
    rng = np.random.default_rng(0); T = np.array([1, 0, 0, 0, 0.01, 0, 0.0]); out = {"image": "1920x1200", "levels": ctx.levels}
    n = 163840
    for lvl in range(ctx.levels):
        wl, hl = w >> lvl, h >> lvl; nl = min(n, (wl - 8) * (hl - 8) // 2)
        u = rng.uniform(4, wl - 5, nl).astype(np.float32); v = rng.uniform(4, hl - 5, nl).astype(np.float32)
        order = np.lexsort((u, v.astype(np.int32))); u, v = u[order], v[order]              # raster order like makeCoarseDepthL0
        tr.setCloud(0, lvl, u, v, rng.uniform(0.02, 0.2, nl).astype(np.float32), rng.uniform(0, 255, nl).astype(np.float32))
        if lvl == 0:
            ms = []
            for rep in range(14):
                tr.calcRes(1, 0, T, 0.0, 0.0, 20.0); ms.append(ctx.last_kernel_ms())
            m = float(np.median(ms[4:]))
            out["coarse_res_gs_level0"] = {"points": int(nl), "us_per_pass": 1e3 * m, "GBps_algorithmic": ALG_BYTES_PER_EVAL * nl / (m * 1e-3) / 1e9}
    ms = []; ev = 0
    for rep in range(6):
        r = tr.trackNewestCoarse(1, np.array([1, 0, 0, 0, 0, 0, 0.0]), [0.0, 0.0]); ms.append(ctx.last_kernel_ms()); ev = int(np.sum(r["evals"]))
    m = float(np.median(ms[2:])); peak, _ = measured_peak()
    out["track_one_job"] = {"ms": m, "point_evals": ev, "GBps_algorithmic": ALG_BYTES_PER_EVAL * ev / (m * 1e-3) / 1e9, "frac_of_peak": ALG_BYTES_PER_EVAL * ev / (m * 1e-3) / 1e9 / peak,
                            "launch": "cluster of 16 CTAs x 256 threads (one job cannot fill 148 SMs: latency-bound, see DESIGN.md 4)"}
    ctx.close()
    return out


def reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores (oracle/_ref when available, else the oracle port), same workload / metric / config keys."""
    if rank != 0:
        return
    seq, synth = load_sequence()
    p4 = np.concatenate([synth.select_points(seq.images[0], seq.clouds[0], args.points), np.full((args.points, 1), 1e-3, np.float32)], 1).astype(np.float32)
    cores = host_cores(); per = 16
    arm = make_cpu_arm(seq, synth, p4, cores)
    W = max(args.warmup, 0)
    for _ in range(W):
        arm.run(2)
    steps = max(1, min(args.steps, 20)); tot_f = 0; tot_t = 0.0
    for _ in range(steps):
        f, t = arm.run(per); tot_f += f; tot_t += t
    fps = tot_f / tot_t
    cfg = common_config(args, max(1, args.gpus), args.seqs)
    cfg["note"] = "same workload definition as the B200 arm; each CPU step is a bounded sample of it: %d threads x %d frames, one sequence per thread" % (cores, per)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": arm.kind, "threads": cores, "frames_per_s_per_thread": fps / cores,
                             "sample": "%d steps x %d threads x %d frames, one sequence per thread; %s" % (steps, cores, per, cpu_note(arm))},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


METRIC = "frames/sec (KITTI 1241x376 + 64-beam): undistort + makeImages + trackNewestCoarse per frame"


_REAL_STDOUT = None


def quiet_stdout():
    """Everything any library prints on fd 1 from here on (NCCL's version banner, nvcc/ptxas chatter of a first build, ...) goes to stderr; the ONE JSON line of the contract is
    written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush(); _REAL_STDOUT = os.dup(1); os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush(); os.write(_REAL_STDOUT, data)


def bind_to_gpu_numa(local_rank):
    """Pin this rank's host threads to the CPUs of the NUMA node its GPU hangs off BEFORE the pinned staging buffers are allocated (first-touch places their pages on
    that node): at 8 ranks the H2D copies of all GPUs otherwise share whatever node the ranks happened to start on.  No-op where sysfs does not say."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return {"numa_node": None}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-"); cpus |= set(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"numa_node": node, "bound_cpus": 0}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "bound_cpus": len(cpus)}
    except Exception as e:                                                 # noqa: BLE001 — a placement hint, never a reason to fail the run
        return {"numa_node": None, "note": type(e).__name__}


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seqs", type=int, default=1184, help="resident sequences per GPU (148 SMs x 4 co-resident jobs x 2 waves)")
    ap.add_argument("--batches", type=int, default=16, help="batches (one frame of every resident sequence) per step: makes the timed region >= 1 s at the default --steps")
    ap.add_argument("--points", type=int, default=2000, help="LiDAR-depth splats of the keyframe (reference default ~1500-2000 active points)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--track-threads", type=int, default=128, help="threads per trackNewestCoarse job (128: 4 jobs/SM; 64: 8 jobs/SM; 256: latency mode)")
    ap.add_argument("--ba-windows", type=int, default=296, help="BA leg: resident 7-keyframe windows optimised per batch (0 = skip the BA leg)")
    ap.add_argument("--no-refine", action="store_true", help="skip the reprojectMap + structPoseEstimation leg")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the single-sequence and S-STRESS legs")
    ap.add_argument("--kf-every", type=int, default=5, help="keyframe cadence assumed when combining the tracker and BA legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")          # NCCL's version banner / debug lines must not land on stdout next to the JSON line
    if args.impl == "reference":
        reference_arm(args, rank, world); return
    W = max(args.warmup, 3); R = max(1, args.batches)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: the B200 path has no CPU fallback"}); sys.exit(2)
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    seq, synth = load_sequence()
    from sdv_loam_b200 import api
    w, h = seq.wh; wo, ho = WH_RAW; B = args.seqs; K = args.steps
    pts = synth.select_points(seq.images[0], seq.clouds[0], args.points)
    p4 = np.concatenate([pts, np.full((len(pts), 1), 1e-3, np.float32)], 1).astype(np.float32); rh = np.zeros(len(p4), np.int32)

    WBA = max(0, args.ba_windows)
    ctx = api.Context(seq.K, w, h, device=local_rank, n_tracker_slots=B, track_threads=args.track_threads, max_frames=3 * B + 16, max_kf_images=12)
    ctx.setUndistort(seq.und)
    KF = 1 << 40
    for b in range(B):                                                   # per sequence: keyframe -> reference cloud (makeCoarseDepthL0 on device)
        ctx.makeImages(KF, seq.images[0]); api.CoarseTracker(ctx, b).setCoarseTrackingRef(KF, p4, rh); ctx.releaseFrame(KF)
    gts, inits = gt_and_inits(seq, synth, B, N_INIT_POOL, seed=7 + rank)          # pool index s is for frame 1 + s % 3
    frames_np = np.stack(seq.images[1:]).astype(np.float32)              # (3,h,w)
    # device-resident raw inputs (value leg): one private copy per sequence, so nothing is artificially shared in L2
    dev_in = torch.empty((N_FRAMES - 1, B, h, w), dtype=torch.float32, device="cuda")
    for k in range(N_FRAMES - 1):
        dev_in[k] = torch.from_numpy(frames_np[k]).cuda()
    slots = np.arange(B, dtype=np.int32)
    ids_par = [np.arange(B, dtype=np.uint64) * 2 + p for p in (0, 1)]       # frame handles alternate between two pool slots per sequence
    stride = h * w * 4
    dev_ptrs = [np.uint64(dev_in[k].data_ptr()) + np.arange(B, dtype=np.uint64) * np.uint64(stride) for k in range(N_FRAMES - 1)]

    def frame_ids(step):
        return ids_par[step & 1]

    def init_of(step):                                                   # batch `step` shows frame 1 + step % 3: take a pool entry drawn for that frame
        return inits[(3 * (step // 3) + step % (N_FRAMES - 1)) % N_INIT_POOL]

    def step_dev(step):
        ctx.makeImagesBatch(frame_ids(step), dev_ptrs[step % (N_FRAMES - 1)], device=True, adopt=True)   # zero-copy: level-0 plane = the resident input
        T = init_of(step).copy(); ab = np.zeros((B, 2))
        r = ctx.trackBatch(slots, frame_ids(step), T, ab)
        return r, T

    def barrier():
        torch.cuda.synchronize(); ctx.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- leg 1: inputs resident in HBM
    sampler = ClockSampler(local_rank); sampler.start(); time.sleep(0.3)
    for s in range(W * R):
        step_dev(s)
    barrier(); l0 = ctx.launch_count()
    t0 = time.perf_counter(); kern_ms = 0.0; evals = 0; good = 0
    for s in range(W * R, (W + K) * R):
        r, T = step_dev(s)
        kern_ms += ctx.last_kernel_ms(); evals += int(r["evals"].sum()); good += int(r["good"].sum())
    barrier(); t_value = time.perf_counter() - t0
    launches = ctx.launch_count() - l0
    k_last = 1 + ((W + K) * R - 1) % (N_FRAMES - 1)
    errs = [np.abs(synth.se3_log7(synth.se3_mul7(T[b], synth.se3_inv7(gts[k_last])))) for b in range(min(B, 16))]
    pose_err_t = float(max(e[:3].max() for e in errs)); pose_err_r = float(max(e[3:].max() for e in errs))
    NB = K * R                                                            # timed batches per leg

    # ---------------------------------------------------------------- leg 2: end to end through host buffers (H2D + D2H every batch)
    def e2e_leg(upload, first, nb=None, nbatches=None):
        """nb: sequences of the batch this leg drives (default all B); nbatches: timed batches (default NB)"""
        nb = B if nb is None else nb; nbt = NB if nbatches is None else nbatches; sl = slots[:nb]
        upload(first)
        for s in range(first, first + W):                                # warm-up, same software pipeline
            upload(s + 1); T = init_of(s)[:nb].copy(); ab = np.zeros((nb, 2)); ctx.trackBatch(sl, frame_ids(s)[:nb], T, ab)
        barrier(); s1 = first + W
        t0 = time.perf_counter()
        for s in range(s1, s1 + nbt):
            if s + 1 < s1 + nbt:
                upload(s + 1)                                            # async H2D + pyramid of the next batch overlaps this batch's tracking
            T = init_of(s)[:nb].copy(); ab = np.zeros((nb, 2))
            ctx.trackBatch(sl, frame_ids(s)[:nb], T, ab)                 # D2H of poses/residuals inside
        barrier(); t = time.perf_counter() - t0
        # the first batch's upload happened before t0: charge it (one un-overlapped upload) so every batch's H2D is inside the timed region
        tu = time.perf_counter(); upload(s1 + nbt); ctx.sync(); t += time.perf_counter() - tu
        return t, s1 + nbt + 1

    # float leg (secondary: the makeImages(float*) seam): PCIe-bound at 4x the bytes, so it runs on 296 sequences and a quarter of the batches — 1.5 GB of pinned memory
    # per rank instead of 6 GB, same frames/s (the rate is set by the bytes per frame, not by the batch size)
    Bf = min(B, 296); NBf = max(4, NB // 4)
    host_in = torch.empty((N_FRAMES - 1, Bf, h, w), dtype=torch.float32, pin_memory=True)
    for k in range(N_FRAMES - 1):
        host_in[k].copy_(torch.from_numpy(frames_np[k]).expand(Bf, h, w))
    host_ptrs = [np.uint64(host_in[k].data_ptr()) + np.arange(Bf, dtype=np.uint64) * np.uint64(stride) for k in range(N_FRAMES - 1)]
    t_e2e, nxt = e2e_leg(lambda step: ctx.makeImagesBatch(frame_ids(step)[:Bf], host_ptrs[step % (N_FRAMES - 1)]), (W + K) * R, nb=Bf, nbatches=NBf)
    del host_in
    host_u8 = torch.empty((N_FRAMES - 1, B, ho, wo), dtype=torch.uint8, pin_memory=True)     # the RAW camera frames, one private copy per sequence
    for k in range(N_FRAMES - 1):
        host_u8[k].copy_(torch.from_numpy(seq.raw[1 + k]).expand(B, ho, wo))
    u8_ptrs = [np.uint64(host_u8[k].data_ptr()) + np.arange(B, dtype=np.uint64) * np.uint64(ho * wo) for k in range(N_FRAMES - 1)]

    def upload_u8(step):
        ctx.makeImagesBatch(frame_ids(step), u8_ptrs[step % (N_FRAMES - 1)], raw=True)
    t_e2e_u8, s3 = e2e_leg(upload_u8, nxt)
    # ---------------------------------------------------------------- leg 2c: the WHOLE FullSystem::trackNewCoarse per frame through host buffers: mono8 upload ->
    # motion hypotheses + trackNewestCoarse re-track loop -> reprojectMap -> structPoseEstimation (sdv_track_new_coarse_batch), pose D2H
    t_e2e_full = None; full_stats = None
    if not args.no_refine:
        KFM = (1 << 40) + 1; ctx.makeImages(KFM, seq.images[0])
        kf_c2w = np.concatenate([synth._quat_from_R(seq.R[0]), seq.t[0]])
        mp = np.zeros(len(pts), api.MAP_PT_DTYPE); mp["u"] = np.floor(pts[:, 0]); mp["v"] = np.floor(pts[:, 1]); mp["idepth"] = pts[:, 2]; mp["host"] = 0; mp["type"] = (np.arange(len(pts)) % 3 == 0)
        rp = api.Reprojector(ctx)
        for b in range(B):
            rp.setMap(b, [KFM], kf_c2w[None], None, mp)
        order = np.random.default_rng(3).permutation(rp.n_cells).astype(np.int32)
        KF_full = min(K * R, 3 * K); nfull = W + KF_full + 2
        io_all = np.zeros((N_INIT_POOL, B), api.TRACK_NEW_COARSE_DTYPE)
        for i in range(N_INIT_POOL):                                     # history chosen so that the constant-motion hypothesis equals the perturbed initial guess of the other legs:
            io = io_all[i]; io["slot"] = slots; io["poses_valid"] = 1    # slast = lastF (keyframe pose), sprelast = lastF * init  =>  try 0 = init
            io["lastF_c2w"] = kf_c2w; io["slast_c2w"] = kf_c2w; io["lastCoarseRMSE"] = 100.0
            for b in range(B):
                io["sprelast_c2w"][b] = synth.se3_mul7(kf_c2w, inits[i, b])
        upload_u8(s3)

        def full_step(i):
            io = io_all[(3 * ((s3 + i) // 3) + (s3 + i) % (N_FRAMES - 1)) % N_INIT_POOL].copy(); io["frame"] = frame_ids(s3 + i)
            api.trackNewCoarseBatchArray(ctx, io, cell_order=order, max_matches=400)
            return io
        for i in range(W):
            upload_u8(s3 + i + 1); full_step(i)
        barrier(); t0 = time.perf_counter()
        for i in range(W, W + KF_full):
            if i + 1 < W + KF_full:
                upload_u8(s3 + i + 1)
            io = full_step(i)
        barrier(); t_e2e_full = time.perf_counter() - t0
        tu = time.perf_counter(); upload_u8(s3 + W + KF_full); ctx.sync(); t_e2e_full += time.perf_counter() - tu
        k_last = 1 + (s3 + W + KF_full - 1) % (N_FRAMES - 1)
        gt_c2w = np.concatenate([synth._quat_from_R(seq.R[k_last]), seq.t[k_last]])
        full_stats = {"batches": KF_full, "tries_mean": float(io["tries"].mean()), "matches_mean": float(io["n_matches"].mean()), "refine_accepts_mean": float(io["refine_accepts"].mean()),
                      "median_translation_err_m": float(np.median(np.linalg.norm(io["camToWorld"][:, 4:] - gt_c2w[4:], axis=1)))}
    # ---------------------------------------------------------------- leg 2d: the batched runner (sdv_loam_b200/runner.py): world x B Monte-Carlo re-runs of the drive dealt to the
    # ranks by dist.shard_sequences, every re-run a CHAIN (frame k+1 starts from the constant-motion prediction of its own results), raw mono8 uploads, final reduction over ranks
    from sdv_loam_b200 import runner
    seeds = 100000 + np.arange(world * B)
    be = runner.GpuBackend(ctx, B, u8_ptrs, raw=True)
    runner.run_monte_carlo(be, seeds, N_FRAMES - 1, gts[1], rank, world, device="cuda")                  # warm-up pass
    mc = runner.run_monte_carlo(be, seeds, N_FRAMES - 1, gts[1], rank, world, device="cuda")
    mc_err = [np.abs(synth.se3_log7(synth.se3_mul7(mc["local_poses"][-1][b], synth.se3_inv7(gts[N_FRAMES - 1])))) for b in range(min(B, 64))]
    mc_runner = {"sequences": mc["sequences"], "frames_per_sequence": N_FRAMES - 1, "frames": mc["frames"], "seconds_max_over_ranks": mc["seconds"], "frames_per_s": mc["frames_per_s"],
                 "pose_digest": mc["pose_digest"], "tracked_ok_fraction_rank0": mc["local_ok_fraction"],
                 "final_pose_err_vs_gt_m_rad": [float(max(e[:3].max() for e in mc_err)), float(max(e[3:].max() for e in mc_err))],
                 "api": "runner.run_monte_carlo over runner.GpuBackend: sequence i -> rank i mod world (dist.shard_sequences), per frame sdv_frame_upload_batch_raw_u8 + "
                        "sdv_tracker_track_batch, next guess = constant motion from the re-run's own last two poses; the next frame's upload overlaps the tracking of the current one"}
    clocks = sampler.stop()
    del host_u8

    ctx.close()
    # back-end and refinement legs: their own context (the 8-frame window sequence of tests/, rendered for synth.KITTI_K)
    Bref = min(B, 592)
    ctx = api.Context(synth.KITTI_K, w, h, device=local_rank, n_tracker_slots=Bref, track_threads=args.track_threads, max_frames=Bref + 32 + 8 * WBA, max_kf_images=max(12, 7 * WBA))
    ba = ba_leg(ctx, api, synth, local_rank, WBA) if WBA > 0 else None
    refine = refine_leg(ctx, api, synth, Bref, cpu=not args.no_cpu_baseline) if not args.no_refine else None
    keyframe = None
    if not args.no_refine and rank == 0:
        try:
            keyframe = keyframe_leg(ctx, api, synth, min(Bref, 148), cpu=not args.no_cpu_baseline)
        except Exception as e:                                            # an auxiliary leg must not take the headline measurement down
            keyframe = {"error": repr(e)[:300]}
    h2d_local = B * ho * wo * NB / t_e2e_u8 / 1e9                          # this rank's raw-frame H2D rate over the e2e leg
    hv = torch.tensor([h2d_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(hv, op=dist.ReduceOp.MIN)
    tv = torch.tensor([t_value, t_e2e, kern_ms, t_e2e_u8, t_e2e_full or 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
    t_value, t_e2e, kern_ms_max, t_e2e_u8, t_e2e_full_max = [float(x) for x in tv.cpu()]
    ev = torch.tensor([float(evals), float(good)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ev, op=dist.ReduceOp.SUM)
    ctx.close()
    # ---------------------------------------------------------------- one sequence per GPU (BASELINE configs as written): every rank runs its own chain
    single = None
    if not args.no_extra_legs:
        single = single_sequence_leg(api, synth, seq, p4, rh, local_rank)
        sv = torch.tensor([single["frames_per_s_e2e"], 1.0 / single["track_kernel_ms"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(sv, op=dist.ReduceOp.SUM)
        single["one_seq_per_gpu_frames_per_s_e2e"] = float(sv[0].item()); single["gpus"] = world
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    job_bytes = api.track_job_bytes()
    achieved = evals * ALG_BYTES_PER_EVAL / (kern_ms * 1e-3) / 1e9      # this rank's kernel: algorithmic GB/s
    traffic = ncu_traffic()
    cfg = common_config(args, world, B)
    cfg.update({"l2_policy": "inputs larger than L2: %.1f GB of per-sequence pyramids+clouds per batch vs 126 MB L2" % (B * 9.3e-3),
                "timed_region_s": t_value, "tracked_ok_fraction": float(ev[1].item()) / (world * B * NB), "pose_err_vs_gt_m_rad": [pose_err_t, pose_err_r]})
    line = {
        "metric": METRIC, "value": world * B * NB / t_value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * t_value / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "e2e": {"value": world * B * NB / t_e2e_u8, "unit": "frames/s", "h2d_bytes_per_step": R * (B * ho * wo + B * job_bytes), "d2h_bytes_per_step": R * B * job_bytes,
                "h2d_GBps_per_gpu_min_over_ranks": float(hv.item()), "host_placement_rank0": numa,
                "api": "sdv_frame_upload_batch_raw_u8 (pinned host RAW 1241x376 mono8 = the sensor_msgs/Image wire format the reference ingests; Undistort::undistort crop-remap + u8->float "
                       "fused into the level-0/1 pyramid kernel, tables of calib/KITTI/00.txt via sdv_set_undistort) + sdv_tracker_track_batch; upload of batch k+1 overlapped with tracking of "
                       "batch k; the rectified images are bit-identical to the reference's undistort<unsigned char> output (tests/test_undistort.py), i.e. to what the other legs track"},
        "e2e_float32": {"value": world * Bf * NBf / t_e2e, "unit": "frames/s", "sequences_per_gpu": Bf, "batches": NBf, "h2d_bytes_per_batch": Bf * h * w * 4 + Bf * job_bytes, "d2h_bytes_per_batch": Bf * job_bytes,
                        "api": "sdv_frame_upload_batch(float*, pinned, host-rectified frames) = FrameHessian::makeImages(float*) signature + sdv_tracker_track_batch; PCIe-bound (4x the bytes of the wire format)"},
        "e2e_trackNewCoarse": (None if not full_stats else dict({"value": world * B * full_stats["batches"] / t_e2e_full_max, "unit": "frames/s", "h2d_bytes_per_batch": B * ho * wo + B * (job_bytes + 416),
                               "d2h_bytes_per_batch": B * (job_bytes + 416),
                               "api": "sdv_frame_upload_batch_raw_u8 + sdv_track_new_coarse_batch (the whole FullSystem::trackNewCoarse: hypotheses, trackNewestCoarse re-track loop, reprojectMap, "
                                      "structPoseEstimation) on host buffers; single-keyframe map of the bench sequence"}, **full_stats)),
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "track_cluster_kernel<128,4> (device-resident trackNewestCoarse: calcRes+calcGSSSE+LM; TMA bulk point staging + cp.async tap ring)", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": evals * ALG_BYTES_PER_EVAL / NB, "point_evals_per_launch": evals / NB,
                     "avg_launch_ms": kern_ms / NB, "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_source": (traffic or {}).get("source"),
                     "kernel_share_of_step": kern_ms * 1e-3 / t_value},
    }
    line["mc_runner"] = mc_runner
    if single is not None:
        line["single_sequence"] = single
        line["s_stress"] = stress_leg(api, local_rank)
    if ba is not None:
        ba["windows_per_s"] *= world   # every rank optimises its own windows (weak scaling, no collective)
        line["ba"] = ba
        fps_track = line["value"]; wps = ba["windows_per_s"]
        line["combined"] = {"kf_every": args.kf_every, "frames_per_s_track_plus_ba": 1.0 / (1.0 / fps_track + 1.0 / (args.kf_every * wps)),
                            "note": "tracking every frame + one FullSystem::optimize per kf_every frames, both legs measured separately on resident data"}
        if refine is not None:
            fr = refine["frames_per_s_device"] * world
            line["combined"]["frames_per_s_track_refine_ba"] = 1.0 / (1.0 / fps_track + 1.0 / fr + 1.0 / (args.kf_every * wps))
    if refine is not None:
        refine["frames_per_s_device"] *= world; refine["frames_per_s_wall"] *= world
        line["refine"] = refine
    if keyframe is not None:
        line["keyframe_rate"] = keyframe
        if "error" not in keyframe and ba is not None and refine is not None:
            # every stage of the per-frame path together, each measured separately on this GPU through the C-ABI: tracking + refinement every frame, the LiDAR front-end every
            # frame (one sweep per image, main.cpp:785), makeNewTraces + activation walk + FullSystem::optimize every kf_every-th frame
            kfe = args.kf_every; per_frame = 1.0 / line["value"] + 1.0 / refine["frames_per_s_device"] + 1.0 / (kfe * ba["windows_per_s"]) \
                + (1.0 / keyframe["lidar_front_end"]["sweeps_per_s"] + (1.0 / keyframe["make_new_traces"]["keyframes_per_s"] + 1.0 / keyframe["activate_select"]["sequences_per_s"]) / kfe) / world
            line["combined"]["frames_per_s_all_stages"] = 1.0 / per_frame
            line["combined"]["all_stages_note"] = "tracking + refinement + LiDAR front-end every frame; makeNewTraces + activation walk + optimize every kf_every-th frame; legs measured separately (keyframe-rate legs: wall time of the Python/ctypes call incl. host buffers)"
    if not args.no_cpu_baseline:
        arm = make_cpu_arm(seq, synth, p4, 1); arm.run(3)
        nf, tw = arm.run(10 ** 9, budget_s=12.0)
        if single is not None:
            single["cpu_1thread_frames_per_s"] = nf / tw
        if ba is not None:
            cms = ba_cpu_ms(synth); line["ba"]["cpu_ms_per_window_1core"] = cms
            line["combined"]["cpu_frames_per_s_track_plus_ba_1core"] = 1.0 / (tw / nf + cms * 1e-3 / args.kf_every)
            if refine is not None:
                line["combined"]["cpu_frames_per_s_track_refine_ba_1core"] = 1.0 / (tw / nf + refine["cpu_ms_per_frame_1core"] * 1e-3 + cms * 1e-3 / args.kf_every)
                if keyframe is not None and "cpu_ms_1core" in keyframe:
                    kc = keyframe["cpu_ms_1core"]
                    line["combined"]["cpu_frames_per_s_all_stages_1core"] = 1.0 / (tw / nf + refine["cpu_ms_per_frame_1core"] * 1e-3 + kc["lidar_front_end"] * 1e-3
                                                                                   + (cms + kc["make_new_traces"] + kc["activate_select"]) * 1e-3 / args.kf_every)
        line["cpu_baseline"] = {"value": nf / tw, "unit": "frames/s", "cores": 1, "kind": arm.kind,
                                "sample": "%d frames (undistort + makeImages + trackNewestCoarse, same inputs/inits distribution) in %.1f s on 1 host core; %s" % (nf, tw, cpu_note(arm))}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
