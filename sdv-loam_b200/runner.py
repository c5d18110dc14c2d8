"""Batched-mode runner (north_star / SURVEY.md §8e): N independent sequences — here Monte-Carlo re-runs of one drive with perturbed initial guesses — sharded over
the ranks of a torch.distributed job (sequence i -> rank i mod world, dist.shard_sequences), every rank driving its shard frame by frame through ONE batched
trackNewestCoarse launch per frame, the initial guess of frame k+1 predicted from the results of frames k and k-1 (the constant-motion hypothesis FullSystem::trackNewCoarse
tries first, FullSystem.cpp:346-352).  No data-path collective: ranks meet only in the final reduction {frames, max seconds, pose digest} (dist.reduce_step_report).

The arithmetic sits behind a small backend protocol so that the host logic can be exercised without a GPU (tests/test_dist_cpu.py passes an oracle-backed backend;
this package never imports the oracle):
    backend.n                      number of LOCAL sequences (== len(shard))
    backend.track(step, T_pred)    upload/track frame `step` (1-based) of every local sequence from the predicted refToNew poses (n,7) -> (T_est (n,7), good (n,) bool)
    backend.sync()                 drain outstanding device work (timing fence)
"""
from __future__ import annotations
import time
import numpy as np
from . import dist as sdist
from . import synth


def perturbed_start(T_true: np.ndarray, seed: int, sigma_t: float = 0.04, sigma_r: float = 0.002) -> np.ndarray:
    """initial guess of one Monte-Carlo re-run: ground-truth relative pose perturbed like a constant-motion prediction error"""
    rng = np.random.default_rng(seed)
    return synth.se3_mul7(synth.se3_exp7(np.concatenate([rng.normal(0, sigma_t, 3), rng.normal(0, sigma_r, 3)])), T_true)


def _qmul_b(a, b):
    return np.stack([a[:, 0]*b[:, 0] - a[:, 1]*b[:, 1] - a[:, 2]*b[:, 2] - a[:, 3]*b[:, 3], a[:, 0]*b[:, 1] + a[:, 1]*b[:, 0] + a[:, 2]*b[:, 3] - a[:, 3]*b[:, 2],
                     a[:, 0]*b[:, 2] + a[:, 2]*b[:, 0] + a[:, 3]*b[:, 1] - a[:, 1]*b[:, 3], a[:, 0]*b[:, 3] + a[:, 3]*b[:, 0] + a[:, 1]*b[:, 2] - a[:, 2]*b[:, 1]], 1)


def _qrot_b(q, v):
    qv = q[:, 1:]; uv = 2.0 * np.cross(qv, v); return v + q[:, :1] * uv + np.cross(qv, uv)


def _mul_b(a, b):
    q = _qmul_b(a[:, :4], b[:, :4]); q = q / np.linalg.norm(q, axis=1, keepdims=True); return np.concatenate([q, a[:, 4:] + _qrot_b(a[:, :4], b[:, 4:])], 1)


def _inv_b(a):
    q = a[:, :4] * np.array([1.0, -1.0, -1.0, -1.0]); return np.concatenate([q, _qrot_b(q, -a[:, 4:])], 1)


def constant_motion(T_prev: np.ndarray, T_cur: np.ndarray) -> np.ndarray:
    """refToNew prediction for the next frame: the last inter-frame motion applied once more, (T_cur T_prev^-1) T_cur.  (7,) or batched (n,7): same formulas as
    synth.se3_mul7 / se3_inv7, vectorised over the local sequences (a Python loop over 1 184 re-runs cost 30x the device work of a chain step)."""
    a = np.asarray(T_prev, np.float64); b = np.asarray(T_cur, np.float64)
    if a.ndim == 1:
        return _mul_b(_mul_b(b[None], _inv_b(a[None])), b[None])[0]
    return _mul_b(_mul_b(b, _inv_b(a)), b)


def run_monte_carlo(backend, seeds, n_steps: int, T_first, rank: int = 0, world: int = 1, device=None):
    """Drive the local shard of `seeds` through frames 1..n_steps.  T_first: ground-truth refToNew of frame 1 (7,) — each re-run starts from its own perturbation of it.
    Returns a dict with the whole-job totals (identical on every rank) and this rank's poses."""
    mine = sdist.shard_sequences(len(seeds), rank, world)
    assert backend.n == len(mine), (backend.n, len(mine))
    n = len(mine); ID = np.array([1, 0, 0, 0, 0, 0, 0.0])
    T_prev = np.tile(ID, (n, 1))                                              # the keyframe itself (frame 0): identity
    T_pred = np.stack([perturbed_start(np.asarray(T_first, np.float64), int(seeds[i])) for i in mine]) if n else np.zeros((0, 7))
    poses = np.zeros((n_steps, n, 7)); ok = np.ones(n, bool)
    backend.sync(); t0 = time.perf_counter()
    for k in range(1, n_steps + 1):
        T_est, good = backend.track(k, T_pred)
        poses[k - 1] = T_est; ok &= np.asarray(good, bool)
        if k < n_steps:
            T_pred = constant_motion(T_prev, T_est) if n else T_pred
            T_prev = T_est
    backend.sync(); secs = time.perf_counter() - t0
    digest = float(np.abs(poses[-1]).sum()) if n and n_steps else 0.0          # order-independent checksum of the final poses of the shard
    frames, max_secs, digest_sum = sdist.reduce_step_report(n * n_steps, secs, digest, device=device)
    return {"sequences": len(seeds), "local_sequences": n, "frames": frames, "seconds": max_secs, "frames_per_s": frames / max_secs if max_secs > 0 else 0.0,
            "pose_digest": digest_sum, "local_ok_fraction": float(ok.mean()) if n else 1.0, "local_poses": poses, "shard": mine}


class GpuBackend:
    """The product path: one api.Context per rank, one tracker slot per local sequence (reference clouds set by the caller), frames uploaded per step from host buffers."""

    def __init__(self, ctx, n_local: int, frame_ptrs, raw: bool = False, u8: bool = False):
        """frame_ptrs[k-1]: uint64 array (n_local,) of host addresses of frame k of every local sequence (pinned for full-rate asynchronous copies)"""
        self.ctx, self.n, self.ptrs, self.raw, self.u8 = ctx, n_local, frame_ptrs, raw, u8
        self.slots = np.arange(n_local, dtype=np.int32)
        self.ids = [np.arange(n_local, dtype=np.uint64) * 2 + p for p in (0, 1)]   # two frame handles per sequence, alternating
        self.uploaded = 0                                                      # last step whose frames are on their way (the ingest is asynchronous)

    def _upload(self, step):
        self.ctx.makeImagesBatch(self.ids[step & 1], self.ptrs[step - 1], u8=self.u8, raw=self.raw); self.uploaded = step

    def track(self, step, T_pred):
        ids = self.ids[step & 1]
        if self.uploaded != step:
            self._upload(step)
        if step < len(self.ptrs):
            self._upload(step + 1)                                             # the NEXT frame's copy + pyramid overlap this frame's tracking: only the pose guess depends on its result
        T = np.ascontiguousarray(T_pred, np.float64).copy(); ab = np.zeros((self.n, 2))
        r = self.ctx.trackBatch(self.slots, ids, T, ab)
        return T, r["good"]

    def sync(self):
        self.ctx.sync(); self.uploaded = 0
