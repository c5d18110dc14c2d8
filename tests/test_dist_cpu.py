"""N>1 plumbing on CPU: world_size-2 gloo run of the batched-mode sharding + reductions (the data path has no collective)."""
import os
import socket
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "oracle"))
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import dist as sd
    import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sd.shard_sequences(5, rank, world)
    # each "sequence" = an independent Monte-Carlo pose chain evaluated with the oracle's SE3 (stands in for a tracked sequence)
    digest = 0.0
    for s in mine:
        rng = np.random.default_rng(s); T = np.array([1, 0, 0, 0, 0, 0, 0.0])
        for _ in range(10):
            T = orc.se3_mul(orc.se3_exp(rng.normal(0, 0.01, 6)), T)
        digest += float(np.abs(T).sum())
    dist.barrier()
    frames, secs, dsum = sd.reduce_step_report(10 * len(mine), 0.5 + rank, digest)
    out[rank] = (mine, frames, secs, dsum)
    dist.destroy_process_group()


def test_world2_gloo_sharding_and_reduction():
    world = 2; port = _free_port()
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert sorted(out[0][0] + out[1][0]) == [0, 1, 2, 3, 4] and not set(out[0][0]) & set(out[1][0])     # disjoint cover: seq i -> rank i mod N
    assert out[0][1] == out[1][1] == 50                                                                   # whole-job frames
    assert out[0][2] == out[1][2] == 1.5                                                                  # max over ranks, not one rank's clock
    assert abs(out[0][3] - out[1][3]) < 1e-12 and out[0][3] > 0


# ---------------------------------------------------------------------------------------------- the batched runner (sdv_loam_b200/runner.py) on 2 gloo ranks
class _OracleBackend:
    """CPU stand-in for runner.GpuBackend in the gloo test: one oracle tracker per local sequence (test infrastructure; the product backend is the CUDA context)."""

    def __init__(self, seq, n_local, levels):
        import orc
        from sdv_loam_b200 import synth
        self.orc = orc; self.n = n_local; self.L = levels; w, h = seq.wh
        pts = synth.select_points(seq.images[0], seq.clouds[0], 600); p4 = np.concatenate([pts, np.full((len(pts), 1), 1e-3, np.float32)], 1).astype(np.float32)
        f0 = orc.Frame(seq.images[0], levels); self.frames = [orc.Frame(im, levels) for im in seq.images]
        self.tr = []
        for _ in range(n_local):
            t = orc.CoarseTracker(w, h, levels, seq.K); t.setCoarseTrackingRef(f0, p4, np.zeros(len(p4), np.int32)); self.tr.append(t)

    def track(self, step, T_pred):
        T = np.zeros((self.n, 7)); good = np.zeros(self.n, bool)
        for j in range(self.n):
            r = self.tr[j].trackNewestCoarse(self.frames[step], T_pred[j], (0.0, 0.0), self.L - 1); T[j] = r["T"]; good[j] = r["good"]
        return T, good

    def sync(self):
        pass


def _runner_worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "oracle")); sys.path.insert(0, os.path.join(root, "tests"))
    import sdv_loam_b200  # noqa
    from sdv_loam_b200 import runner, synth, dist as sd
    from conftest import cached_sequence, SMALL_K, SMALL_WH
    import orc
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    seq = cached_sequence(3, 1000, SMALL_K, SMALL_WH); seeds = [11, 12, 13, 14, 15]
    T1 = orc.se3_from_rt(*synth.rel_pose(seq.R[0], seq.t[0], seq.R[1], seq.t[1]))
    be = _OracleBackend(seq, len(sd.shard_sequences(len(seeds), rank, world)), 4)
    r = runner.run_monte_carlo(be, seeds, 2, T1, rank, world)
    T2 = orc.se3_from_rt(*synth.rel_pose(seq.R[0], seq.t[0], seq.R[2], seq.t[2]))
    err = max(float(np.abs(orc.se3_log(orc.se3_mul(p, orc.se3_inv(T2)))).max()) for p in r["local_poses"][-1])
    out[(world, rank)] = (r["shard"], r["frames"], r["pose_digest"], r["local_ok_fraction"], err, r["seconds"])
    if world > 1:
        dist.destroy_process_group()


def test_batched_runner_shards_sequences_over_two_gloo_ranks():
    mgr = mp.Manager(); out = mgr.dict()
    _runner_worker(0, 1, 0, out)                                                                            # single process: the whole list
    mp.spawn(_runner_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    one, a, b = out[(1, 0)], out[(2, 0)], out[(2, 1)]
    assert one[0] == [0, 1, 2, 3, 4] and a[0] == [0, 2, 4] and b[0] == [1, 3]                              # sequence i -> rank i mod 2
    assert one[1] == a[1] == b[1] == 10                                                                     # whole-job frames on every rank
    assert abs(a[2] - b[2]) < 1e-12 and abs(a[2] - one[2]) < 1e-9                                           # same poses, however the sequences were dealt
    assert one[3] == a[3] == b[3] == 1.0 and max(one[4], a[4], b[4]) < 2e-2                                # every re-run tracked both frames onto the ground truth
    assert a[5] == b[5] > 0                                                                                  # job time = max over ranks
