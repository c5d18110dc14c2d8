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
