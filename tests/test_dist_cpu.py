"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: pair sharding and the hypothesis-sharded
registration with its single all-reduce(MIN).  The compute backend here is the CPU oracle (tests may use
it); on the GPU box the same logic runs over kss_icp_b200.Context and NCCL (tests/test_gpu_dist.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """adapter with the method names of kss_icp_b200.Context"""
    def __init__(self, okss):
        self.o = okss

    def middle_align(self, s, t): return self.o.middle_align(s, t)
    def rotation_sweep(self, al, t, step): return self.o.sweep(al, t, step, self.o.SCORE_AVE, self.o.NN_KDTREE)
    def apply_similarity(self, p, a7, ang): return self.o.apply_similarity(p, a7, ang)
    def icp(self, s, t, max_iter=1000): return self.o.icp(s, t, max_iter=max_iter, sum_order=self.o.SUM_CANON256)
    def apply_transform(self, T, p): return self.o.apply_transform(T, p)
    def nn_metrics(self, a, t): return self.o.nn_metrics(a, t)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    from oracle import okss
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = entry.load_package()
    p = pkg.synth.modelnet_pair(33, n_full=300)

    def allreduce_min(v):
        t = torch.from_numpy(v.copy())
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return t.numpy()
    _, lst = okss.sweep_angles(8.0)
    r = pkg.dist.register_hypothesis_sharded(OracleBackend(okss), p["sim_s"], p["sim_t"], p["full_s"], p["full_t"],
                                             rank=rank, world=world, allreduce_min=allreduce_min, list_angles=lst)
    # pair sharding: every pair owned exactly once
    lo, hi = pkg.dist.shard_range(2468, world, rank)
    owned = torch.zeros(2468, dtype=torch.int32); owned[lo:hi] = 1
    dist.all_reduce(owned)
    q.put((rank, r["winner"], r["branch_multi"], r["rmse"], np.asarray(r["T"]).tolist(), int(owned.min()), int(owned.max())))
    dist.destroy_process_group()


def test_hypothesis_sharding_two_ranks_gloo(okss, pkg):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    out = [q.get(timeout=300) for _ in range(2)]
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    p = pkg.synth.modelnet_pair(33, n_full=300)
    ref = okss.register(p["sim_s"], p["sim_t"], p["full_s"], p["full_t"], sum_order=okss.SUM_CANON256)
    for rank, winner, multi, rmse, T, omin, omax in out:
        assert multi == ref["branch_multi"] and winner == ref["winner"]
        assert rmse == ref["rmse"] and np.array_equal(np.array(T, np.float32), ref["T"])
        assert omin == 1 and omax == 1


def test_select_winner_rule(pkg):
    sw = pkg.dist.select_winner
    assert sw([0.3, 0.1, 0.1, 0.2]) == 1                     # first strict minimum
    assert sw([np.inf, 12000.0, 9999.0]) == 0                # nothing below Q = 9999: default index 0
    assert sw([-1.0, 0.5]) == 1                              # ri >= 0 guard
    assert sw([]) == 0
    for total, world in ((2468, 8), (2468, 3), (5, 8), (0, 2)):
        rs = [pkg.dist.shard_range(total, world, r) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == total and all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
