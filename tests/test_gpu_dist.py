"""GPU multi-rank tests: hypothesis-sharded registration over NCCL (one process per GPU).
Needs >= 2 GPUs (skipped otherwise); with 1 GPU the same host logic is covered by the
single-rank call and by tests/test_dist_cpu.py (gloo, world_size 2)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _list_angles(pkg):
    return pkg.sweep_angles(8.0)[1]


def test_hypothesis_sharded_single_rank_matches_batch(ctx, okss, pkg):
    """world = 1: the sharded driver over the single-object C ABI equals the batched pipeline and the oracle"""
    p = pkg.synth.modelnet_pair(33, n_full=600)
    r = pkg.dist.register_hypothesis_sharded(ctx, p["sim_s"], p["sim_t"], p["full_s"], p["full_t"],
                                             list_angles=_list_angles(pkg))
    b = ctx.register(p["sim_s"], p["sim_t"], p["full_s"], p["full_t"])
    o = okss.register(p["sim_s"], p["sim_t"], p["full_s"], p["full_t"], sum_order=okss.SUM_CANON256)
    assert r["winner"] == int(b["winner"]) == o["winner"]
    assert np.array_equal(r["T"], np.asarray(b["T"]).reshape(4, 4)) and np.array_equal(r["T"], o["T"])
    assert r["rmse"] == float(b["rmse"]) == o["rmse"]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pkg = entry.load_package()
    ctx = pkg.Context(rank)
    p = pkg.synth.modelnet_pair(34, n_full=800)

    def allreduce_min(v):                                   # the single NCCL all-reduce of this path
        t = torch.from_numpy(v).to("cuda:%d" % rank)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return t.cpu().numpy()
    r = pkg.dist.register_hypothesis_sharded(ctx, p["sim_s"], p["sim_t"], p["full_s"], p["full_t"], rank=rank,
                                             world=world, allreduce_min=allreduce_min, list_angles=pkg.sweep_angles(8.0)[1])
    q.put((rank, r["winner"], r["branch_multi"], r["rmse"], np.asarray(r["T"]).tolist()))
    ctx.close()
    dist.destroy_process_group()


def test_hypothesis_sharded_two_gpus_nccl(okss, pkg):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [mpc.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    out = [q.get(timeout=600) for _ in range(2)]
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    p = pkg.synth.modelnet_pair(34, n_full=800)
    ref = okss.register(p["sim_s"], p["sim_t"], p["full_s"], p["full_t"], sum_order=okss.SUM_CANON256)
    for rank, winner, multi, rmse, T in out:
        assert multi == ref["branch_multi"] and winner == ref["winner"]
        assert rmse == ref["rmse"] and np.array_equal(np.array(T, np.float32), ref["T"])
