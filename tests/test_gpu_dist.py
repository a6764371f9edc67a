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


def _check_same(a, b):
    for name in a.dtype.names:
        assert np.array_equal(a[name], b[name]), name


def test_library_hyp_sharded_world1_equals_batch(ctx, pkg):
    """kss_register_batch_hyp_sharded without a communicator is the unsharded pipeline"""
    b, _ = pkg.synth.modelnet_batch(5, n_full=512, first=70)
    a = ctx.register_batch(b["sim_s"], b["sim_t"], b["full_s"], b["full_t"])
    s = ctx.register_batch_hyp_sharded(b["sim_s"], b["sim_t"], b["full_s"], b["full_t"])
    _check_same(a, s)


@pytest.mark.parametrize("world", [2, 4])
def test_library_hyp_sharded_nccl_in_library(okss, pkg, world):
    """the product path of SURVEY.md 8e rows 2-3: sweep slabs + ncclAllGather, ICP runs round-robin + ONE
    ncclAllReduce(min) inside the library (kss_register_batch_hyp_sharded), one context per GPU in one process
    (kss_ctx_nccl_init_all), every rank's results identical to the oracle and to the unsharded pipeline -- raw clouds
    (AIVS inside), ragged hypothesis counts, more local minima than slots on some pairs"""
    import threading
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    P = 6
    b, _ = pkg.synth.modelnet_batch(P, n_full=512, first=20)
    ctxs = [pkg.Context(r) for r in range(world)]
    try:
        pkg.nccl_init_all(ctxs)
        for c in ctxs:
            c.set_hyp_slots(3)                                 # some pairs have > 3 * world local minima: slots stride
        out = [None] * world
        err = [None] * world

        def run(r):
            try:
                out[r] = ctxs[r].register_batch_hyp_sharded(b["sim_s"], b["sim_t"], b["full_s"], b["full_t"], want_points=True)
            except Exception as e:                             # noqa: BLE001
                err[r] = e
        th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join(timeout=600)
        assert all(e is None for e in err), err
        ref = ctxs[0].register_batch(b["sim_s"], b["sim_t"], b["full_s"], b["full_t"], want_points=True)
        for r in range(world):
            res, pa = out[r]
            assert np.array_equal(pa, ref[1])
            for name in res.dtype.names:
                if name == "overflow":
                    continue
                assert np.array_equal(res[name], ref[0][name]), (r, name)
        for p in range(P):
            o = okss.register(b["sim_s"][p], b["sim_t"][p], b["full_s"][p], b["full_t"][p], sum_order=okss.SUM_CANON256)
            assert int(out[0][0][p]["winner"]) == o["winner"] and int(out[0][0][p]["total_icp_iters"]) == o["total_icp_iters"]
            assert np.array_equal(out[0][0][p]["T"].reshape(4, 4), o["T"]) and float(out[0][0][p]["rmse"]) == o["rmse"]
    finally:
        for c in ctxs:
            c.close()


def test_register_batch_multi_contexts(ctx, pkg):
    """kss_register_batch_multi: pairs in contiguous blocks over several contexts, host threads inside the library
    (two contexts on one GPU when there is only one; one per GPU otherwise) -- same bits as a single context"""
    import torch
    n = max(2, min(4, torch.cuda.device_count()))
    ctxs = [pkg.Context(i % torch.cuda.device_count()) for i in range(n)]
    try:
        b, _ = pkg.synth.modelnet_batch(11, n_full=512, first=90)
        ref = ctx.register_batch(None, None, b["full_s"], b["full_t"], want_points=True)
        res, pa = pkg.register_batch_multi(ctxs, None, None, b["full_s"], b["full_t"], want_points=True)
        _check_same(res, ref[0])
        assert np.array_equal(pa, ref[1])
    finally:
        for c in ctxs:
            c.close()
