"""GPU parity tests for the <= 2048-point path: every call goes through the C ABI
(libkss_icp_b200.so) and is compared with the CPU oracle on the same seeded inputs.

Bar: bit-exact for indices and for everything the oracle computes in the same
arithmetic (d2, sweep grid, MiddleAlign, ICP with the CANON256 reduction order,
PCR_QM); fp32-level tolerance only between the oracle's two reduction orders."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cloud(rng, n, scale=1.0):
    return (rng.normal(size=(n, 3)) * scale).astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("nq,nt", [(1, 1), (5, 31), (32, 32), (33, 65), (1000, 777), (2048, 2048), (300, 2047)])
def test_nn_search_bit_exact(ctx, okss, nq, nt):
    rng = np.random.default_rng(nq * 7919 + nt)
    q = _cloud(rng, nq); t = _cloud(rng, nt)
    idx, d2 = ctx.nn_search(q, t)
    oi, od = okss.nn(q, t, okss.NN_BRUTE)
    assert np.array_equal(d2, od)
    assert np.array_equal(idx, oi)


def test_nn_search_ties_lowest_index(ctx, okss):
    """duplicated target points and lattice data produce fp32-equal distances: lowest index wins"""
    rng = np.random.default_rng(5)
    t = _cloud(rng, 500)
    t[400:450] = t[100:150]                       # exact duplicates
    g = np.stack(np.meshgrid(np.arange(8.0), np.arange(8.0), np.arange(8.0)), -1).reshape(-1, 3)
    t = np.concatenate([t, g])                    # lattice: equidistant neighbours
    q = np.concatenate([t[100:150], g + 0.5, _cloud(rng, 100)])
    idx, d2 = ctx.nn_search(q, t)
    oi, od = okss.nn(q, t, okss.NN_BRUTE)
    assert np.array_equal(d2, od) and np.array_equal(idx, oi)
    assert (idx[:50] == np.arange(100, 150)).all()


def test_nn_search_far_queries(ctx, okss):
    """queries far outside the target box (culling bounds must stay conservative)"""
    rng = np.random.default_rng(11)
    t = _cloud(rng, 1500, 0.1)
    q = _cloud(rng, 700, 50.0)
    idx, d2 = ctx.nn_search(q, t)
    oi, od = okss.nn(q, t, okss.NN_BRUTE)
    assert np.array_equal(d2, od) and np.array_equal(idx, oi)


@pytest.mark.parametrize("ns,nt", [(3, 3), (100, 257), (1024, 1024), (2000, 1999)])
def test_middle_align_bit_exact(ctx, okss, ns, nt):
    rng = np.random.default_rng(ns + nt)
    s = _cloud(rng, ns, 2.0) + 3.0; t = _cloud(rng, nt, 0.7) - 1.0
    o7, oal = okss.middle_align(s, t)
    g7, gal = ctx.middle_align(s, t)
    assert np.array_equal(g7, o7)
    assert np.array_equal(gal, oal)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_sweep_bit_exact(ctx, okss, pkg, mode):
    p = pkg.synth.modelnet_pair(3, n_full=600)
    _, al = okss.middle_align(p["sim_s"], p["sim_t"])
    o = okss.sweep(al, p["sim_t"], 8.0, mode, okss.NN_KDTREE)
    g = ctx.rotation_sweep(al, p["sim_t"], 8.0, mode)
    assert g["G"] == 9 and o["G"] == 9
    assert np.array_equal(g["value"], o["value"])
    assert np.array_equal(g["best_index"], o["best_index"])
    assert np.array_equal(g["best_angle"], o["best_angle"])
    assert np.array_equal(g["minima"], o["minima"])


def test_sweep_voxel_score_mode(ctx, okss, pkg):
    """KSS_SCORE_VOXEL (optional, NOT a mode of the released sources: a reading of the authors' closed CUDA build): the
    occupancy-grid score of every hypothesis, its arg-min and its local minima equal the oracle's bit for bit"""
    for seed, n in ((3, 600), (9, 2000)):
        p = pkg.synth.modelnet_pair(seed, n_full=n)
        _, al = okss.middle_align(p["sim_s"], p["sim_t"])
        o = okss.sweep(al, p["sim_t"], 8.0, okss.SCORE_VOXEL, okss.NN_KDTREE)
        g = ctx.rotation_sweep(al, p["sim_t"], 8.0, pkg.SCORE_VOXEL)
        assert np.array_equal(g["value"], o["value"])
        assert np.array_equal(g["best_index"], o["best_index"]) and np.array_equal(g["minima"], o["minima"])
        assert 0.0 <= g["value"].min() and g["value"].max() <= 1.0
    t = np.zeros((300, 3)); s_ = np.random.default_rng(1).normal(size=(300, 3))       # zero-extent target: every point misses
    g = ctx.rotation_sweep(s_, t, 8.0, pkg.SCORE_VOXEL)
    o = okss.sweep(s_, t, 8.0, okss.SCORE_VOXEL, okss.NN_KDTREE)
    assert np.array_equal(g["value"], o["value"]) and np.all(g["value"] == 1.0)


def test_sweep_other_steps(ctx, okss, pkg):
    p = pkg.synth.modelnet_pair(4, n_full=256)
    _, al = okss.middle_align(p["sim_s"], p["sim_t"])
    for step, G in ((6.0, 6), (12.0, 12)):
        o = okss.sweep(al, p["sim_t"], step, 0, okss.NN_KDTREE)
        g = ctx.rotation_sweep(al, p["sim_t"], step, 0)
        assert g["G"] == G
        assert np.array_equal(g["value"], o["value"])
        assert np.array_equal(g["minima"], o["minima"])


def test_similarity_and_transform_bit_exact(ctx, okss):
    rng = np.random.default_rng(2)
    pts = _cloud(rng, 5000)
    a7 = np.array([0.1, -0.2, 0.3, 1.5, -0.7, 0.2, 1.37])
    ang = np.array([0.7875, 3.15, 5.5125])
    assert np.array_equal(ctx.apply_similarity(pts, a7, ang), okss.apply_similarity(pts, a7, ang))
    T = np.eye(4, dtype=np.float32); T[:3, :3] = rng.normal(size=(3, 3)); T[:3, 3] = [0.1, 0.2, 0.3]
    assert np.array_equal(ctx.apply_transform(T, pts), okss.apply_transform(T, pts))


@pytest.mark.parametrize("n", [64, 700, 2048])
def test_nn_metrics_bit_exact(ctx, okss, n):
    rng = np.random.default_rng(n)
    a = _cloud(rng, n); t = _cloud(rng, n + 13 if n < 2048 else n)
    assert np.array_equal(ctx.nn_metrics(a, t), okss.nn_metrics(a, t, okss.NN_BRUTE))


def _icp_inputs(pkg, okss, index, n_full, hyp=None):
    p = pkg.synth.modelnet_pair(index, n_full=n_full)
    a7, al = okss.middle_align(p["sim_s"], p["sim_t"])
    sw = okss.sweep(al, p["sim_t"], 8.0, 0, okss.NN_KDTREE)
    ang = sw["best_angle"] if hyp is None else okss.sweep_angles(8.0)[1][sw["minima"][hyp % len(sw["minima"])]]
    src = okss.apply_similarity(p["sim_s"], a7, ang)
    return src, p["sim_t"]


@pytest.mark.parametrize("index,n_full,hyp", [(0, 400, None), (1, 1024, None), (2, 2048, 0), (5, 2048, 3), (6, 4000, 1)])
def test_icp_free_running_bit_exact(ctx, okss, pkg, index, n_full, hyp):
    """whole ICP runs: per-iteration correspondences, T_k, MSE, iteration count, final T and
    fitness are identical to the oracle in CANON256 order (free running, not teacher forced)"""
    src, tgt = _icp_inputs(pkg, okss, index, n_full, hyp)
    cap = 64
    o = okss.icp(src, tgt, sum_order=okss.SUM_CANON256, method=okss.NN_KDTREE, trace_iters=cap)
    g = ctx.icp(src, tgt, trace_iters=cap)
    assert g["iters"] == o["iters"] and g["converged"] == o["converged"]
    k = min(cap, o["iters"])
    assert np.array_equal(g["trace"]["src_k"][:k], o["trace"]["src_k"][:k])
    assert np.array_equal(g["trace"]["corr_idx"][:k], o["trace"]["corr_idx"][:k])
    assert np.array_equal(g["trace"]["T_k"][:k], o["trace"]["T_k"][:k])
    assert np.array_equal(g["trace"]["mse"][:k], o["trace"]["mse"][:k])
    assert np.array_equal(g["T"], o["T"])
    assert g["fitness"] == o["fitness"]
    # reduction order of the reference's Eigen sums is unknown: the serial order agrees to fp32 rounding
    s = okss.icp(src, tgt, sum_order=okss.SUM_SERIAL, method=okss.NN_KDTREE)
    R1, R2 = s["T"][:3, :3].astype(np.float64), g["T"][:3, :3].astype(np.float64)
    ang = 2.0 * np.arcsin(min(1.0, np.linalg.norm(R1 - R2) / (2.0 * np.sqrt(2.0))))   # chord: well conditioned near zero
    dt = np.abs(s["T"][:3, 3] - g["T"][:3, 3]).max()
    if s["iters"] == o["iters"]:
        assert ang < 1e-4                               # north_star: rotation error < 1e-4 rad
        assert dt < 1e-4
        assert abs(s["fitness"] - g["fitness"]) < 1e-4 * s["fitness"] + 1e-9   # fp32 tolerance, stated
    else:
        # the two summation orders crossed PCL's relative-MSE stop (1e-3) on different iterations: the results then
        # differ by one ICP update at the stopping threshold -- bounded by that threshold, not by fp32 rounding
        assert abs(s["iters"] - o["iters"]) <= 2
        assert ang < 2e-3 and dt < 2e-3
        assert abs(s["fitness"] - g["fitness"]) < 1e-2 * s["fitness"] + 1e-9


def test_icp_rejection_and_too_few(ctx, okss):
    """max correspondence distance drops pairs (A.3); < 3 correspondences ends the run unconverged"""
    rng = np.random.default_rng(9)
    t = _cloud(rng, 500, 0.2)
    s = t[:300] + 0.01
    s[:40] += 5.0                                     # far outliers: d2 > 1 -> rejected
    o = okss.icp(s, t, sum_order=okss.SUM_CANON256, trace_iters=8)
    g = ctx.icp(s, t, trace_iters=8)
    assert (o["trace"]["corr_idx"][0][:40] == -1).all()
    assert np.array_equal(g["trace"]["corr_idx"][:min(8, o["iters"])], o["trace"]["corr_idx"][:min(8, o["iters"])])
    assert g["iters"] == o["iters"] and np.array_equal(g["T"], o["T"]) and g["fitness"] == o["fitness"]
    far = t[:10] + 100.0
    o = okss.icp(far, t, sum_order=okss.SUM_CANON256)
    g = ctx.icp(far, t)
    assert o["converged"] == 0 and o["iters"] == 0
    assert g["converged"] == 0 and g["iters"] == 0 and g["fitness"] == o["fitness"]


def test_icp_max_iterations(ctx, okss, pkg):
    src, tgt = _icp_inputs(pkg, okss, 8, 800)
    o = okss.icp(src, tgt, max_iter=3, sum_order=okss.SUM_CANON256)
    g = ctx.icp(src, tgt, max_iter=3)
    assert g["iters"] == o["iters"] <= 3 and g["converged"] == o["converged"]
    assert np.array_equal(g["T"], o["T"]) and g["fitness"] == o["fitness"]


def _check_pair(r, o, G=9):
    acc = None
    assert np.array_equal(np.asarray(r["align"])[:7], o["align"])
    bi = o["best_index"]
    assert int(r["best_h"]) == (bi[0] * G + bi[1]) * G + bi[2]
    assert int(r["n_minima"]) == o["n_minima"]
    assert int(r["branch_multi"]) == o["branch_multi"]
    assert int(r["winner"]) == o["winner"]
    assert float(r["judge_fitness"]) == o["judge_fitness"]
    assert float(r["final_fitness"]) == o["final_fitness"]
    assert int(r["judge_iters"]) == o["judge_iters"] and int(r["final_iters"]) == o["final_iters"]
    assert int(r["total_icp_iters"]) == o["total_icp_iters"] and int(r["n_icp_runs"]) == o["n_icp_runs"]
    assert np.array_equal(np.asarray(r["T"]).reshape(4, 4), o["T"])
    assert float(r["mse"]) == o["mse"] and float(r["rmse"]) == o["rmse"] and float(r["mae"]) == o["mae"]
    return acc


def test_register_batch_matches_oracle(ctx, okss, pkg):
    """KSSICP_Registration + PCR_QM for a small batch: every reported quantity equals the oracle's"""
    P = 6
    b, _ = pkg.synth.modelnet_batch(P, n_full=512, first=20)
    res, pa = ctx.register_batch(b["sim_s"], b["sim_t"], b["full_s"], b["full_t"], want_points=True)
    for p in range(P):
        o = okss.register(b["sim_s"][p], b["sim_t"][p], b["full_s"][p], b["full_t"][p],
                          sum_order=okss.SUM_CANON256, method=okss.NN_KDTREE, want_points=True)
        _check_pair(res[p], o)
        assert np.array_equal(pa[p], o["point_align"])


def test_register_ragged_counts(ctx, okss, pkg):
    """per-pair point counts below the batch capacity (ragged inputs)"""
    P = 3
    b, _ = pkg.synth.modelnet_batch(P, n_full=400, first=40)
    cs = np.array([200, 150, 97], np.int32); ct = np.array([200, 180, 33], np.int32)
    cS = np.array([400, 333, 100], np.int32); cT = np.array([400, 64, 399], np.int32)
    res = ctx.register_batch(b["sim_s"], b["sim_t"], b["full_s"], b["full_t"], counts=(cs, ct, cS, cT))
    for p in range(P):
        o = okss.register(b["sim_s"][p][:cs[p]], b["sim_t"][p][:ct[p]], b["full_s"][p][:cS[p]], b["full_t"][p][:cT[p]],
                          sum_order=okss.SUM_CANON256, method=okss.NN_KDTREE)
        _check_pair(res[p], o)


@pytest.mark.parametrize("slots", [1, 2, 5])
def test_register_hypothesis_slot_overflow(ctx, okss, pkg, slots):
    """more local minima than hypothesis slots: a slot's CTA then runs several hypotheses in turn inside the same
    launch (KSS_ICP.hpp:102-118 keeps the first strict minimum, which is the lexicographic minimum (fitness, l)) --
    same answer through the host-buffer entry AND the device entry, which the bench's headline `value` is timed on"""
    import torch
    P = 8
    b, _ = pkg.synth.modelnet_batch(P, n_full=512, first=20)
    ora = [okss.register(b["sim_s"][p], b["sim_t"][p], b["full_s"][p], b["full_t"][p], sum_order=okss.SUM_CANON256,
                         method=okss.NN_KDTREE) for p in range(P)]
    assert sum(1 for o in ora if o["branch_multi"] and o["n_minima"] > slots) >= 2      # the case is really exercised
    ctx.set_hyp_slots(slots)
    try:
        res = ctx.register_batch(b["sim_s"], b["sim_t"], b["full_s"], b["full_t"])
        dev = {k: torch.from_numpy(np.ascontiguousarray(b[k])).cuda() for k in ("sim_s", "sim_t", "full_s", "full_t")}
        d_res = torch.zeros(P * pkg.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        caps = (b["sim_s"].shape[1], b["sim_t"].shape[1], b["full_s"].shape[1], b["full_t"].shape[1])
        ctx.register_batch_device(P, caps, tuple(dev[k].data_ptr() for k in ("sim_s", "sim_t", "full_s", "full_t")),
                                  d_res.data_ptr())
        ctx.synchronize()
        res_d = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=pkg.RESULT_DTYPE)
    finally:
        ctx.set_hyp_slots(32)
    for p in range(P):
        _check_pair(res[p], ora[p])
        _check_pair(res_d[p], ora[p])
        assert int(res[p]["overflow"]) == int(ora[p]["branch_multi"] and ora[p]["n_minima"] > slots)
        assert int(res[p]["final_converged"]) == 1


def test_bad_arguments(ctx, pkg):
    with pytest.raises(pkg.KssError):
        ctx.rotation_sweep(np.zeros((3000, 3)), np.zeros((10, 3)))
    with pytest.raises(pkg.KssError):
        ctx.middle_align(np.zeros((0, 3)), np.zeros((10, 3)))


@pytest.mark.parametrize("model", ["Bunny", "Horse", "Dog"])
def test_golden_fixtures_from_reference_data(ctx, pkg, model):
    """the reference's own shipped pairs (data/registration/*.wlop|.gird, decimated; see
    tests/golden/make_golden.py) against the committed golden vectors: no oracle at run time"""
    import os
    gd = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gd, "golden_oracle.npz")); fx = np.load(os.path.join(gd, "fixtures_pairs.npz"))
    s = fx[model + "_src"].astype(np.float64); t = fx[model + "_tgt"].astype(np.float64)
    a7, al = ctx.middle_align(s, t)
    assert np.array_equal(a7, g[model + "_align7"])
    sw = ctx.rotation_sweep(al, t, 8.0)
    assert np.array_equal(sw["value"], g[model + "_value"])
    assert np.array_equal(sw["minima"], g[model + "_minima"])
    r = ctx.register(s, t, s, t)
    assert int(r["winner"]) == int(g[model + "_canon_winner"])
    assert np.array_equal(np.asarray(r["T"]).reshape(4, 4), g[model + "_canon_T"])
    assert float(r["final_fitness"]) == float(g[model + "_canon_final_fitness"])
    assert float(r["rmse"]) == float(g[model + "_canon_rmse"]) and float(r["mae"]) == float(g[model + "_canon_mae"])
    assert int(r["total_icp_iters"]) == int(g[model + "_canon_total_icp_iters"])


def test_icp_correspondence_ties_through_the_candidate_grid(ctx, okss):
    """duplicated target points and a lattice target give fp32-equal distances inside the ICP loop: the candidate
    grid's lists are ascending in original index, so the first (lowest index) of equal distances must win"""
    rng = np.random.default_rng(77)
    g = np.stack(np.meshgrid(np.arange(9.0), np.arange(9.0), np.arange(9.0)), -1).reshape(-1, 3) / 8.0 - 0.5
    t = np.concatenate([g, g[100:300], rng.uniform(-0.5, 0.5, (100, 3))]).astype(np.float32).astype(np.float64)
    perm = rng.permutation(len(t)); t = t[perm]
    s = np.concatenate([g + 1.0 / 16.0, g[::3]]).astype(np.float32).astype(np.float64)      # cell centres: 8-way ties
    o = okss.icp(s, t, sum_order=okss.SUM_CANON256, method=okss.NN_BRUTE, trace_iters=6)
    r = ctx.icp(s, t, trace_iters=6)
    k = min(o["iters"], 6)
    assert r["iters"] == o["iters"] and k >= 1
    assert np.array_equal(r["trace"]["corr_idx"][:k], o["trace"]["corr_idx"][:k])
    assert np.array_equal(r["T"], o["T"]) and r["fitness"] == o["fitness"]


@pytest.mark.parametrize("env", [{"KSS_NO_CG": "1"}, {"KSS_CG_NO_REFINE": "1"}, {"KSS_CG_COOP_LEVELS": "0"},
                                 {"KSS_CG_COOP_LEVELS": "14"}, {"KSS_LANES": "1"}, {"KSS_CHUNKS": "5"}])
def test_engine_switches_do_not_change_results(ctx, pkg, monkeypatch, env):
    """three NN engines, one result: the tile search, the candidate grid without / with its sparse levels, either
    build kernel for the dense levels, and any split of the batch over lanes and chunks give identical registrations"""
    P = 10
    b, _ = pkg.synth.modelnet_batch(P, n_full=700, first=200)
    ref = ctx.register_batch(None, None, b["full_s"], b["full_t"])
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got = ctx.register_batch(None, None, b["full_s"], b["full_t"])
    for name in ref.dtype.names:
        assert np.array_equal(ref[name], got[name]), name
