"""GPU parity tests for clouds beyond the shared-memory path (> 2048 points): hierarchical
Morton-tile NN, canonical-order reductions, full-resolution ICP (KSS_ICP.hpp:133-183)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cloud(rng, n, scale=1.0):
    return (rng.normal(size=(n, 3)) * scale).astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("nq,nt", [(2049, 2049), (5000, 3000), (100, 40000), (30000, 2100), (70000, 70001)])
def test_large_nn_bit_exact(ctx, okss, nq, nt):
    rng = np.random.default_rng(nq + 3 * nt)
    q = _cloud(rng, nq); t = _cloud(rng, nt)
    t[nt // 2:nt // 2 + 20] = t[5:25]                      # duplicates: lowest index must win
    q[:20] = t[5:25]
    idx, d2 = ctx.nn_search(q, t)
    oi, od = okss.nn(q, t, okss.NN_KDTREE)
    assert np.array_equal(d2, od)
    assert np.array_equal(idx, oi)


def test_large_nn_surface_and_far(ctx, okss, pkg):
    """surface-like data (the real use) plus queries far from the target"""
    p = pkg.synth.scan_pair(1, 60000)
    q = np.concatenate([p["full_s"], p["full_s"][:500] * 40.0 + 7.0])
    idx, d2 = ctx.nn_search(q, p["full_t"])
    oi, od = okss.nn(q, p["full_t"], okss.NN_KDTREE)
    assert np.array_equal(d2, od) and np.array_equal(idx, oi)


def test_large_metrics(ctx, okss, pkg):
    """PCR_QM at full resolution: the sums run in the canonical parallel order on the GPU and
    serially in the oracle (registrationMeasure.hpp:66-88): stated tolerance 1e-12 relative"""
    p = pkg.synth.scan_pair(2, 30000)
    g = ctx.nn_metrics(p["full_s"], p["full_t"])
    o = okss.nn_metrics(p["full_s"], p["full_t"], okss.NN_KDTREE)
    assert np.allclose(g, o, rtol=1e-12, atol=0)


@pytest.mark.parametrize("n,index", [(3000, 0), (20000, 1), (70000, 2)])
def test_large_icp_bit_exact(ctx, okss, pkg, n, index):
    """full-resolution ICP: iteration count, final transform and fitness equal the oracle (CANON256)"""
    p = pkg.synth.scan_pair(index, n)
    o = okss.icp(p["full_s"], p["full_t"], sum_order=okss.SUM_CANON256, method=okss.NN_KDTREE)
    g = ctx.icp(p["full_s"], p["full_t"])
    assert g["iters"] == o["iters"] and g["converged"] == o["converged"]
    assert np.array_equal(g["T"], o["T"])
    assert g["fitness"] == o["fitness"]
    # and the recovered motion is the generated one (5 degrees, 0.02 shift) up to sampling noise
    R = g["T"][:3, :3].astype(np.float64)
    ang = np.arccos(np.clip((np.trace(R @ p["R"]) - 1) / 2, -1, 1))
    assert ang < 0.02


def test_large_icp_stepping_matches_run(ctx, okss, pkg):
    """begin / iterate / end (what bench.py times) gives the same answer as kss_icp"""
    p = pkg.synth.scan_pair(3, 10000)
    a = ctx.icp(p["full_s"], p["full_t"])
    ctx.icp_large_begin(p["full_s"], p["full_t"])
    ctx.icp_large_iterate(a["iters"] + 5)                  # extra iterations are no-ops after convergence
    b = ctx.icp_large_end()
    assert b["iters"] == a["iters"] and np.array_equal(a["T"], b["T"]) and a["fitness"] == b["fitness"]


def test_register_with_large_full_clouds(ctx, okss, pkg):
    """simplified clouds on the small path, full-resolution apply + PCR_QM on the large path"""
    rng = np.random.default_rng(3)
    p = pkg.synth.modelnet_pair(50, n_full=6000)
    sim_s = pkg.synth.simplify(p["full_s"], 1500, rng); sim_t = pkg.synth.simplify(p["full_t"], 1500, rng)
    r, pa = ctx.register(sim_s, sim_t, p["full_s"], p["full_t"], want_points=True)
    o = okss.register(sim_s, sim_t, p["full_s"], p["full_t"], sum_order=okss.SUM_CANON256, want_points=True)
    assert np.array_equal(pa, o["point_align"])
    assert np.array_equal(np.asarray(r["T"]).reshape(4, 4), o["T"])
    assert int(r["winner"]) == o["winner"]
    assert np.allclose([r["mse"], r["rmse"], r["mae"]], [o["mse"], o["rmse"], o["mae"]], rtol=1e-12, atol=0)


@pytest.mark.parametrize("case", ["identical", "line", "two_clusters", "plane_dup", "tiny_spread", "outside"])
def test_large_nn_degenerate_geometry(ctx, okss, case):
    """the block grid's cell size comes from a nearest-neighbour probe of the target: degenerate targets (zero extent,
    collinear, clusters far apart, duplicated sheets) and queries outside the target's box must stay exact"""
    rng = np.random.default_rng(abs(hash(case)) % 1000)
    n = 6000
    if case == "identical":
        t = np.tile(np.array([[0.25, -1.5, 3.0]]), (n, 1)); q = _cloud(rng, 3000) + t[0]
    elif case == "line":
        t = np.zeros((n, 3)); t[:, 0] = rng.uniform(-2, 2, n); q = _cloud(rng, 3000, 0.3); q[:, 0] *= 8
    elif case == "two_clusters":
        t = np.concatenate([_cloud(rng, n // 2, 0.01), _cloud(rng, n // 2, 0.01) + 50.0]); q = np.concatenate([_cloud(rng, 1500, 0.02), _cloud(rng, 1500, 20.0) + 25.0])
    elif case == "plane_dup":
        a = rng.uniform(-1, 1, (n // 2, 3)); a[:, 2] = 0.0
        t = np.concatenate([a, a]); q = a[:3000] + rng.normal(size=(3000, 3)) * 0.002
    elif case == "tiny_spread":
        t = _cloud(rng, n, 1e-6) + 1000.0; q = _cloud(rng, 3000, 2e-6) + 1000.0
    else:
        t = rng.uniform(0, 1, (n, 3)); q = rng.uniform(-3, 4, (3000, 3))
    t = t.astype(np.float32).astype(np.float64); q = q.astype(np.float32).astype(np.float64)
    idx, d2 = ctx.nn_search(q, t)
    oi, od = okss.nn(q, t, okss.NN_KDTREE)
    assert np.array_equal(d2, od)
    assert np.array_equal(idx, oi)


def test_large_nn_grid_ties(ctx, okss):
    """a lattice target with queries on cell centres: 8-way fp32 ties across fine cells and blocks of the grid"""
    g = np.stack(np.meshgrid(np.arange(20.0), np.arange(20.0), np.arange(20.0)), -1).reshape(-1, 3) / 16.0
    rng = np.random.default_rng(5)
    t = g[rng.permutation(len(g))]
    q = g[::2] + 1.0 / 32.0
    idx, d2 = ctx.nn_search(q, t)
    oi, od = okss.nn(q, t, okss.NN_BRUTE)
    assert np.array_equal(d2, od) and np.array_equal(idx, oi)


def test_register_batch_ragged_large_full_clouds(ctx, okss, pkg):
    """ragged batch whose FULL clouds take the large path (> 2048 points): final apply + PCR_QM must honour the per-pair
    counts (rows beyond a pair's count are padding)"""
    rng = np.random.default_rng(11)
    P, cap = 3, 5000
    cS = np.array([5000, 3100, 2500], np.int32); cT = np.array([4200, 5000, 2050], np.int32)
    fs = np.zeros((P, cap, 3)); ft = np.zeros((P, cap, 3)); ss = []; st = []
    for p in range(P):
        pr = pkg.synth.modelnet_pair(60 + p, n_full=cap)
        fs[p, :cS[p]] = pr["full_s"][:cS[p]]; ft[p, :cT[p]] = pr["full_t"][:cT[p]]
        fs[p, cS[p]:] = 1e3; ft[p, cT[p]:] = -1e3                       # padding that would wreck the metrics if it were used
        ss.append(pkg.synth.simplify(fs[p, :cS[p]], 1000, rng)); st.append(pkg.synth.simplify(ft[p, :cT[p]], 1000, rng))
    ss = np.stack(ss); st = np.stack(st)
    res, pa = ctx.register_batch(ss, st, fs, ft, counts=(None, None, cS, cT), want_points=True)
    for p in range(P):
        o = okss.register(ss[p], st[p], fs[p, :cS[p]], ft[p, :cT[p]], sum_order=okss.SUM_CANON256, want_points=True)
        assert np.array_equal(np.asarray(res[p]["T"]).reshape(4, 4), o["T"]) and int(res[p]["winner"]) == o["winner"]
        assert np.array_equal(pa[p, :cS[p]], o["point_align"])
        assert np.allclose([res[p]["mse"], res[p]["rmse"], res[p]["mae"]], [o["mse"], o["rmse"], o["mae"]], rtol=1e-12, atol=0)


@pytest.mark.parametrize("n,index,iters", [(20000, 4, 60), (70000, 5, 45)])
def test_large_icp_long_run_bit_exact(ctx, okss, pkg, n, index, iters):
    """far past PCL's convergence (rotation/translation and relative-MSE exits disabled, fixed iteration count): late
    iterations keep certified matches, re-certify from the target's neighbour lists or search only a few points -- every
    one of them has to leave the same correspondences as the oracle's kd-tree, or T and the fitness drift apart"""
    p = pkg.synth.scan_pair(index, n)
    kw = dict(max_iter=iters, trans_eps=-1.0, fit_eps=0.0)
    o = okss.icp(p["full_s"], p["full_t"], sum_order=okss.SUM_CANON256, method=okss.NN_KDTREE, **kw)
    g = ctx.icp(p["full_s"], p["full_t"], **kw)
    assert g["iters"] == o["iters"] and g["iters"] > 30
    assert np.array_equal(g["T"], o["T"])
    assert g["fitness"] == o["fitness"]


def test_large_icp_duplicate_targets_and_rim(ctx, okss, pkg):
    """targets that coincide (fp32-equal distances: the lowest index has to win in every iteration, certificates can
    never hold for them) and a source that reaches beyond the target's footprint (two targets about equally far)"""
    p = pkg.synth.scan_pair(6, 30000)
    t = p["full_t"].copy()
    t[15000:15400] = t[100:500]                            # 400 coincident pairs
    s = np.concatenate([p["full_s"], p["full_s"][:2000] * np.array([1.0, 1.0, 1.0]) + np.array([0.35, 0.0, 0.0])])
    kw = dict(max_iter=40, trans_eps=-1.0, fit_eps=0.0, max_corr_dist=0.2)
    o = okss.icp(s, t, sum_order=okss.SUM_CANON256, method=okss.NN_KDTREE, **kw)
    g = ctx.icp(s, t, **kw)
    assert g["iters"] == o["iters"]
    assert np.array_equal(g["T"], o["T"])
    assert g["fitness"] == o["fitness"]


@pytest.mark.parametrize("case", ["tiny_target", "identical", "line_target", "far_apart", "few_source_many_target", "two_sheets"])
def test_large_icp_degenerate_cases(ctx, okss, case):
    """the certificate machinery on inputs it was not tuned for: targets with fewer points than a neighbour list holds,
    coincident clouds (every distance 0), collinear targets, clouds that never meet (every match rejected by the
    max-distance test from some iteration on), a handful of source points, two sheets closer than the cell size"""
    rng = np.random.default_rng(abs(hash(case)) % 997)
    f = lambda a: a.astype(np.float32).astype(np.float64)
    kw = dict(max_iter=30, trans_eps=-1.0, fit_eps=0.0)
    if case == "tiny_target":
        t = f(rng.normal(size=(5, 3))); s = f(rng.normal(size=(3000, 3)) * 0.5)
    elif case == "identical":
        t = f(rng.uniform(-1, 1, (4000, 3))); s = t.copy()
    elif case == "line_target":
        t = np.zeros((3000, 3)); t[:, 0] = rng.uniform(-2, 2, 3000); t = f(t); s = f(rng.normal(size=(2500, 3)) * 0.2); s[:, 0] *= 8
    elif case == "far_apart":
        t = f(rng.uniform(0, 1, (3000, 3))); s = f(rng.uniform(0, 1, (3000, 3)) + 5.0); kw["max_corr_dist"] = 4.0
    elif case == "few_source_many_target":
        t = f(rng.uniform(-1, 1, (20000, 3))); s = f(rng.uniform(-0.8, 0.8, (7, 3)))
    else:
        a = rng.uniform(-1, 1, (3000, 3)); a[:, 2] = 0.0
        t = f(np.concatenate([a, a + np.array([0.0, 0.0, 1e-4])])); s = f(a[:2500] + rng.normal(size=(2500, 3)) * 0.003 + np.array([0.01, 0.0, 0.0]))
    o = okss.icp(s, t, sum_order=okss.SUM_CANON256, method=okss.NN_KDTREE, **kw)
    g = ctx.icp(s, t, **kw)
    assert g["iters"] == o["iters"] and g["converged"] == o["converged"]
    assert np.array_equal(g["T"], o["T"], equal_nan=True)
    assert g["fitness"] == o["fitness"] or (np.isnan(g["fitness"]) and np.isnan(o["fitness"]))


def test_large_icp_refine_lanes_switch(ctx, pkg):
    """lg_refine_kernel searches a flagged point with one thread or with four lanes (every fourth row of cells each, merged
    by shuffles), by the number of flagged points of the segment: both orders forced through KSS_LG_REFINE_G (read once per
    process, hence the subprocesses) leave iteration count, transform and fitness bit-identical to the default"""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, json; sys.path.insert(0, %r); import __graft_entry__ as e; pkg = e.load_package(); c = pkg.Context(0); "
            "p = pkg.synth.scan_pair(5, 40000); g = c.icp(p['full_s'], p['full_t'], max_iter=40, trans_eps=-1.0, fit_eps=0.0); "
            "print('RES ' + json.dumps({'iters': int(g['iters']), 'T': [float(x).hex() for x in g['T'].ravel()], 'fitness': float(g['fitness']).hex()}))") % root
    p = pkg.synth.scan_pair(5, 40000)
    g = ctx.icp(p["full_s"], p["full_t"], max_iter=40, trans_eps=-1.0, fit_eps=0.0)
    want = {"iters": int(g["iters"]), "T": [float(x).hex() for x in g["T"].ravel()], "fitness": float(g["fitness"]).hex()}
    for lanes in ("1", "4"):
        env = dict(os.environ, KSS_LG_REFINE_G=lanes)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        line = [l for l in out.stdout.splitlines() if l.startswith("RES ")][-1]
        assert json.loads(line[4:]) == want, "KSS_LG_REFINE_G=%s changes the result" % lanes
