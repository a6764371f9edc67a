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
