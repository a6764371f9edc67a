"""BASELINE.json's full sizes, through size-independent properties (the oracle would need minutes to hours here):
configs[2] = the whole 2,468-pair ModelNet40-shape batch, configs[3] = one 1M-point full-resolution ICP."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rot_angle(R1, R2):
    """chord formula: well conditioned near zero, unlike arccos of the trace"""
    return 2.0 * float(np.arcsin(min(1.0, np.linalg.norm(R1 - R2) / (2.0 * np.sqrt(2.0)))))


def test_c3_full_batch_properties(ctx, okss, pkg):
    P = 2468
    b, prs = pkg.synth.modelnet_batch(P, n_full=2048)
    res = ctx.register_batch(None, None, b["full_s"], b["full_t"])
    # (1) every field is sane
    assert np.isfinite(res["final_fitness"]).all() and (res["final_fitness"] >= 0).all()
    assert np.isfinite(res["T"]).all() and (res["overflow"] == 0).all()
    assert ((res["n_minima"] >= 1) & (res["n_minima"] <= 729)).all()
    assert np.allclose(res["rmse"] ** 2, res["mse"], rtol=1e-12, atol=0)
    assert (res["mae"] <= res["rmse"] * (1 + 1e-12)).all()                    # Jensen
    # (2) the batch is only a batch: any pair registered alone gives the same bits
    rng = np.random.default_rng(0)
    for p in rng.choice(P, 6, replace=False):
        one = ctx.register_batch(None, None, b["full_s"][p:p + 1], b["full_t"][p:p + 1])[0]
        for name in res.dtype.names:
            assert np.array_equal(res[p][name], one[name]), (int(p), name)
    # (3) sampled parity with the oracle (AIVS + registration), exact
    for p in rng.choice(P, 3, replace=False):
        ss = okss.aivs_simplify(b["full_s"][p], 1024)[0]; st = okss.aivs_simplify(b["full_t"][p], 1024)[0]
        o = okss.register(ss, st, b["full_s"][p], b["full_t"][p])
        assert int(res[p]["winner"]) == o["winner"] and int(res[p]["n_minima"]) == o["n_minima"]
        assert np.array_equal(res[p]["T"].reshape(4, 4), o["T"]) and float(res[p]["final_fitness"]) == o["final_fitness"]
    # (4) it registers: the residual is at sampling-noise level for the bulk of the batch (the clouds are two
    #     independent samplings of the same CAD-like shape, unit-sphere normalised: spacing ~ 0.05)
    assert np.median(res["rmse"]) < 0.05
    assert (res["rmse"] < 0.1).mean() > 0.8


def test_c4_one_million_points_round_trip(ctx):
    """1M-point source = the target moved by a small rigid motion: the full-resolution ICP must undo it exactly
    enough that every point finds ITSELF again (indices are a permutation check, distances ~ fp32 rounding)"""
    n = 1000000
    rng = np.random.default_rng(7)
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    t = (u * (1.0 + 0.15 * np.sin(4 * u[:, :1]) * np.cos(3 * u[:, 1:2]))).astype(np.float32).astype(np.float64)
    a = 0.001                                  # motion of ~0.3 point spacings (spacing ~ 3.5e-3): most points match at once
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    s = (t @ R.T + np.array([2e-4, -1e-4, 1.5e-4])).astype(np.float32).astype(np.float64)
    ctx.icp_large_begin(s, t)
    ctx.icp_large_iterate(40)
    r = ctx.icp_large_end()
    assert r["converged"] == 1 and r["iters"] <= 40
    assert r["fitness"] < 1e-10, r
    assert _rot_angle(np.asarray(r["T"], np.float64).reshape(4, 4)[:3, :3], R.T) < 1e-5
    moved = ctx.apply_transform(np.asarray(r["T"], np.float32).reshape(4, 4), s)
    sel = rng.choice(n, 200000, replace=False)
    idx, d2 = ctx.nn_search(moved[sel], t)
    assert (idx == sel).mean() > 0.999                                       # exact duplicates aside
    assert float(np.median(d2)) < 1e-12 and float(d2.max()) < 1e-8
