"""CPU tests of the oracle (oracle/kss_oracle.cpp): golden vectors made from the reference's own
shipped pairs, the known-answer rotations of data/registration/transfer.txt, and the reference
quirks of SURVEY.md Appendix B.  No GPU needed."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "golden_oracle.npz")), np.load(os.path.join(GOLD, "fixtures_pairs.npz"))


def _rot(axis, a):
    c, s = np.cos(a), np.sin(a)
    return [np.array([[1, 0, 0], [0, c, -s], [0, s, c]]), np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]),
            np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])][axis]


@pytest.mark.parametrize("model", ["Bunny", "Horse", "Dog"])
def test_oracle_matches_golden(okss, gold, model):
    g, fx = gold
    s = fx[model + "_src"].astype(np.float64); t = fx[model + "_tgt"].astype(np.float64)
    a7, al = okss.middle_align(s, t)
    assert np.array_equal(a7, g[model + "_align7"])
    sw = okss.sweep(al, t, 8.0, okss.SCORE_AVE, okss.NN_KDTREE)
    assert np.array_equal(sw["value"], g[model + "_value"])
    assert np.array_equal(sw["minima"], g[model + "_minima"])
    assert np.array_equal(sw["best_index"], g[model + "_best_index"])


@pytest.mark.parametrize("model", ["Bunny", "Dog"])
def test_oracle_registration_golden_and_known_rotation(okss, gold, model):
    """end to end on the reference's data: the recovered rotation is the one transfer.txt documents
    (tolerance = resampling noise between the .wlop and .gird samplings, not fp32)"""
    g, fx = gold
    s = fx[model + "_src"].astype(np.float64); t = fx[model + "_tgt"].astype(np.float64)
    for tag, order in (("canon", okss.SUM_CANON256), ("serial", okss.SUM_SERIAL)):
        r = okss.register(s, t, s, t, sum_order=order, method=okss.NN_KDTREE)
        assert r["winner"] == int(g[model + "_%s_winner" % tag])
        assert np.array_equal(r["T"], g[model + "_%s_T" % tag])
        assert r["rmse"] == float(g[model + "_%s_rmse" % tag])
        ua = r["used_angle"]
        R = r["T"][:3, :3].astype(np.float64) @ _rot(2, ua[2]) @ _rot(1, ua[1]) @ _rot(0, ua[0])
        ax, ang = g[model + "_axis_angle"]
        err = np.arccos(np.clip((np.trace(R.T @ _rot(int(ax), ang)) - 1) / 2, -1, 1))
        assert err < 0.06


def test_serial_and_canonical_orders_agree_to_fp32(gold):
    """Eigen's reduction order is unknown: both oracle orders give the same registration"""
    g, _ = gold
    for m in ("Bunny", "Horse", "Dog"):
        assert int(g[m + "_canon_winner"]) == int(g[m + "_serial_winner"])
        R1 = g[m + "_canon_T"][:3, :3].astype(np.float64); R2 = g[m + "_serial_T"][:3, :3].astype(np.float64)
        assert np.arccos(np.clip((np.trace(R1.T @ R2) - 1) / 2, -1, 1)) < 1e-4
        assert abs(float(g[m + "_canon_rmse"]) - float(g[m + "_serial_rmse"])) < 1e-6


def test_angle_grid_quirk(okss):
    """B1: `for (a = 0; a < 6.3; a += 6.3/step)` gives 9 values at step 8 (729 hypotheses, not 512)"""
    acc, lst = okss.sweep_angles(8.0)
    assert len(acc) == 9 and acc[-1] < 6.3 and acc[-1] > 6.2999
    assert len(okss.sweep_angles(12.0)[0]) == 12 and len(okss.sweep_angles(6.0)[0]) == 6
    a = 0.0
    for v in acc:
        assert v == a
        a = a + 6.3 / 8.0
    assert np.array_equal(lst, np.arange(9) * 6.3 / 8.0)


def test_kdtree_equals_brute_force(okss):
    rng = np.random.default_rng(1)
    t = rng.normal(size=(4000, 3)).astype(np.float32); q = rng.normal(size=(3000, 3)).astype(np.float32)
    t[2000:2050] = t[10:60]
    q[:50] = t[10:60]
    i0, d0 = okss.nn(q, t, okss.NN_BRUTE); i1, d1 = okss.nn(q, t, okss.NN_KDTREE)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    assert (i0[:50] == np.arange(10, 60)).all()           # tie rule: lowest index
    d = ((q[:, None, :].astype(np.float32) - t[None]) ** 2)
    ref = (d[..., 0] + d[..., 1]) + d[..., 2]               # ((dx*dx)+dy*dy)+dz*dz in float32
    assert np.array_equal(d0, ref.min(1))


def _canon(v, dtype):
    v = np.asarray(v, dtype)
    if len(v) > 256:
        parts = [_canon(v[i:i + 256], dtype) for i in range(0, len(v), 256)]
        return _canon_chunk(np.array(parts, dtype), dtype) if len(parts) <= 256 else _canon(np.array(parts, dtype), dtype)
    return _canon_chunk(v, dtype)


def _canon_chunk(v, dtype):
    p = np.zeros(32, dtype)
    for i, x in enumerate(v):
        p[i & 31] = dtype(p[i & 31] + x)
    for off in (16, 8, 4, 2, 1):
        p = np.array([dtype(p[l] + p[l ^ off]) for l in range(32)], dtype)
    return p[0]


@pytest.mark.parametrize("n", [1, 31, 256, 257, 1024, 2000, 70000])
def test_canonical_sum_definition(okss, n):
    """CANON256: the order contract shared with the CUDA reductions"""
    rng = np.random.default_rng(n)
    v = rng.normal(size=n)
    assert okss.canon_sum_f64(v) == _canon(v, np.float64)
    assert okss.canon_sum_f32(v.astype(np.float32)) == float(_canon(v.astype(np.float32), np.float32))


def test_svd_and_umeyama(okss):
    rng = np.random.default_rng(3)
    for k in range(300):
        A = rng.normal(size=(3, 3)).astype(np.float32)
        if k % 4 == 0:
            A[:, 2] = A[:, 1]
        U, s, V = okss.svd3(A)
        assert np.abs((U * s) @ V.T - A).max() < 3e-6 * max(1.0, np.abs(A).max())
        assert s[0] >= s[1] >= s[2] >= 0
    R = _rot(0, 0.3) @ _rot(1, -0.7) @ _rot(2, 1.9)
    src = rng.normal(size=(400, 3)); dst = src @ R.T + np.array([0.3, -0.2, 0.1])
    for order in (okss.SUM_SERIAL, okss.SUM_CANON256):
        T = okss.umeyama(src, dst, order)
        assert np.abs(T[:3, :3] - R).max() < 2e-6 and np.abs(T[:3, 3] - [0.3, -0.2, 0.1]).max() < 2e-6
    refl = src * np.array([1, 1, -1.0])                      # mirrored target: det fix keeps a proper rotation
    T = okss.umeyama(src, refl)
    assert np.linalg.det(T[:3, :3].astype(np.float64)) > 0.999


def test_icp_convergence_rules(okss):
    rng = np.random.default_rng(4)
    t = rng.normal(size=(600, 3)).astype(np.float32).astype(np.float64) * 0.3
    s = t[:400] @ _rot(2, 0.05).T + 0.01
    r = okss.icp(s, t, trace_iters=64)
    assert r["converged"] == 1 and 1 <= r["iters"] < 64
    mse = r["trace"]["mse"][:r["iters"]]; Tk = r["trace"]["T_k"][:r["iters"]].reshape(-1, 4, 4)

    def fired(i):                                            # A.6 tests (2) and (3) after iteration i
        prev = mse[i - 1] if i > 0 else np.finfo(np.float64).max
        cos = 0.5 * float(np.float32(np.float32(Tk[i][0, 0] + Tk[i][1, 1]) + Tk[i][2, 2]) - np.float32(1))
        tr2 = float(np.float32(np.float32(Tk[i][0, 3] ** 2 + Tk[i][1, 3] ** 2) + Tk[i][2, 3] ** 2))
        return (cos >= 1 - 1e-10 and tr2 <= 1e-10) or abs(mse[i] - prev) < 1e-12 or abs(mse[i] - prev) / prev < 1e-3
    assert fired(r["iters"] - 1)
    assert not any(fired(i) for i in range(r["iters"] - 1))
    r3 = okss.icp(s, t, max_iter=2)
    assert r3["iters"] <= 2 and r3["converged"] == 1        # hasConverged() is true for "max iterations" too
    far = okss.icp(t[:5] + 50.0, t)
    assert far["converged"] == 0 and far["iters"] == 0      # < 3 correspondences within distance 1
    assert far["fitness"] > 1000


def test_minima_are_clamped_window_plateaus(okss, gold):
    """B5: non-periodic +-2 window, `centre > neighbour` rejects (plateaus count)"""
    g, _ = gold
    v = g["Bunny_value"]; G = v.shape[0]
    exp = []
    for i in range(G):
        for j in range(G):
            for k in range(G):
                w = v[max(0, i - 2):i + 3, max(0, j - 2):j + 3, max(0, k - 2):k + 3]
                if not (v[i, j, k] > w).any():
                    exp.append((i, j, k))
    assert np.array_equal(np.array(exp), g["Bunny_minima"])
    assert tuple(g["Bunny_best_index"]) == np.unravel_index(np.argmin(v), v.shape)


def test_pcr_qm_definition(okss):
    rng = np.random.default_rng(6)
    a = rng.normal(size=(300, 3)).astype(np.float32).astype(np.float64); t = rng.normal(size=(500, 3)).astype(np.float32).astype(np.float64)
    _, d2 = okss.nn(a, t, okss.NN_BRUTE)
    mse = 0.0; mae = 0.0
    for x in d2:
        mse += float(x); mae += float(np.sqrt(np.float64(x)))
    m = okss.nn_metrics(a, t)
    assert m[0] == mse / 300 and m[1] == np.sqrt(mse / 300) and m[2] == mae / 300


# ------------------------------------------------------------------ AIVS simplification
@pytest.mark.parametrize("model", ["Bunny", "Horse", "Dog"])
def test_aivs_matches_golden(okss, gold, model):
    g, fx = gold
    s = fx[model + "_src"].astype(np.float64); t = fx[model + "_tgt"].astype(np.float64)
    pn = min(len(s), len(t)) // 2
    out, idx = okss.aivs_simplify(s, pn)
    assert np.array_equal(idx, g[model + "_aivs_src_idx"])
    assert np.array_equal(out, s[idx])
    assert np.array_equal(okss.aivs_simplify(t, pn)[1], g[model + "_aivs_tgt_idx"])


def test_aivs_properties(okss):
    """exact target size after the trim, no point twice, an even spread: the largest hole of the simplified cloud
    stays within a small multiple of the mean sample spacing (what farthest point sampling is for)"""
    rng = np.random.default_rng(3)
    u = rng.normal(size=(4000, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    u[:2000] *= 1.0 + 0.05 * rng.random((2000, 1))              # uneven density
    out, idx = okss.aivs_simplify(u, 1000)
    assert len(idx) == 1000 and len(np.unique(idx)) == 1000
    assert np.array_equal(out, u[idx])
    d = np.sqrt(((u[:, None, :] - out[None, :, :]) ** 2).sum(-1)).min(1)
    spacing = np.sqrt(4 * np.pi / 1000)
    assert d.max() < 1.5 * spacing
    # pointNum above what the quotas give: no trim, fewer points
    few, fidx = okss.aivs_simplify(u, 20)
    assert len(fidx) <= 20


def test_aivs_box_scale_rule(okss):
    """EstimateBoxScale (ballRegionCompute.hpp:1194-1214): 10 boxes along the longest axis below 1e4 points -> a cloud
    on a 10-cell lattice with one point per cell and pointNum = n keeps every point"""
    g = np.stack(np.meshgrid(np.arange(10.0), np.arange(10.0), np.arange(10.0)), -1).reshape(-1, 3) + 0.5
    g = np.concatenate([g, [[0.0, 0.0, 0.0], [10.0, 10.0, 10.0]]])
    out, idx = okss.aivs_simplify(g, len(g))
    assert len(idx) == len(g) and len(np.unique(idx)) == len(g)


def test_raw_batch_equals_simplify_then_register(okss):
    rng = np.random.default_rng(4)
    P, n = 2, 600
    fs = rng.normal(size=(P, n, 3)); ft = fs[:, ::-1, :] * 1.3 + 0.2
    res, _ = okss.register_batch(None, None, fs, ft, threads=2)
    for p in range(P):
        a = okss.aivs_simplify(fs[p], n // 2)[0]; b = okss.aivs_simplify(ft[p], n // 2)[0]
        one = okss.register(a, b, fs[p], ft[p])
        assert res[p]["final_fitness"] == one["final_fitness"] and res[p]["rmse"] == one["rmse"]
        assert np.array_equal(res[p]["T"], one["T"])
