"""GPU parity tests for the AIVS simplification: the C ABI against the CPU oracle's
restatement of BallRegion + AIVS_simplification on the same clouds.

Bar: the kept points are the same input positions in the same order (bit-exact indices, hence
bit-exact coordinates), for every cloud of a batch; and a raw-cloud registration equals the
registration of the oracle-simplified clouds bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _shape_cloud(pkg, index, n):
    p = pkg.synth.modelnet_pair(index, n_full=n)
    return p["full_s"], p["full_t"]


def _check(ctx, okss, pts, point_num):
    out, idx = ctx.aivs_simplify(pts, point_num)
    o_out, o_idx = okss.aivs_simplify(pts, point_num)
    assert len(idx) == len(o_idx)
    assert np.array_equal(idx, o_idx)
    assert np.array_equal(out, o_out)
    assert np.array_equal(out, pts[idx])
    return idx


@pytest.mark.parametrize("n,point_num", [(2048, 1024), (2048, 700), (1000, 500), (300, 150), (5000, 2000), (64, 32), (9, 4)])
def test_aivs_single_cloud_bit_exact(ctx, okss, pkg, n, point_num):
    s, t = _shape_cloud(pkg, n + point_num, n)
    for c in (s, t):
        idx = _check(ctx, okss, c, point_num)
        assert len(np.unique(idx)) == len(idx)


def test_aivs_gaussian_and_scaled_clouds(ctx, okss):
    rng = np.random.default_rng(11)
    for scale, shift in ((1.0, 0.0), (250.0, 1000.0), (1e-3, -5.0)):
        pts = rng.normal(size=(3000, 3)) * scale + shift
        _check(ctx, okss, pts, 1200)


def test_aivs_flat_and_line_like_clouds(ctx, okss):
    """thin clouds give one or two boxes along an axis: exercises the box decode quirks"""
    rng = np.random.default_rng(12)
    plane = rng.uniform(-1, 1, (2500, 3)); plane[:, 2] *= 1e-3
    _check(ctx, okss, plane, 1000)
    plane2 = rng.uniform(-1, 1, (2500, 3)); plane2[:, 0] *= 0.05
    _check(ctx, okss, plane2, 1000)
    rod = rng.uniform(-1, 1, (1500, 3)); rod[:, :2] *= 0.02
    _check(ctx, okss, rod, 600)


def test_aivs_duplicates_and_lattice(ctx, okss):
    rng = np.random.default_rng(13)
    g = np.stack(np.meshgrid(np.arange(12.0), np.arange(12.0), np.arange(12.0)), -1).reshape(-1, 3)
    _check(ctx, okss, g, 800)                                   # equal distances everywhere: first-strict rules decide
    pts = rng.normal(size=(1500, 3))
    pts[700:900] = pts[100:300]                                 # exact duplicates
    _check(ctx, okss, pts, 700)


def test_aivs_more_samples_than_fit_one_block(ctx, okss):
    """>= 1e5 points: 40 boxes per axis, several thousand one-sample boxes, thousands of trim steps"""
    rng = np.random.default_rng(14)
    u = rng.normal(size=(120000, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    u[:, 1] *= 0.6
    idx = _check(ctx, okss, u, 2000)
    assert len(idx) == 2000


def test_aivs_quota_shortfall(ctx, okss):
    """a tiny rate leaves boxes without a quota: fewer samples than asked for (even none), no trim"""
    rng = np.random.default_rng(15)
    u = rng.normal(size=(20000, 3)); u[:, 0] *= 2; u /= np.linalg.norm(u, axis=1, keepdims=True)
    assert 0 < len(_check(ctx, okss, u, 80)) < 80
    assert 0 < len(_check(ctx, okss, u, 100)) < 100
    assert len(_check(ctx, okss, u, 40)) == 0
    assert len(_check(ctx, okss, rng.uniform(-1, 1, (20000, 3)), 300)) == 0


def test_aivs_batch_ragged(ctx, okss, pkg):
    P, cap = 12, 2048
    rng = np.random.default_rng(16)
    pts = np.zeros((P, cap, 3)); cnt = np.zeros(P, np.int32); pn = np.zeros(P, np.int32)
    for p in range(P):
        n = int(rng.integers(200, cap + 1)) if p else cap
        s, _ = _shape_cloud(pkg, 900 + p, n)
        pts[p, :n] = s; pts[p, n:] = 1e6                        # padding must never be read
        cnt[p] = n; pn[p] = min(2000, n // 2)
    out, ocnt, oidx = ctx.aivs_simplify_batch(pts, pn, counts=cnt)
    for p in range(P):
        o_out, o_idx = okss.aivs_simplify(pts[p, :cnt[p]], int(pn[p]))
        assert ocnt[p] == len(o_idx)
        assert np.array_equal(oidx[p, :ocnt[p]], o_idx)
        assert np.array_equal(out[p, :ocnt[p]], o_out)


def test_aivs_degenerate_cloud_is_an_error(ctx, pkg):
    pts = np.ones((100, 3))
    with pytest.raises(pkg.KssError):
        ctx.aivs_simplify(pts, 50)
    # the context stays usable
    rng = np.random.default_rng(17)
    out, idx = ctx.aivs_simplify(rng.normal(size=(500, 3)), 250)
    assert len(idx) > 0


def test_register_raw_equals_register_of_simplified(ctx, okss, pkg):
    """kss_batch with sim_s = sim_t = NULL: AIVS on the device, then the same path"""
    P, n = 6, 2048
    b, _ = pkg.synth.modelnet_batch(P, n_full=n, first=40)
    fs, ft = b["full_s"], b["full_t"]
    raw, pa_raw = ctx.register_batch(None, None, fs, ft, want_points=True)
    pn = n // 2
    ss = np.zeros((P, pn, 3)); st = np.zeros((P, pn, 3)); cs = np.zeros(P, np.int32); ct = np.zeros(P, np.int32)
    for p in range(P):
        a = okss.aivs_simplify(fs[p], pn)[0]; c = okss.aivs_simplify(ft[p], pn)[0]
        ss[p, :len(a)] = a; st[p, :len(c)] = c; cs[p], ct[p] = len(a), len(c)
    ref, pa_ref = ctx.register_batch(ss, st, fs, ft, counts=(cs, ct, None, None), want_points=True)
    for k in raw.dtype.names:
        assert np.array_equal(raw[k], ref[k]), k
    assert np.array_equal(pa_raw, pa_ref)
    # and against the oracle end to end for one pair
    o = okss.register(ss[0, :cs[0]], st[0, :ct[0]], fs[0], ft[0])
    assert raw[0]["winner"] == o["winner"] and raw[0]["n_minima"] == o["n_minima"]
    assert raw[0]["final_fitness"] == o["final_fitness"]
    assert raw[0]["rmse"] == o["rmse"]


def test_register_raw_ragged_counts(ctx, okss, pkg):
    P, cap = 4, 2048
    rng = np.random.default_rng(18)
    fs = np.zeros((P, cap, 3)); ft = np.zeros((P, cap, 3)); cS = np.zeros(P, np.int32); cT = np.zeros(P, np.int32)
    for p in range(P):
        nS, nT = int(rng.integers(600, cap + 1)), int(rng.integers(600, cap + 1))
        pr = pkg.synth.modelnet_pair(70 + p, n_full=cap)
        fs[p, :nS] = pr["full_s"][:nS]; ft[p, :nT] = pr["full_t"][:nT]
        cS[p], cT[p] = nS, nT
    raw = ctx.register_batch(None, None, fs, ft, counts=(None, None, cS, cT))
    for p in range(P):
        pn = min(2000, min(cS[p], cT[p]) // 2)
        ss = okss.aivs_simplify(fs[p, :cS[p]], pn)[0]; st = okss.aivs_simplify(ft[p, :cT[p]], pn)[0]
        one = ctx.register(ss, st, fs[p, :cS[p]], ft[p, :cT[p]])
        for k in ("final_fitness", "judge_fitness", "rmse", "mse", "mae", "winner", "n_minima"):
            assert raw[p][k] == one[k], (p, k)
        assert np.array_equal(raw[p]["T"], one["T"])


def test_aivs_general_kernels_match_single_cta_kernel(ctx, okss, pkg, monkeypatch):
    """clouds <= 2048 points normally take the one-CTA-per-cloud kernel; KSS_AIVS_GENERAL=1 forces the multi-kernel
    path used for large clouds: both must keep the same points"""
    P, cap = 5, 1500
    rng = np.random.default_rng(19)
    pts = np.zeros((P, cap, 3)); cnt = np.zeros(P, np.int32)
    for p in range(P):
        n = int(rng.integers(100, cap + 1))
        pts[p, :n] = _shape_cloud(pkg, 300 + p, n)[0]; cnt[p] = n
    a = ctx.aivs_simplify_batch(pts, 600, counts=cnt)
    monkeypatch.setenv("KSS_AIVS_GENERAL", "1")
    b = ctx.aivs_simplify_batch(pts, 600, counts=cnt)
    monkeypatch.delenv("KSS_AIVS_GENERAL")
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    for p in range(P):
        assert np.array_equal(a[2][p, :a[1][p]], okss.aivs_simplify(pts[p, :cnt[p]], 600)[1])


def test_aivs_many_random_shapes(ctx, okss, pkg):
    """96 clouds with random anisotropy and sizes: any dependence on the order in which same-colour boxes are
    processed (misplaced centres of the boxes at the end of an x/y row) would show up here"""
    P, cap = 96, 2048
    rng = np.random.default_rng(20)
    pts = np.zeros((P, cap, 3)); cnt = np.zeros(P, np.int32); pn = np.zeros(P, np.int32)
    for p in range(P):
        n = int(rng.integers(300, cap + 1))
        kind = p % 4
        if kind == 0:
            c = _shape_cloud(pkg, 500 + p, n)[0]
        elif kind == 1:
            c = rng.uniform(-1, 1, (n, 3))
        elif kind == 2:
            c = rng.normal(size=(n, 3)); c /= np.linalg.norm(c, axis=1, keepdims=True)
        else:
            c = rng.normal(size=(n, 3))
        c = c * rng.uniform(0.05, 1.0, 3) * rng.choice([1.0, 1.0, 0.02], 3)
        pts[p, :n] = c; cnt[p] = n; pn[p] = int(n * rng.uniform(0.2, 0.6))
    out, ocnt, oidx = ctx.aivs_simplify_batch(pts, pn, counts=cnt)
    for p in range(P):
        o_idx = okss.aivs_simplify(pts[p, :cnt[p]], int(pn[p]))[1]
        assert ocnt[p] == len(o_idx), p
        assert np.array_equal(oidx[p, :ocnt[p]], o_idx), p


@pytest.mark.parametrize("model", ["Bunny", "Horse", "Dog"])
def test_aivs_golden_fixtures_from_reference_data(ctx, model):
    """the reference's own shipped clouds (decimated fixtures) against the committed golden vectors: AIVS index lists
    and the raw-cloud registration behind them, no oracle at run time"""
    import os
    gd = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gd, "golden_oracle.npz")); fx = np.load(os.path.join(gd, "fixtures_pairs.npz"))
    s = fx[model + "_src"].astype(np.float64); t = fx[model + "_tgt"].astype(np.float64)
    pn = min(len(s), len(t)) // 2
    assert np.array_equal(ctx.aivs_simplify(s, pn)[1], g[model + "_aivs_src_idx"])
    assert np.array_equal(ctx.aivs_simplify(t, pn)[1], g[model + "_aivs_tgt_idx"])
    cap = max(len(s), len(t))
    fs = np.zeros((1, cap, 3)); ft = np.zeros((1, cap, 3)); fs[0, :len(s)] = s; ft[0, :len(t)] = t
    r = ctx.register_batch(None, None, fs, ft, counts=(None, None, np.array([len(s)], np.int32), np.array([len(t)], np.int32)))[0]
    assert int(r["winner"]) == int(g[model + "_raw_winner"]) and int(r["n_minima"]) == int(g[model + "_raw_n_minima"])
    assert np.array_equal(r["T"].reshape(4, 4), g[model + "_raw_T"])
    for k in ("final_fitness", "mse", "rmse", "mae"):
        assert float(r[k]) == float(g[model + "_raw_" + k]), k


def test_register_raw_large_clouds(ctx, okss, pkg):
    """raw clouds beyond the shared-memory path (20k points, pNumber capped at 2000): the any-size AIVS kernels, then
    the same registration; final metrics through the large-cloud NN"""
    n = 20000
    pr = pkg.synth.modelnet_pair(5, n_full=n)
    fs, ft = pr["full_s"], pr["full_t"]
    raw = ctx.register_batch(None, None, fs[None], ft[None])[0]
    ss = okss.aivs_simplify(fs, 2000)[0]; st = okss.aivs_simplify(ft, 2000)[0]
    one = ctx.register(ss, st, fs, ft)
    for k in ("final_fitness", "judge_fitness", "winner", "n_minima", "total_icp_iters"):
        assert raw[k] == one[k], k
    assert np.array_equal(raw["T"], one["T"])
    for k in ("mse", "rmse", "mae"):
        assert abs(raw[k] - one[k]) <= 1e-12 * max(1.0, abs(one[k])), k


def test_aivs_big_boxes_take_the_cta_path(ctx, okss, pkg):
    """clouds from 250,000 points on hand the boxes with more than 512 members to aivs_fps_big_kernel (one CTA per box,
    members staged in shared memory) and stage the others per warp: a non-uniform 300k-point scan has both kinds -- the
    kept points are still the oracle's, index for index"""
    p = pkg.synth.scan_pair(7, 300000)
    idx = _check(ctx, okss, p["full_t"], 2000)
    assert len(np.unique(idx)) == len(idx)
