"""Deterministic synthetic point clouds for the BASELINE.json configurations.

Everything is numpy + an own splitmix64/xoshiro-free PRNG path: numpy's PCG64 seeded with
fixed integers (bit-reproducible across numpy versions for the distributions used here:
uniform and Box-Muller built on uniform; no np.random.normal).  Coordinates are rounded to
float32-representable doubles, because the reference's PLY loader delivers floats
(PlyLoad.cpp:93-101).

`simplify` is a seeded subset used only where a test wants GIVEN simplified clouds (the sim_s / sim_t mode of the
batch API); the registration path itself runs the real AIVS simplification (Method_AIVS_SimPro.hpp) on the device
(pNumber = min(|S|,|T|)/2 capped at 2000, KSS_ICP.hpp:57-66).
"""
import numpy as np

BASE_SEED = 0x4B5353494350  # "KSSICP"


def _rng(config, index):
    return np.random.Generator(np.random.PCG64(BASE_SEED + 1000003 * config + index))


def _normal(rng, shape):
    u1 = 1.0 - rng.random(shape)
    u2 = rng.random(shape)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def random_rotation(rng):
    q = _normal(rng, 4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def random_similarity(rng, scale_range=(0.5, 2.0), trans=1.0):
    R = random_rotation(rng)
    s = rng.uniform(*scale_range)
    t = rng.uniform(-trans, trans, 3)
    return R, s, t


def apply_similarity(p, R, s, t):
    return f32r(t + s * (p @ R.T))


def f32r(a):
    """round to float32-representable doubles"""
    return np.asarray(a, np.float64).astype(np.float32).astype(np.float64)


# ------------------------------------------------------------------ CAD-like shapes (config 3)
def _sample_box(rng, n, c, h):
    areas = np.array([h[1] * h[2], h[0] * h[2], h[0] * h[1]])
    face = rng.choice(3, size=n, p=areas / areas.sum())
    p = rng.uniform(-1, 1, (n, 3)) * h
    sign = np.where(rng.random(n) < 0.5, -1.0, 1.0)
    p[np.arange(n), face] = sign * h[face]
    return p + c


def _sample_cyl(rng, n, c, r, hh, axis):
    side = 2 * np.pi * r * 2 * hh
    cap = np.pi * r * r
    which = rng.random(n) < side / (side + 2 * cap)
    th = rng.uniform(0, 2 * np.pi, n)
    rr = np.where(which, r, r * np.sqrt(rng.random(n)))
    z = np.where(which, rng.uniform(-hh, hh, n), np.where(rng.random(n) < 0.5, -hh, hh))
    p = np.stack([rr * np.cos(th), rr * np.sin(th), z], 1)
    p = np.roll(p, axis, axis=1)
    return p + c


def _sample_plane(rng, n, c, h, axis):
    p = rng.uniform(-1, 1, (n, 3)) * h
    p[:, axis] = 0.0
    return p + c


def cad_shape_sampler(rng):
    """random union of 2-5 boxes / cylinders / planes; returns f(rng2, n) -> points"""
    k = int(rng.integers(2, 6))
    prims = []
    for _ in range(k):
        kind = int(rng.integers(0, 3))
        c = rng.uniform(-0.5, 0.5, 3)
        if kind == 0:
            h = rng.uniform(0.1, 0.6, 3)
            area = 8 * (h[0] * h[1] + h[1] * h[2] + h[0] * h[2])
            prims.append(("box", c, h, area))
        elif kind == 1:
            r = rng.uniform(0.1, 0.4); hh = rng.uniform(0.1, 0.7); ax = int(rng.integers(0, 3))
            area = 2 * np.pi * r * 2 * hh + 2 * np.pi * r * r
            prims.append(("cyl", c, (r, hh, ax), area))
        else:
            h = rng.uniform(0.2, 0.8, 3); ax = int(rng.integers(0, 3))
            hh = h.copy(); hh[ax] = 1.0
            area = 4 * np.prod(hh)
            prims.append(("plane", c, (h, ax), area))
    areas = np.array([p[3] for p in prims])
    probs = areas / areas.sum()

    def sample(rng2, n):
        counts = rng2.multinomial(n, probs)
        out = []
        for (kind, c, par, _), m in zip(prims, counts):
            if m == 0:
                continue
            if kind == "box":
                out.append(_sample_box(rng2, m, c, par))
            elif kind == "cyl":
                out.append(_sample_cyl(rng2, m, c, par[0], par[1], par[2]))
            else:
                out.append(_sample_plane(rng2, m, c, par[0], par[1]))
        p = np.concatenate(out, 0)
        return p[rng2.permutation(len(p))]
    return sample


def unit_sphere(p, ref=None):
    ref = p if ref is None else ref
    c = ref.mean(0)
    r = np.sqrt(((ref - c) ** 2).sum(1)).max()
    return (p - c) / r, c, r


def simplify(points, n, rng):
    """seeded-subset stand-in for AIVS_simplification(pNumber); keeps original order"""
    if n >= len(points):
        return points.copy()
    keep = np.sort(rng.permutation(len(points))[:n])
    return points[keep]


def modelnet_pair(index, n_full=2048, config=3):
    """config 3: one ModelNet40-shape pair (full_s, full_t, sim_s, sim_t, ground truth)"""
    rng = _rng(config, index)
    sampler = cad_shape_sampler(rng)
    t = sampler(rng, n_full)
    s = sampler(rng, n_full)                       # independent sampling of the same shape
    t, c, r = unit_sphere(t)
    s = (s - c) / r
    R, sc, tr = random_similarity(rng)
    full_t = f32r(t)
    full_s = apply_similarity(s, R, sc, tr)
    pn = min(len(full_s), len(full_t)) // 2        # KSS_ICP.hpp:57-66
    pn = min(pn, 2000)
    sim_t = simplify(full_t, pn, rng)
    sim_s = simplify(full_s, pn, rng)
    return dict(full_s=full_s, full_t=full_t, sim_s=sim_s, sim_t=sim_t, R=R, scale=sc, t=tr)


def modelnet_batch(n_pairs, n_full=2048, first=0, config=3):
    ps = [modelnet_pair(first + i, n_full, config) for i in range(n_pairs)]
    return {k: np.stack([p[k] for p in ps]) for k in ("full_s", "full_t", "sim_s", "sim_t")}, ps


# ------------------------------------------------------------------ smooth surface / scan (configs 2, 4)
def _blob_params(rng):
    k = 3
    return dict(cs=rng.uniform(-0.25, 0.25, (k, 3)), ax=rng.uniform(0.15, 0.45, (k, 3)), R0=0.35, r0=0.08)


def _blob_surface(rng, prm, n, nonuniform=True):
    """points on the union of 3 ellipsoids and a torus; density ~ exp(2x) when nonuniform"""
    k = len(prm["cs"])
    m = n // (k + 1)
    parts = []
    for i in range(k):
        v = _normal(rng, (m * 4, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
        parts.append(prm["cs"][i] + v * prm["ax"][i])
    q = (n - k * m) * 4
    th = rng.uniform(0, 2 * np.pi, q); ph = rng.uniform(0, 2 * np.pi, q)
    R0, r0 = prm["R0"], prm["r0"]
    parts.append(np.stack([(R0 + r0 * np.cos(ph)) * np.cos(th), (R0 + r0 * np.cos(ph)) * np.sin(th), r0 * np.sin(ph)], 1))
    p = np.concatenate(parts, 0)
    if nonuniform:                                  # accept with probability ~ exp(2x)
        w = np.exp(2.0 * (p[:, 0] - p[:, 0].max()))
        p = p[rng.random(len(p)) < w]
    p = p[rng.permutation(len(p))]
    while len(p) < n:                               # top up (rare): jittered copies
        m2 = min(len(p), n - len(p))
        p = np.concatenate([p, p[:m2] + 1e-4 * _normal(rng, (m2, 3))], 0)
    return p[:n]


def surface_pair(index, n, noise_sigma_rel=0.005, config=2):
    """config 2: two independent samplings of one procedural surface, non-uniform density,
    Gaussian noise (sigma = 0.005 * bbox diagonal) on the source only"""
    rng = _rng(config, index)
    prm = _blob_params(rng)
    t = _blob_surface(rng, prm, n)
    s = _blob_surface(rng, prm, n)
    diag = np.linalg.norm(t.max(0) - t.min(0))
    s = s + noise_sigma_rel * diag * _normal(rng, s.shape)
    R, sc, tr = random_similarity(rng)
    return dict(full_t=f32r(t), full_s=apply_similarity(s, R, sc, tr), R=R, scale=sc, t=tr)


def scan_pair(index, n, config=4, angle_deg=5.0, shift=0.02):
    """config 4: noisy terrain + objects, S = independent sampling perturbed by a small rigid motion"""
    rng = _rng(config, index)
    freq = rng.uniform(1.0, 4.0, (6, 2)); amp = rng.uniform(0.01, 0.05, 6); ph = rng.uniform(0, 6.28, 6)

    def height(xy):
        h = np.zeros(len(xy))
        for f, a, p in zip(freq, amp, ph):
            h += a * np.sin(f[0] * xy[:, 0] * 6.28 + f[1] * xy[:, 1] * 6.28 + p)
        return h
    obj_c = rng.uniform(0.1, 0.9, (12, 2)); obj_r = rng.uniform(0.02, 0.06, 12); obj_h = rng.uniform(0.03, 0.12, 12)

    def sample(r, m):
        xy = r.random((m, 2)) ** np.array([1.0, 1.5])          # non-uniform density
        z = height(xy)
        for c, rr, hh in zip(obj_c, obj_r, obj_h):
            d = np.linalg.norm(xy - c, axis=1)
            z = np.where(d < rr, z + hh * np.sqrt(np.maximum(0.0, 1 - (d / rr) ** 2)), z)
        p = np.concatenate([xy, z[:, None]], 1)
        return p + 0.0005 * _normal(r, p.shape)
    t = sample(rng, n)
    s = sample(rng, n)
    a = np.deg2rad(angle_deg)
    axis = _normal(rng, 3); axis /= np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * (K @ K)
    c = t.mean(0)
    s = (s - c) @ R.T + c + shift * axis
    return dict(full_t=f32r(t), full_s=f32r(s), R=R, t=shift * axis)


def transfer_pc(points, cord, angle, rate=1.0, dis=0.0):
    """the reference's TransferPC (transferPC.hpp:66-130): rotate about axis `cord` (1 x, 2 y, else z; the convention of
    initRegistration_Transfer), scale about the centroid, translate all three coordinates by `dis` -- in that order"""
    p = np.array(points, np.float64)
    c, s_ = np.cos(angle), np.sin(angle)
    x, y, z = p[:, 0].copy(), p[:, 1].copy(), p[:, 2].copy()
    if cord == 1:
        p[:, 1] = y * c - z * s_; p[:, 2] = y * s_ + z * c
    elif cord == 2:
        p[:, 0] = z * s_ + x * c; p[:, 2] = z * c - x * s_
    else:
        p[:, 0] = x * c - y * s_; p[:, 1] = x * s_ + y * c
    if rate != 1.0:
        m = np.array([np.add.reduce(p[:, a]) / len(p) for a in range(3)])
        p = (p - m) * rate + m
    if dis != 0.0:
        p = p + dis
    return p


def c1_pair(model_points, n=10000, index=1, config=1):
    """config 1 (BASELINE.json configs[0]): source = `model_points` (the reference's Armadillo.gird,
    43 871 points; tests/golden/fullsize_pairs.npz) decimated to n points by a seeded shuffle; target = a random
    similarity of an INDEPENDENT n-point subsample of the same file"""
    rng = _rng(config, index)
    pts = np.asarray(model_points, np.float64)
    src = pts[rng.permutation(len(pts))[:n]]
    tgt0 = pts[rng.permutation(len(pts))[:n]]
    R, sc, tr = random_similarity(rng)
    return dict(full_s=f32r(src), full_t=apply_similarity(tgt0, R, sc, tr), R=R, scale=sc, t=tr)


def c5_pair(model_points, keep_frac, index=0, n=10000, config=5):
    """config 5: a C1-style pair whose SOURCE is cropped by a random half-space to keep_frac of its points
    (partial / defective overlap)"""
    p = c1_pair(model_points, n=n, index=100 + index, config=config)
    rng = _rng(config, 1000 + index)
    p["full_s"] = crop_halfspace(p["full_s"], keep_frac, rng)
    p["keep_frac"] = keep_frac
    return p


def crop_halfspace(points, keep_frac, rng):
    """config 5: keep the `keep_frac` of points on one side of a random plane"""
    n = _normal(rng, 3); n /= np.linalg.norm(n)
    d = points @ n
    thr = np.quantile(d, 1.0 - keep_frac)
    return points[d >= thr]
