"""diagnostics: candidate-grid list lengths, per cell and weighted by the sweep's queries"""
import sys, os, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

P = 8
pkg = entry.load_package()
b, _ = pkg.synth.modelnet_batch(P, n_full=2048)
ctx = pkg.Context(0)
res = ctx.register_batch(None, None, b["full_s"], b["full_t"])
lib = ctx.lib
lib.kss_debug_read.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_size_t, C.c_void_p]
def rd(name, off, n, dt):
    a = np.empty(n, dt)
    rc = lib.kss_debug_read(ctx.h, name.encode(), off, a.nbytes, a.ctypes.data_as(C.c_void_p))
    assert rc == 0, (name, rc)
    return a
NG = 32; LV = [4, 8, 16, 32]; tot = sum(x ** 3 for x in LV); base = sum(x ** 3 for x in LV[:-1])
geom = rd("cg_geom", 0, P * 8, np.float32).reshape(P, 8)
cap_s = 1072
acc, lst = pkg.sweep_angles(8.0)
allw = []; allc = []
kinds = np.zeros(6, np.int64)     # inline, ext, huge, empty, (refined->) counted by the child
refq = 0; totq = 0
for p in range(P):
    hdr = rd("cg_hdr", (p * tot + base) * 8, NG ** 3, np.uint64)
    arena = rd("cg_arena", p * (3 << 19) * 2, 3 << 19, np.uint16)
    a64 = arena.view(np.uint64)
    wl = rd("cg_wl_cnt", p * 4, 1, np.uint32)[0]; cur = rd("cg_cursor", p * 4, 1, np.uint32)[0]
    if p == 0: print("pair 0: refined cells %d, arena used %d of %d" % (wl, cur, 3 << 19))
    s_al = rd("s_al", p * cap_s * 3 * 8, cap_s * 3, np.float64).reshape(cap_s, 3)
    cs = rd("aivs_cnt_s", p * 4, 1, np.int32)[0]
    s = s_al[:cs]
    lo = geom[p, :3] - geom[p, 3]; inv = NG / (2 * geom[p, 3])
    w = []
    rng = np.random.default_rng(p)
    for _ in range(40):
        a, bb, c = acc[rng.integers(0, 9, 3)]
        def rx(v, t): return np.stack([v[:, 0], v[:, 1] * np.cos(t) - v[:, 2] * np.sin(t), v[:, 1] * np.sin(t) + v[:, 2] * np.cos(t)], 1)
        def ry(v, t): return np.stack([v[:, 0] * np.cos(t) + v[:, 2] * np.sin(t), v[:, 1], -v[:, 0] * np.sin(t) + v[:, 2] * np.cos(t)], 1)
        def rz(v, t): return np.stack([v[:, 0] * np.cos(t) - v[:, 1] * np.sin(t), v[:, 0] * np.sin(t) + v[:, 1] * np.cos(t), v[:, 2]], 1)
        q = rz(ry(rx(s, a), bb), c)
        f = np.floor((q - lo) * inv).astype(int)
        ok = ((f >= 0) & (f < NG)).all(1)
        cell = f[:, 0] + NG * (f[:, 1] + NG * f[:, 2])
        cc = np.clip(cell, 0, NG ** 3 - 1)
        h = hdr[cc].copy()
        tag = (h >> np.uint64(60)).astype(np.int64)
        ref = tag == 15
        fr = (q - lo) * inv - f
        octv = (fr[:, 0] >= 0.5) * 1 + (fr[:, 1] >= 0.5) * 2 + (fr[:, 2] >= 0.5) * 4
        off64 = ((h & np.uint64(0xffffffff)).astype(np.int64) << 2) // 4      # u16 offset -> u64 index
        ch = a64[np.where(ref, off64 + octv, 0)]
        h = np.where(ref, ch, h)
        tag = (h >> np.uint64(60)).astype(np.int64)
        ref2 = tag == 15                                                      # second sparse level (128^3)
        fr2 = fr * 2 - np.floor(fr * 2)
        oct2 = (fr2[:, 0] >= 0.5) * 1 + (fr2[:, 1] >= 0.5) * 2 + (fr2[:, 2] >= 0.5) * 4
        off64 = ((h & np.uint64(0xffffffff)).astype(np.int64) << 2) // 4
        ch = a64[np.where(ref2, off64 + oct2, 0)]
        h = np.where(ref2, ch, h)
        tag = (h >> np.uint64(60)).astype(np.int64)
        ref2q = globals().get("ref2q", 0) + int((ref2 & ok).sum()); globals()["ref2q"] = ref2q
        k = np.where((tag >= 1) & (tag <= 5), tag, np.where(tag == 14, ((h >> np.uint64(32)) & np.uint64(0xfffffff)).astype(np.int64), np.where(tag == 13, 9999, 0)))
        kinds[0] += ((tag >= 1) & (tag <= 5) & ok).sum(); kinds[1] += ((tag == 14) & ok).sum(); kinds[2] += ((tag == 13) & ok).sum()
        kinds[3] += ((tag == 0) | ~ok).sum()
        refq += (ref & ok).sum(); totq += len(q)
        w.append(np.where(ok, k, -1))
    w = np.concatenate(w)
    allw.append(w); allc.append(np.zeros(1))
print("queries: inline %.2f%%, external %.2f%%, huge %.3f%%, empty/outside %.3f%%; through a refined cell %.1f%%, through two %.1f%%" % tuple(
    [100.0 * kinds[i] / totq for i in range(4)] + [100.0 * refq / totq, 100.0 * globals().get("ref2q", 0) / totq]))
w = np.concatenate(allw); c = np.concatenate(allc)
print("queries: outside/empty %.2f%%, mean cnt %.2f, mean padded %.2f" % (100 * (w <= 0).mean(), w[w > 0].mean(), (((w[w > 0] + 3) // 4) * 4).mean()))
print("query cnt percentiles 50/75/90/99/max:", np.percentile(w[w > 0], [50, 75, 90, 99, 100]))
print("query hist:", np.bincount(np.minimum(w[w > 0], 40))[:41])
pad = ((w[w > 0] + 3) // 4) * 4
k = len(pad) // 32 * 32
print("warp-max padded (random grouping, pessimistic): %.2f" % pad[:k].reshape(-1, 32).max(1).mean())
