"""distribution of ICP run lengths in the ModelNet40-shape batch and what launch order would cost: list-scheduling model
(S concurrent CTA slots, time of a run = its iteration count) for the current order, longest-first by a predictor, and
the true longest-first"""
import ctypes as C
import heapq
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("KSS_LANES", "1"); os.environ.setdefault("KSS_CHUNKS", "1")
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
P = int(sys.argv[1]) if len(sys.argv) > 1 else 309
b, _ = pkg.synth.modelnet_batch(P, n_full=2048)
ctx = pkg.Context(0)
res = ctx.register_batch(None, None, b["full_s"], b["full_t"])
R = 33


def read(name, dtype, count, sfx):
    buf = np.zeros(count, dtype)
    rc = ctx.lib.kss_debug_read(ctx.h, (name + sfx).encode(), C.c_size_t(0), C.c_size_t(buf.nbytes), buf.ctypes.data_as(C.c_void_p))
    return buf if rc == 0 else None


for sfx in ("", "@0", "#0", ":0", ".0", "_0"):
    it = read("run_iters", np.int32, P * R, sfx)
    if it is not None:
        break
assert it is not None, "run_iters not found"
it = it.reshape(P, R)
nm = np.array([int(r["n_minima"]) for r in res])
runs = []            # (pair, slot, iters) in launch order
for p in range(P):
    for s in range(R):
        if s == 0 or s - 1 < nm[p]:
            runs.append((p, s, int(it[p, s])))
L = np.array([r[2] for r in runs])
print("pairs %d runs %d  iterations: mean %.1f median %d p90 %d p99 %d max %d  total %d" % (P, len(L), L.mean(), np.median(L), np.percentile(L, 90), np.percentile(L, 99), L.max(), L.sum()))


def schedule(order, S):
    h = [0.0] * S
    heapq.heapify(h)
    end = 0.0
    for i in order:
        t = heapq.heappop(h) + L[i]
        end = max(end, t)
        heapq.heappush(h, t)
    return end


# predictors: the sweep score of the start hypothesis (mean NN distance after the start rotation)
G3 = 729
val = read("value", np.float64, P * G3, sfx)
hp = int(os.environ.get("HPAD", "0")) or None
bh = read("best_h", np.int32, P, sfx)
pred = np.zeros(len(runs))
if val is not None and bh is not None:
    val = val.reshape(P, G3)
    for cap in (729, 736, 768, 1024):
        mn = read("minima", np.int32, P * cap, sfx)
        if mn is None:
            continue
        mn = mn.reshape(P, cap)
        if all(0 <= mn[p, :nm[p]].min() and mn[p, :nm[p]].max() < G3 for p in range(P)):
            break
    for k, (p_, s_, l_) in enumerate(runs):
        h = bh[p_] if s_ == 0 else mn[p_, s_ - 1]
        pred[k] = val[p_, h] / max(val[p_].min(), 1e-30)
    print("correlation(iters, score / best score of the pair) = %.3f ; judge runs mean %.1f, hypothesis runs mean %.1f" % (
        np.corrcoef(L, pred)[0, 1], L[[r[1] == 0 for r in runs]].mean(), L[[r[1] != 0 for r in runs]].mean()))
    for lo, hi in ((1.0, 1.0001), (1.0001, 1.05), (1.05, 1.1), (1.1, 1.2), (1.2, 1.5), (1.5, 9)):
        m = (pred >= lo) & (pred < hi)
        if m.any():
            print("  score ratio [%.4f, %.4f): %5d runs, mean iters %.1f, p90 %d, max %d" % (lo, hi, m.sum(), L[m].mean(), np.percentile(L[m], 90), L[m].max()))
for S in (592, 444):
    ideal = L.sum() / S
    print("slots %d: ideal %.0f  launch order %.0f  true longest-first %.0f  (longest run %d)" % (
        S, ideal, schedule(range(len(L)), S), schedule(np.argsort(-L), S), L.max()))
    print("   by descending score ratio %.0f, ascending %.0f" % (schedule(np.argsort(-pred), S), schedule(np.argsort(pred), S)))
