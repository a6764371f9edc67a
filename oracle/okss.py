"""ctypes binding of the CPU oracle (oracle/kss_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by kss-icp_b200/.
PARITY UNPINNED: see oracle/kss_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libkss_oracle.so")

NN_BRUTE, NN_KDTREE = 0, 1
SUM_SERIAL, SUM_CANON256 = 0, 1
SCORE_AVE, SCORE_MAX, SCORE_DIFF, SCORE_VOXEL = 0, 1, 2, 3


def build(force=False):
    src = os.path.join(_DIR, "kss_oracle.cpp")
    hdr = os.path.join(_DIR, "kss_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(_SO) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _DIR, "-B", "libkss_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


class IcpParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("max_corr_dist", C.c_double),
                ("transformation_eps", C.c_double), ("fitness_eps", C.c_double),
                ("sum_order", C.c_int), ("nn_method", C.c_int)]


class IcpTrace(C.Structure):
    _fields_ = [("cap_iters", C.c_int), ("corr_idx", C.c_void_p), ("T_k", C.c_void_p),
                ("mse", C.c_void_p), ("src_k", C.c_void_p)]


class PairResult(C.Structure):
    _fields_ = [("align", C.c_double * 7), ("best_angle", C.c_double * 3),
                ("best_index", C.c_int * 3), ("G", C.c_int), ("n_minima", C.c_int),
                ("branch_multi", C.c_int), ("winner", C.c_int),
                ("used_angle", C.c_double * 3), ("judge_fitness", C.c_double),
                ("final_fitness", C.c_double), ("judge_iters", C.c_int),
                ("final_iters", C.c_int), ("total_icp_iters", C.c_int),
                ("n_icp_runs", C.c_int), ("T", C.c_float * 16),
                ("mse", C.c_double), ("rmse", C.c_double), ("mae", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.okss_icp.restype = C.c_double
        _lib.okss_sweep_score.restype = C.c_double
        _lib.okss_canon_sum_f32.restype = C.c_float
        _lib.okss_canon_sum_f64.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def nn(q, t, method=NN_BRUTE):
    q = _f32(q); t = _f32(t)
    idx = np.empty(len(q), np.int32); d2 = np.empty(len(q), np.float32)
    lib().okss_nn(_p(q), C.c_int(len(q)), _p(t), C.c_int(len(t)), C.c_int(method), _p(idx), _p(d2))
    return idx, d2


def middle_align(src, tgt):
    src = _f64(src); tgt = _f64(tgt)
    out7 = np.empty(7, np.float64); al = np.empty_like(src)
    lib().okss_middle_align(_p(src), C.c_int(len(src)), _p(tgt), C.c_int(len(tgt)), _p(out7), _p(al))
    return out7, al


def sweep_angles(step):
    acc = np.empty(1024, np.float64); lst = np.empty(1024, np.float64)
    g = lib().okss_sweep_angles(C.c_double(step), _p(acc), _p(lst), C.c_int(1024))
    return acc[:g].copy(), lst[:g].copy()


def sweep(src_aligned, tgt, step=8.0, score_mode=SCORE_AVE, method=NN_KDTREE):
    s = _f64(src_aligned); t = _f64(tgt)
    g = len(sweep_angles(step)[0])
    value = np.empty(g * g * g, np.float64)
    best_angle = np.empty(3, np.float64); best_index = np.empty(3, np.int32)
    minima = np.empty((g * g * g, 3), np.int32); nmin = C.c_int(0)
    lib().okss_sweep(_p(s), C.c_int(len(s)), _p(t), C.c_int(len(t)), C.c_double(step),
                     C.c_int(score_mode), C.c_int(method), _p(value), _p(best_angle),
                     _p(best_index), _p(minima), C.byref(nmin))
    return dict(G=g, value=value.reshape(g, g, g), best_angle=best_angle, best_index=best_index,
                minima=minima[:nmin.value].copy())


def sweep_score(src_aligned, tgt, angles, score_mode=SCORE_AVE, method=NN_BRUTE):
    s = _f64(src_aligned); t = _f64(tgt); a = _f64(angles)
    q = np.empty((len(s), 3), np.float32); idx = np.empty(len(s), np.int32); d2 = np.empty(len(s), np.float32)
    v = lib().okss_sweep_score(_p(s), C.c_int(len(s)), _p(t), C.c_int(len(t)), _p(a),
                               C.c_int(score_mode), C.c_int(method), _p(q), _p(idx), _p(d2))
    return v, q, idx, d2


def apply_similarity(pts, align7, angles):
    pts = _f64(pts); out = np.empty_like(pts)
    lib().okss_apply_similarity(_p(pts), C.c_int(len(pts)), _p(_f64(align7)), _p(_f64(angles)), _p(out))
    return out


def icp(src, tgt, max_iter=1000, max_corr_dist=1.0, trans_eps=1e-10, fit_eps=1e-3,
        sum_order=SUM_CANON256, method=NN_KDTREE, trace_iters=0):
    s = _f64(src); t = _f64(tgt)
    p = IcpParams(max_iter, max_corr_dist, trans_eps, fit_eps, sum_order, method)
    T = np.empty(16, np.float32); it = C.c_int(0); cv = C.c_int(0)
    tr = None; keep = None
    if trace_iters > 0:
        keep = dict(corr_idx=np.full((trace_iters, len(s)), -2, np.int32),
                    T_k=np.zeros((trace_iters, 16), np.float32),
                    mse=np.zeros(trace_iters, np.float64),
                    src_k=np.zeros((trace_iters, len(s), 3), np.float32))
        tr = IcpTrace(trace_iters, _p(keep["corr_idx"]), _p(keep["T_k"]), _p(keep["mse"]), _p(keep["src_k"]))
    fit = lib().okss_icp(_p(s), C.c_int(len(s)), _p(t), C.c_int(len(t)), C.byref(p), _p(T),
                         C.byref(it), C.byref(cv), C.byref(tr) if tr is not None else None)
    out = dict(fitness=fit, T=T.reshape(4, 4), iters=it.value, converged=cv.value)
    if keep is not None:
        out["trace"] = keep
    return out


def icp_iteration(src_f32, tgt_f32, max_corr_dist=1.0, sum_order=SUM_CANON256, method=NN_KDTREE):
    s = _f32(src_f32); t = _f32(tgt_f32)
    idx = np.empty(len(s), np.int32); d2 = np.empty(len(s), np.float32)
    T = np.empty(16, np.float32); mse = C.c_double(0); out = np.empty_like(s)
    cnt = lib().okss_icp_iteration(_p(s), C.c_int(len(s)), _p(t), C.c_int(len(t)),
                                   C.c_double(max_corr_dist), C.c_int(sum_order), C.c_int(method),
                                   _p(idx), _p(d2), _p(T), C.byref(mse), _p(out))
    return dict(n_corr=cnt, idx=idx, d2=d2, T=T.reshape(4, 4), mse=mse.value, src_out=out)


def apply_transform(T, pts):
    pts = _f64(pts); out = np.empty_like(pts)
    lib().okss_apply_transform(_p(_f32(T).reshape(-1)), _p(pts), C.c_int(len(pts)), _p(out))
    return out


def nn_metrics(a, t, method=NN_KDTREE):
    a = _f64(a); t = _f64(t); out = np.empty(3, np.float64)
    lib().okss_nn_metrics(_p(a), C.c_int(len(a)), _p(t), C.c_int(len(t)), C.c_int(method), _p(out))
    return out


def umeyama(src, dst, sum_order=SUM_SERIAL):
    s = _f32(src); d = _f32(dst); T = np.empty(16, np.float32)
    lib().okss_umeyama(_p(s), _p(d), C.c_int(len(s)), C.c_int(sum_order), _p(T))
    return T.reshape(4, 4)


def svd3(A):
    A = _f32(A).reshape(-1); U = np.empty(9, np.float32); s = np.empty(3, np.float32); V = np.empty(9, np.float32)
    lib().okss_svd3(_p(A), _p(U), _p(s), _p(V))
    return U.reshape(3, 3), s, V.reshape(3, 3)


def aivs_simplify(pts, point_num):
    """AIVS_simplification(pointNum) on the BallRegion of `pts`: returns (points [m,3], original indices [m])"""
    pts = _f64(pts)
    out = np.empty_like(pts); idx = np.empty(len(pts), np.int32)
    m = lib().okss_aivs_simplify(_p(pts), C.c_int(len(pts)), C.c_int(point_num), _p(out), _p(idx))
    return out[:m].copy(), idx[:m].copy()


def canon_sum_f32(v):
    v = _f32(v); return float(lib().okss_canon_sum_f32(_p(v), C.c_int(len(v))))


def canon_sum_f64(v):
    v = _f64(v); return float(lib().okss_canon_sum_f64(_p(v), C.c_int(len(v))))


def _res_dict(r):
    return dict(align=np.array(r.align[:]), best_angle=np.array(r.best_angle[:]),
                best_index=np.array(r.best_index[:]), G=r.G, n_minima=r.n_minima,
                branch_multi=r.branch_multi, winner=r.winner, used_angle=np.array(r.used_angle[:]),
                judge_fitness=r.judge_fitness, final_fitness=r.final_fitness,
                judge_iters=r.judge_iters, final_iters=r.final_iters,
                total_icp_iters=r.total_icp_iters, n_icp_runs=r.n_icp_runs,
                T=np.array(r.T[:], np.float32).reshape(4, 4), mse=r.mse, rmse=r.rmse, mae=r.mae)


def register(sim_s, sim_t, full_s, full_t, step=8.0, max_iter=1000,
             sum_order=SUM_CANON256, method=NN_KDTREE, want_points=False):
    ss = _f64(sim_s); st = _f64(sim_t); fs = _f64(full_s); ft = _f64(full_t)
    r = PairResult(); pa = np.empty_like(fs) if want_points else None
    lib().okss_register(_p(ss), C.c_int(len(ss)), _p(st), C.c_int(len(st)), _p(fs), C.c_int(len(fs)),
                        _p(ft), C.c_int(len(ft)), C.c_double(step), C.c_int(max_iter),
                        C.c_int(sum_order), C.c_int(method), C.byref(r), _p(pa))
    d = _res_dict(r)
    if want_points:
        d["point_align"] = pa
    return d


def register_batch(sim_s, sim_t, full_s, full_t, step=8.0, max_iter=1000,
                   sum_order=SUM_CANON256, method=NN_KDTREE, threads=0):
    """arrays shaped [P, n, 3]; returns (list of result dicts, threads used).
    sim_s = sim_t = None: raw clouds, AIVS-simplified per pair inside (KSSICP_init + KSSICP_Registration)."""
    fs = _f64(full_s); ft = _f64(full_t)
    raw = sim_s is None
    ss = None if raw else _f64(sim_s); st = None if raw else _f64(sim_t)
    P = fs.shape[0]
    res = (PairResult * P)()
    used = lib().okss_register_batch(C.c_int(P), _p(ss), C.c_int(0 if raw else ss.shape[1]), _p(st),
                                     C.c_int(0 if raw else st.shape[1]),
                                     _p(fs), C.c_int(fs.shape[1]), _p(ft), C.c_int(ft.shape[1]),
                                     C.c_double(step), C.c_int(max_iter), C.c_int(sum_order),
                                     C.c_int(method), C.c_int(threads), res)
    return [_res_dict(r) for r in res], used


def max_threads():
    return lib().okss_max_threads()
