"""kss_icp_b200 -- thin ctypes binding over the C ABI in include/kss_icp_b200.h.

This is test/bench plumbing only: the product is the shared library
(kss-icp_b200/libkss_icp_b200.so, built from csrc/ for sm_100a) and the C++ headers in
host/ that mirror the reference's KSS_ICP.hpp / initRegistrationKSS.hpp /
registrationMeasure.hpp classes.  There is no CPU fallback: a missing library or a missing
GPU raises.
"""
import ctypes as C
import os

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KSS_ICP_B200_LIB") or os.path.join(_DIR, "libkss_icp_b200.so")   # override: A/B builds in tools/
HEADER_PATH = os.path.join(os.path.dirname(_DIR), "include", "kss_icp_b200.h")

KSS_OK = 0
SMALL_MAX = 2048
SCORE_AVE, SCORE_MAX, SCORE_DIFF, SCORE_VOXEL = 0, 1, 2, 3
STAGES = ("prep", "sweep", "sweep_finalize", "icp_judge", "icp_hyp", "select_apply", "metrics",
          "large_build", "large_nn", "large_reduce", "cg_build", "aivs")


class KssError(RuntimeError):
    pass


class IcpParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("max_corr_dist", C.c_double),
                ("transformation_eps", C.c_double), ("fitness_eps", C.c_double)]


class IcpTrace(C.Structure):
    _fields_ = [("cap_iters", C.c_int), ("corr_idx", C.c_void_p), ("T_k", C.c_void_p),
                ("mse", C.c_void_p), ("src_k", C.c_void_p)]


class PairResult(C.Structure):
    _fields_ = [("align", C.c_double * 8), ("judge_fitness", C.c_double), ("final_fitness", C.c_double),
                ("mse", C.c_double), ("rmse", C.c_double), ("mae", C.c_double), ("T", C.c_float * 16),
                ("G", C.c_int), ("best_h", C.c_int), ("n_minima", C.c_int), ("branch_multi", C.c_int),
                ("winner", C.c_int), ("used_h", C.c_int), ("use_list", C.c_int), ("judge_iters", C.c_int),
                ("final_iters", C.c_int), ("total_icp_iters", C.c_int), ("n_icp_runs", C.c_int),
                ("overflow", C.c_int), ("final_converged", C.c_int), ("reserved_", C.c_int * 2)]


class Batch(C.Structure):
    _fields_ = [("n_pairs", C.c_int), ("cap_s", C.c_int), ("cap_t", C.c_int), ("cap_S", C.c_int), ("cap_T", C.c_int),
                ("sim_s", C.c_void_p), ("sim_t", C.c_void_p), ("full_s", C.c_void_p), ("full_t", C.c_void_p),
                ("cnt_s", C.c_void_p), ("cnt_t", C.c_void_p), ("cnt_S", C.c_void_p), ("cnt_T", C.c_void_p),
                ("step", C.c_double), ("icp", IcpParams), ("judge_threshold", C.c_double)]


RESULT_DTYPE = np.dtype([("align", "f8", 8), ("judge_fitness", "f8"), ("final_fitness", "f8"), ("mse", "f8"),
                         ("rmse", "f8"), ("mae", "f8"), ("T", "f4", 16), ("G", "i4"), ("best_h", "i4"),
                         ("n_minima", "i4"), ("branch_multi", "i4"), ("winner", "i4"), ("used_h", "i4"),
                         ("use_list", "i4"), ("judge_iters", "i4"), ("final_iters", "i4"),
                         ("total_icp_iters", "i4"), ("n_icp_runs", "i4"), ("overflow", "i4"),
                         ("final_converged", "i4"), ("reserved_", "i4", 2)], align=True)
assert RESULT_DTYPE.itemsize == C.sizeof(PairResult), (RESULT_DTYPE.itemsize, C.sizeof(PairResult))

_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KssError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.kss_last_error.restype = C.c_char_p
        _lib.kss_ctx_launch_count.restype = C.c_longlong
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def sweep_angles(step):
    lib = load_library()
    acc = np.empty(64, np.float64); lst = np.empty(64, np.float64)
    g = lib.kss_sweep_angles(C.c_double(step), _p(acc), _p(lst), C.c_int(64))
    return acc[:g].copy(), lst[:g].copy()


class Context:
    """One kss_ctx (one GPU, one stream, one host thread)."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.kss_ctx_create_on_stream(C.c_int(device), C.c_void_p(stream or 0), C.byref(self.h))
        if rc != KSS_OK:
            raise KssError("kss_ctx_create(device=%d) failed with %d (no CUDA device? there is no CPU fallback)"
                           % (device, rc))

    def close(self):
        if self.h:
            self.lib.kss_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != KSS_OK:
            raise KssError("kss error %d: %s" % (rc, self.lib.kss_last_error(self.h).decode()))

    def launch_count(self):
        return int(self.lib.kss_ctx_launch_count(self.h))

    def set_hyp_slots(self, slots):
        self._ck(self.lib.kss_ctx_set_hyp_slots(self.h, C.c_int(int(slots))))

    def set_timing(self, enable=True):
        self._ck(self.lib.kss_ctx_set_timing(self.h, C.c_int(1 if enable else 0)))

    def stage_ms(self, stage):
        ms = C.c_double(0); calls = C.c_longlong(0)
        self._ck(self.lib.kss_ctx_stage_ms(self.h, C.c_int(stage), C.byref(ms), C.byref(calls)))
        return ms.value, calls.value

    def synchronize(self):
        self._ck(self.lib.kss_ctx_synchronize(self.h))

    # ---- single-object calls ------------------------------------------------
    def middle_align(self, src, tgt):
        s = _f64(src); t = _f64(tgt)
        out7 = np.empty(7, np.float64); al = np.empty_like(s)
        self._ck(self.lib.kss_middle_align(self.h, _p(s), C.c_int(len(s)), _p(t), C.c_int(len(t)), _p(out7), _p(al)))
        return out7, al

    def rotation_sweep(self, src_aligned, tgt, step=8.0, score_mode=SCORE_AVE):
        s = _f64(src_aligned); t = _f64(tgt)
        g = len(sweep_angles(step)[0])
        value = np.empty(g ** 3, np.float64); G = C.c_int(0)
        best_angle = np.empty(3, np.float64); best_index = np.empty(3, np.int32)
        minima = np.empty((g ** 3, 3), np.int32); nmin = C.c_int(0)
        self._ck(self.lib.kss_rotation_sweep(self.h, _p(s), C.c_int(len(s)), _p(t), C.c_int(len(t)), C.c_double(step),
                                             C.c_int(score_mode), _p(value), C.byref(G), _p(best_angle),
                                             _p(best_index), _p(minima), C.byref(nmin)))
        return dict(G=G.value, value=value.reshape(g, g, g), best_angle=best_angle, best_index=best_index,
                    minima=minima[:nmin.value].copy())

    def apply_similarity(self, pts, align7, angles):
        p = _f64(pts); out = np.empty_like(p)
        self._ck(self.lib.kss_apply_similarity(self.h, _p(p), C.c_int(len(p)), _p(_f64(align7)), _p(_f64(angles)), _p(out)))
        return out

    def apply_transform(self, T, pts):
        p = _f64(pts); out = np.empty_like(p)
        T = np.ascontiguousarray(T, np.float32).reshape(-1)
        self._ck(self.lib.kss_apply_transform(self.h, _p(T), _p(p), C.c_int(len(p)), _p(out)))
        return out

    def icp(self, src, tgt, max_iter=1000, max_corr_dist=1.0, trans_eps=1e-10, fit_eps=1e-3, trace_iters=0):
        s = _f64(src); t = _f64(tgt)
        prm = IcpParams(max_iter, max_corr_dist, trans_eps, fit_eps)
        T = np.empty(16, np.float32); fit = C.c_double(0); it = C.c_int(0); cv = C.c_int(0)
        tr = None; keep = None
        if trace_iters > 0:
            keep = dict(corr_idx=np.full((trace_iters, len(s)), -2, np.int32),
                        T_k=np.zeros((trace_iters, 16), np.float32), mse=np.zeros(trace_iters, np.float64),
                        src_k=np.zeros((trace_iters, len(s), 3), np.float32))
            tr = IcpTrace(trace_iters, _p(keep["corr_idx"]), _p(keep["T_k"]), _p(keep["mse"]), _p(keep["src_k"]))
        self._ck(self.lib.kss_icp(self.h, _p(s), C.c_int(len(s)), _p(t), C.c_int(len(t)), C.byref(prm), _p(T),
                                  C.byref(fit), C.byref(it), C.byref(cv), C.byref(tr) if tr is not None else None))
        out = dict(fitness=fit.value, T=T.reshape(4, 4), iters=it.value, converged=cv.value)
        if keep is not None:
            out["trace"] = keep
        return out

    def icp_large_begin(self, src, tgt):
        s = _f64(src); t = _f64(tgt)
        self._ck(self.lib.kss_icp_large_begin(self.h, _p(s), C.c_int(len(s)), _p(t), C.c_int(len(t))))

    def icp_large_iterate(self, count, max_iter=1000, max_corr_dist=1.0, trans_eps=1e-10, fit_eps=1e-3):
        prm = IcpParams(max_iter, max_corr_dist, trans_eps, fit_eps)
        self._ck(self.lib.kss_icp_large_iterate(self.h, C.byref(prm), C.c_int(count)))

    def icp_large_end(self, max_iter=1000, max_corr_dist=1.0, trans_eps=1e-10, fit_eps=1e-3):
        prm = IcpParams(max_iter, max_corr_dist, trans_eps, fit_eps)
        T = np.empty(16, np.float32); fit = C.c_double(0); it = C.c_int(0); cv = C.c_int(0)
        self._ck(self.lib.kss_icp_large_end(self.h, C.byref(prm), _p(T), C.byref(fit), C.byref(it), C.byref(cv)))
        return dict(fitness=fit.value, T=T.reshape(4, 4), iters=it.value, converged=cv.value)

    def nn_metrics(self, a, t):
        a = _f64(a); t = _f64(t); out = np.empty(3, np.float64)
        self._ck(self.lib.kss_nn_metrics(self.h, _p(a), C.c_int(len(a)), _p(t), C.c_int(len(t)), _p(out)))
        return out

    def nn_search(self, q, t):
        q = _f64(q); t = _f64(t)
        idx = np.empty(len(q), np.int32); d2 = np.empty(len(q), np.float32)
        self._ck(self.lib.kss_nn_search(self.h, _p(q), C.c_int(len(q)), _p(t), C.c_int(len(t)), _p(idx), _p(d2)))
        return idx, d2

    # ---- batched registration -------------------------------------------------
    @staticmethod
    def _batch(P, caps, ptrs, cnts, step, max_iter, judge_threshold):
        b = Batch()
        b.n_pairs = P
        b.cap_s, b.cap_t, b.cap_S, b.cap_T = caps
        b.sim_s, b.sim_t, b.full_s, b.full_t = ptrs
        b.cnt_s, b.cnt_t, b.cnt_S, b.cnt_T = cnts
        b.step = step
        b.icp = IcpParams(max_iter, 1.0, 1e-10, 0.001)
        b.judge_threshold = judge_threshold
        return b

    def register_batch(self, sim_s, sim_t, full_s, full_t, step=8.0, max_iter=1000, counts=None,
                       want_points=False, judge_threshold=0.0005, results=None):
        """host arrays [P, cap, 3] float64 (numpy; pinned memory makes the copies asynchronous).
        sim_s = sim_t = None: the library simplifies full_s / full_t itself (AIVS, pNumber rule)."""
        fs = _f64(full_s); ft = _f64(full_t)
        raw = sim_s is None
        ss = None if raw else _f64(sim_s); st = None if raw else _f64(sim_t)
        P = fs.shape[0]
        cn = [None] * 4
        if counts is not None:
            cn = [np.ascontiguousarray(c, np.int32) if c is not None else None for c in counts]
        b = self._batch(P, (0 if raw else ss.shape[1], 0 if raw else st.shape[1], fs.shape[1], ft.shape[1]),
                        (None if raw else _p(ss).value, None if raw else _p(st).value, _p(fs).value, _p(ft).value),
                        tuple((_p(c).value if c is not None else None) for c in cn), step, max_iter, judge_threshold)
        res = results if results is not None else np.zeros(P, RESULT_DTYPE)
        pa = np.empty_like(fs) if want_points else None
        self._ck(self.lib.kss_register_batch(self.h, C.byref(b), _p(res), _p(pa)))
        return (res, pa) if want_points else res

    def aivs_simplify_batch(self, pts, point_num, counts=None, want_index=True):
        """pts [P, cap, 3] float64; point_num int or [P] ints -> (out [P, out_cap, 3], out_cnt [P], out_idx [P, out_cap])"""
        pts = _f64(pts); P, cap = pts.shape[0], pts.shape[1]
        each = None if np.isscalar(point_num) else np.ascontiguousarray(point_num, np.int32)
        out_cap = cap               # the greedy trim may stop early (stale neighbour lists): up to all samples can stay
        cn = None if counts is None else np.ascontiguousarray(counts, np.int32)
        out = np.zeros((P, out_cap, 3), np.float64); ocnt = np.zeros(P, np.int32)
        oidx = np.full((P, out_cap), -1, np.int32) if want_index else None
        self._ck(self.lib.kss_aivs_simplify_batch(self.h, C.c_int(P), _p(pts), _p(cn), C.c_int(cap), _p(each),
                                                  C.c_int(0 if each is not None else int(point_num)), _p(out),
                                                  C.c_int(out_cap), _p(ocnt), _p(oidx)))
        return out, ocnt, oidx

    def aivs_simplify(self, pts, point_num):
        """AIVS_simplification(point_num) of one cloud: (points [m, 3], positions in pts [m])"""
        out, cnt, idx = self.aivs_simplify_batch(_f64(pts)[None], int(point_num))
        return out[0, :cnt[0]].copy(), idx[0, :cnt[0]].copy()

    def aivs_simplify_batch_device(self, P, d_pts, cap, point_num, d_out, out_cap, d_out_cnt, d_cnt=None, d_point_num=None,
                                   d_out_idx=None):
        self._ck(self.lib.kss_aivs_simplify_batch_device(self.h, C.c_int(P), C.c_void_p(d_pts), C.c_void_p(d_cnt or 0),
                                                         C.c_int(cap), C.c_void_p(d_point_num or 0), C.c_int(point_num),
                                                         C.c_void_p(d_out), C.c_int(out_cap), C.c_void_p(d_out_cnt),
                                                         C.c_void_p(d_out_idx or 0)))

    def aivs_status(self):
        self._ck(self.lib.kss_aivs_status(self.h))

    def register_batch_device(self, P, caps, dev_ptrs, d_results, d_point_align=None, dev_counts=(None,) * 4,
                              step=8.0, max_iter=1000, judge_threshold=0.0005):
        """device pointers (ints) for [P, cap, 3] float64 clouds; results land in d_results (device)."""
        b = self._batch(P, caps, dev_ptrs, dev_counts, step, max_iter, judge_threshold)
        self._ck(self.lib.kss_register_batch_device(self.h, C.byref(b), C.c_void_p(d_results),
                                                    C.c_void_p(d_point_align or 0)))

    # ---- multi-GPU ---------------------------------------------------------------
    def nccl_init(self, unique_id, rank, world):
        """join the communicator of `unique_id` (128 bytes from nccl_unique_id() on rank 0) as `rank` of `world`"""
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.lib.kss_ctx_nccl_init(self.h, buf, C.c_int(rank), C.c_int(world)))

    def register_batch_hyp_sharded(self, sim_s, sim_t, full_s, full_t, step=8.0, max_iter=1000, counts=None,
                                   want_points=False, judge_threshold=0.0005):
        """register_batch with the rotation hypotheses of every pair sharded over the ranks of this context's
        communicator (every rank passes the same batch and gets the same results)"""
        fs = _f64(full_s); ft = _f64(full_t)
        raw = sim_s is None
        ss = None if raw else _f64(sim_s); st = None if raw else _f64(sim_t)
        P = fs.shape[0]
        cn = [None] * 4
        if counts is not None:
            cn = [np.ascontiguousarray(c, np.int32) if c is not None else None for c in counts]
        b = self._batch(P, (0 if raw else ss.shape[1], 0 if raw else st.shape[1], fs.shape[1], ft.shape[1]),
                        (None if raw else _p(ss).value, None if raw else _p(st).value, _p(fs).value, _p(ft).value),
                        tuple((_p(c).value if c is not None else None) for c in cn), step, max_iter, judge_threshold)
        res = np.zeros(P, RESULT_DTYPE)
        pa = np.empty_like(fs) if want_points else None
        self._ck(self.lib.kss_register_batch_hyp_sharded(self.h, C.byref(b), _p(res), _p(pa)))
        return (res, pa) if want_points else res

    def register(self, sim_s, sim_t, full_s, full_t, step=8.0, max_iter=1000, want_points=False):
        out = self.register_batch(_f64(sim_s)[None], _f64(sim_t)[None], _f64(full_s)[None], _f64(full_t)[None],
                                  step=step, max_iter=max_iter, want_points=want_points)
        if want_points:
            return out[0][0], out[1][0]
        return out[0]


def nccl_unique_id():
    """128 bytes identifying a new NCCL communicator (call on rank 0, hand to every rank's Context.nccl_init)"""
    buf = (C.c_ubyte * 128)()
    rc = load_library().kss_nccl_get_unique_id(buf)
    if rc != KSS_OK:
        raise KssError("kss_nccl_get_unique_id failed with %d (libnccl.so.2 not loadable?)" % rc)
    return bytes(buf)


def nccl_init_all(contexts):
    """one process, several GPUs: a communicator over `contexts` (rank = position)"""
    arr = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
    rc = load_library().kss_ctx_nccl_init_all(arr, C.c_int(len(contexts)))
    if rc != KSS_OK:
        raise KssError("kss_ctx_nccl_init_all failed with %d: %s" % (rc, load_library().kss_last_error(contexts[0].h).decode()))


def register_batch_multi(contexts, sim_s, sim_t, full_s, full_t, step=8.0, max_iter=1000, want_points=False,
                         judge_threshold=0.0005):
    """pairs in contiguous blocks over `contexts` (one per GPU), one host thread each inside the library"""
    fs = _f64(full_s); ft = _f64(full_t)
    raw = sim_s is None
    ss = None if raw else _f64(sim_s); st = None if raw else _f64(sim_t)
    P = fs.shape[0]
    b = Context._batch(P, (0 if raw else ss.shape[1], 0 if raw else st.shape[1], fs.shape[1], ft.shape[1]),
                       (None if raw else _p(ss).value, None if raw else _p(st).value, _p(fs).value, _p(ft).value),
                       (None,) * 4, step, max_iter, judge_threshold)
    res = np.zeros(P, RESULT_DTYPE)
    pa = np.empty_like(fs) if want_points else None
    arr = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
    rc = load_library().kss_register_batch_multi(arr, C.c_int(len(contexts)), C.byref(b), _p(res), _p(pa))
    if rc != KSS_OK:
        raise KssError("kss_register_batch_multi failed with %d: %s" % (rc, load_library().kss_last_error(contexts[0].h).decode()))
    return (res, pa) if want_points else res


def exported_symbols():
    """names the public header declares (used by the CPU-side ABI test)"""
    import re
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kss_[a-z0-9_]+)\s*\(", txt)))


from . import synth, dist  # noqa: E402,F401  (synthetic clouds; multi-GPU host logic)
