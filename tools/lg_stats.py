"""steady-state diagnostics of the 1M-point ICP iteration: per-iteration counts of second-stage searches, pyramid
fallbacks and CTAs that could not stage their region (LgState::miss, kss_large.cu), plus per-kernel event times"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
p = pkg.synth.scan_pair(0, N)
ctx = pkg.Context(0)
never = dict(max_iter=1 << 30, fit_eps=-1.0, trans_eps=-1.0)


def step():
    """largest displacement T_k - I causes over the source's bounding box, in units of the target's median NN distance"""
    buf = np.zeros(16, np.float32)
    rc = ctx.lib.kss_debug_read(ctx.h, b"run:lg_state", C.c_size_t(0), C.c_size_t(64), buf.ctypes.data_as(C.c_void_p))
    T = buf.reshape(4, 4).astype(np.float64)
    lo, hi = p["full_s"].min(0), p["full_s"].max(0)
    cs = np.array([[x, y, z, 1.0] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    return np.linalg.norm((cs @ T.T - cs)[:, :3], axis=1).max()


def miss():
    buf = np.zeros(60, np.uint32)
    rc = ctx.lib.kss_debug_read(ctx.h, b"run:lg_state", C.c_size_t(0), C.c_size_t(240), buf.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return buf[54:59].astype(np.int64)


ctx.icp_large_begin(p["full_s"], p["full_t"])
ctx.set_timing(True)
ctx.icp_large_begin(p["full_s"], p["full_t"])
ctx.synchronize()
print("build ms", ctx.stage_ms(7)[0])
import torch  # noqa: E402
prev = miss()
NIT = int(sys.argv[2]) if len(sys.argv) > 2 else 12
for it in range(NIT):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    ctx.set_timing(True)
    e0.record(torch.cuda.current_stream())
    ctx.icp_large_iterate(1, **never)
    e1.record(torch.cuda.current_stream())
    ctx.synchronize()
    m = miss()
    print("step %.2e " % step(), end="")
    print("iter %2d: %.3f ms (nn %.3f reduce %.3f)  list-settled %8d  searched %8d  pyramid %8d  global-path CTAs %6d" % (
        it, e0.elapsed_time(e1), ctx.stage_ms(8)[0], ctx.stage_ms(9)[0], m[0] - prev[0], m[1] - prev[1], m[2] - prev[2], m[3] - prev[3]))
    prev = m
ctx.icp_large_iterate(30, **never)
ctx.synchronize()
ctx.set_timing(True)
prev = miss()
K = 50
ctx.icp_large_iterate(K, **never)
ctx.synchronize()
m = miss()
nn, c = ctx.stage_ms(8); rd, _ = ctx.stage_ms(9)
print("steady state per iteration: list-settled %.0f  searched %.0f  pyramid %.0f  global-path CTAs %.0f of %d" %
      ((m[0] - prev[0]) / K, (m[1] - prev[1]) / K, (m[2] - prev[2]) / K, (m[3] - prev[3]) / K, (N + 511) // 512))
print("nn %.4f ms  reduce %.4f ms  (warm, %d iterations)" % (nn / c, rd / c, c))
for extra in (300, 600):
    ctx.icp_large_iterate(extra, **never)
    ctx.synchronize()
    ctx.set_timing(True)
    prev = miss()
    ctx.icp_large_iterate(K, **never)
    ctx.synchronize()
    m = miss()
    nn, c = ctx.stage_ms(8); rd, _ = ctx.stage_ms(9)
    print("after %d more: step %.2e list-settled %.0f searched %.0f  nn %.4f ms  reduce %.4f ms" % (extra, step(), (m[0] - prev[0]) / K, (m[1] - prev[1]) / K, nn / c, rd / c))
print(ctx.icp_large_end(**never))
