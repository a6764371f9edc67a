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
prev = miss()
for it in range(12):
    ctx.icp_large_iterate(1, **never)
    ctx.synchronize()
    m = miss()
    print("iter %2d: not verified %8d  stage-2 %8d  pyramid %8d  global-path CTAs %6d" % (it, m[0] - prev[0], m[1] - prev[1], m[2] - prev[2], m[3] - prev[3]))
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
print("steady state per iteration: not verified %.0f  stage-2 %.0f  pyramid %.0f  global-path CTAs %.0f of %d" %
      ((m[0] - prev[0]) / K, (m[1] - prev[1]) / K, (m[2] - prev[2]) / K, (m[3] - prev[3]) / K, (N + 511) // 512))
print("nn %.4f ms  reduce %.4f ms  (warm, %d iterations)" % (nn / c, rd / c, c))
print(ctx.icp_large_end(**never))
