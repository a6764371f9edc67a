"""a batch of configs[0]-sized pairs (10k-point full clouds): registrations/s through the host-buffer ABI"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prs = [pkg.synth.modelnet_pair(1000 + i, n_full=10000) for i in range(P)]
fs = np.stack([p["full_s"] for p in prs]); ft = np.stack([p["full_t"] for p in prs])
ctx = pkg.Context(0)
ctx.register_batch(None, None, fs, ft)
ctx.set_timing(True)
t0 = time.perf_counter()
r = ctx.register_batch(None, None, fs, ft)
ms = 1000 * (time.perf_counter() - t0)
print("%d pairs of 10k points: %.1f ms = %.0f registrations/s; stage ms: aivs %.1f metrics %.1f icp %.1f sweep %.1f cg %.1f" % (
    P, ms, P / ms * 1e3, ctx.stage_ms(11)[0], ctx.stage_ms(6)[0], ctx.stage_ms(4)[0], ctx.stage_ms(1)[0], ctx.stage_ms(10)[0]))
print("rmse[0:3]", [float(x["rmse"]) for x in r[:3]])
