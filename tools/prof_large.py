"""short driver for ncu: 1M-point ICP, 30 warm-up iterations then a few steady-state ones (see profiles/)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
p = pkg.synth.scan_pair(0, N)
ctx = pkg.Context(0)
never = dict(max_iter=1 << 30, fit_eps=-1.0, trans_eps=-1.0)
ctx.icp_large_begin(p["full_s"], p["full_t"])
ctx.icp_large_iterate(int(sys.argv[2]) if len(sys.argv) > 2 else 36, **never)
ctx.synchronize()
print(ctx.icp_large_end(**never)["fitness"])
