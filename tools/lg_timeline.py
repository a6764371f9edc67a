import os, sys
sys.path.insert(0, "/root/repo")
import __graft_entry__ as g
pkg = g.load_package()
p = pkg.synth.scan_pair(0, 1000000)
ctx = pkg.Context(0)
never = dict(max_iter=1 << 30, fit_eps=-1.0, trans_eps=-1.0)
ctx.icp_large_begin(p["full_s"], p["full_t"])
ctx.icp_large_iterate(430, **never)
ctx.synchronize()
print(ctx.icp_large_end(max_iter=1)["iters"])
