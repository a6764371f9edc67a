"""first-stage candidate statistics in steady state only (library built with -DKSS_LG_COUNT): warm up in one process state,
then read the counters' growth over a few iterations"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
p = pkg.synth.scan_pair(0, 1000000)
ctx = pkg.Context(0)
never = dict(max_iter=1 << 30, fit_eps=-1.0, trans_eps=-1.0)
ctx.icp_large_begin(p["full_s"], p["full_t"])
ctx.icp_large_iterate(40, **never); ctx.synchronize()
print("after warm-up:"); ctx.icp_large_end(**never)
ctx.icp_large_iterate(40, **never); ctx.synchronize()
print("after 40 more:"); ctx.icp_large_end(**never)
