"""stage split of the reference-faithful configs[3] pipeline (1M x 1M pair -> AIVS to 2000 -> hot path -> PCR_QM at 1M)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
p = pkg.synth.scan_pair(0, N)
ctx = pkg.Context(0)
fs, ft = p["full_s"][None], p["full_t"][None]
ctx.register_batch(None, None, fs, ft)
ctx.set_timing(True)
t0 = time.perf_counter()
ctx.register_batch(None, None, fs, ft)
ms = 1000 * (time.perf_counter() - t0)
names = ["prep", "sweep", "sweep_finalize", "icp_judge", "icp_hyp", "select_apply", "metrics", "large_build", "large_nn", "large_reduce", "cg_build", "aivs"]
print("host clock %.1f ms" % ms)
for i, nm in enumerate(names):
    v, c = ctx.stage_ms(i)
    if c:
        print("  %-14s %8.3f ms (%d spans)" % (nm, v, c))
