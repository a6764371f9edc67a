"""profiling driver: AIVS simplification of the bench batch's source clouds (one launch of aivs_small_kernel)"""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import torch

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
pkg = entry.load_package()
b, _ = pkg.synth.modelnet_batch(P, n_full=2048)
ctx = pkg.Context(0)
d = torch.from_numpy(b["full_s"]).cuda()
out = torch.zeros((P, 1072, 3), dtype=torch.float64, device="cuda")
cnt = torch.zeros(P, dtype=torch.int32, device="cuda")
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.aivs_simplify_batch_device(P, d.data_ptr(), 2048, 1024, out.data_ptr(), 1072, cnt.data_ptr())
    ctx.synchronize(); t1 = time.perf_counter()
    print("aivs %d clouds: %.3f ms, mean kept %.1f" % (P, 1e3 * (t1 - t0), cnt.float().mean().item()))
ctx.aivs_status()
