"""profiling driver: one raw-cloud batch registration of P pairs (default 296 = 2 per SM)"""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import torch

P = int(sys.argv[1]) if len(sys.argv) > 1 else 296
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pkg = entry.load_package()
b, _ = pkg.synth.modelnet_batch(P, n_full=2048)
ctx = pkg.Context(0)
fs = torch.from_numpy(b["full_s"]).cuda(); ft = torch.from_numpy(b["full_t"]).cuda()
res = torch.zeros(P * pkg.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
ctx.set_timing(True)
for it in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.register_batch_device(P, (0, 0, 2048, 2048), (None, None, fs.data_ptr(), ft.data_ptr()), res.data_ptr())
    ctx.synchronize(); t1 = time.perf_counter()
    print("register %d pairs: %.3f ms" % (P, 1e3 * (t1 - t0)))
print({pkg.STAGES[i]: round(ctx.stage_ms(i)[0] / reps, 3) for i in range(len(pkg.STAGES))})
if os.environ.get("KSS_ICP_PHASES"):
    import ctypes as C
    a = np.zeros(16, np.uint64)
    ctx.lib.kss_debug_read.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_size_t, C.c_void_p]
    assert ctx.lib.kss_debug_read(ctx.h, b"icp_phase", 0, 128, a.ctypes.data_as(C.c_void_p)) == 0
    names = ["setup", "nn", "passA", "lvl2", "passB", "serial-tail+barrier", "post-loop", "fitness", "umeyama", "matmul", "convergence", "sigma-lvl2", "-", "-", "-", "-"]
    tot = float(a.sum())
    print("icp phases (thread-0 cycles, last call): " + ", ".join("%s %.1f%%" % (n, 100 * v / tot) for n, v in zip(names, a)))
    print("total Mcycles %.1f" % (tot / 1e6))
