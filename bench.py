#!/usr/bin/env python
"""bench.py -- registrations/s on the ModelNet40-shape batch (BASELINE.json configs[2]:
2,468 pairs x 2,048 points, simplified to pNumber = 1,024, step 8 -> 729 hypotheses, ICP iter 1000)
plus the 1M-point full-resolution ICP iteration against the HBM roofline (configs[3]).

  python bench.py --gpus N --steps K --warmup W            # this framework (one rank per GPU)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

A "step" is one pass of the whole of KSSICP_init + KSSICP_Registration (pNumber rule, AIVS
simplification of both clouds, MiddleAlign, sweep, ICP runs, final apply) + PCR_QM over the batch; with N ranks the 2,468 pairs are split in contiguous blocks (strong scaling, no
data-path collective: pairs are independent).  One JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

N_PAIRS_TOTAL = 2468
N_FULL = 2048
STEP = 8.0
MAX_ITER = 1000
METRIC = "registrations/sec (ModelNet40-shape batch, 2468 pairs x 2048 pts, pNumber 1024, 729 hypotheses)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=N_PAIRS_TOTAL, help="total pairs in the batch (default: the named config)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU baseline sample (0: 16 x cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-1m", action="store_true", help="skip the 1M-point ICP-iteration roofline leg")
    ap.add_argument("--points-1m", type=int, default=1000000)
    ap.add_argument("--settle-1m", type=int, default=400, help="iterations past PCL's convergence before the steady-state iteration is timed")
    ap.add_argument("--only-1m", action="store_true", help="debug: run only the 1M-point leg")
    ap.add_argument("--config", default="all", choices=["all", "c3", "c1", "c2", "c4pipe", "c5"],
                    help="all: the headline batch (c3) plus one leg per other BASELINE.json configuration; cN: only that leg")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def c3_kernel_rooflines(stages, res, pairs, clocks):
    """what bounds the headline step: the three big stages are not HBM-bound (24-48 KB per pair lives in shared memory / L2)
    but issue-bound.  Per stage: exact-NN queries/s from the live stage time, and the warp-instruction issue rate against
    148 SMs x 4 schedulers x clock, with the instruction count per pair taken from the committed ncu capture
    (profiles/r02_c3_kernels.json, same workload)."""
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r02_c3_kernels.json")))
    except Exception:
        prof = {}
    clk = (clocks or {}).get("sm_mhz") or 1965.0
    peak = 148 * 4 * clk * 1e6
    n_s = N_FULL // 2
    q = {"sweep": 729.0 * n_s * pairs,
         "icp_hyp": float(res["total_icp_iters"].sum() - res["final_iters"].sum() + res["n_icp_runs"].sum() - len(res)) * n_s}
    out = {"issue_peak_warp_inst_per_s": peak, "sm_mhz": clk,
           "note": "stage times: single-lane untimed extra step (stage_ms_single_lane_rank0); warp instructions per pair: ncu capture"}
    for st_name, kern in (("sweep", "sweep_kernel"), ("icp_hyp", "icp_small_kernel"), ("cg_build", "cg_build (6 kernels)"), ("aivs", "aivs_small_kernel")):
        ms = stages.get(st_name, 0.0)
        if not ms:
            continue
        d = {"kernel": kern, "ms": ms}
        if st_name in q:
            d["nn_queries_per_s"] = q[st_name] / (ms * 1e-3)
        wi = prof.get(kern, {}).get("warp_inst_per_pair")
        if wi:
            d["warp_inst_per_s"] = wi * pairs / (ms * 1e-3)
            d["frac_of_issue_peak"] = d["warp_inst_per_s"] / peak
            d["warp_inst_per_pair_ncu"] = wi
        out[st_name] = d
    return out


def shard(total, world, rank):
    per = (total + world - 1) // world
    lo = min(total, rank * per)
    hi = min(total, lo + per)
    return lo, hi


def cpu_sample_pairs(args, okss):
    cores = okss.max_threads()
    n = args.cpu_sample if args.cpu_sample > 0 else 16 * cores      # ~35 registrations/s on 16 cores: 7-8 s per pass
    return n, cores


def run_reference(args, rank):
    """--impl reference: the reference's CPU path.  The reference itself cannot be built here
    (PCL 1.8.1 / FLANN / Eigen / Win32 absent), so this times the oracle port on all
    host threads, on a bounded sample of the same workload, each step the same sample."""
    if rank != 0:
        return
    from oracle import okss
    okss.build()
    pkg = entry.load_package()
    n, cores = cpu_sample_pairs(args, okss)
    b, _ = pkg.synth.modelnet_batch(n, n_full=N_FULL)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        okss.register_batch(None, None, b["full_s"], b["full_t"], step=STEP, max_iter=MAX_ITER,
                            sum_order=okss.SUM_SERIAL, method=okss.NN_KDTREE, threads=0)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
        if args.warmup and i < args.warmup and dt > 20.0:      # one warm-up is enough for a 20 s+ step
            pass
    ms = 1000.0 * float(np.mean(times))
    v = n / (ms / 1000.0)
    sample = "%d pairs of the same generator (first %d of %d), AIVS + kd-tree NN, serial sums, %d threads" % (n, n, N_PAIRS_TOTAL, cores)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "registrations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": {"workload": "configs[2] ModelNet40-shape batch (bounded sample)", "pairs_per_step": n,
                       "points": N_FULL, "pNumber": N_FULL // 2, "hypotheses": 729, "icp_max_iter": MAX_ITER,
                       "simplification": "AIVS (oracle restatement), inside the timed step"},
            "cpu_baseline": {"value": v, "unit": "registrations/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "registrations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.load_package()
    stream = torch.cuda.Stream(device=local)
    ctx = pkg.Context(local, stream=stream.cuda_stream)
    if world > 1:
        # the library's own NCCL communicator (hypothesis sharding, config 5): rank 0's id handed out through torch
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda:%d" % local)
        if rank == 0:
            idt.copy_(torch.tensor(list(pkg.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.nccl_init(bytes(idt.cpu().tolist()), rank, world)

    import bench_configs
    if args.config not in ("all", "c3"):
        from oracle import okss
        okss.build()
        def _bar():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
        if args.config == "c5":
            leg = bench_configs.leg_c5(pkg, ctx, okss, world, rank, barrier=_bar)
        elif rank == 0:
            leg = {"c1": bench_configs.leg_c1, "c2": bench_configs.leg_c2, "c4pipe": bench_configs.leg_c4pipe}[args.config](pkg, ctx, okss)
        else:
            leg = None
        if rank == 0:
            print(json.dumps({"config": args.config, "n_gpus": world, **leg}))
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    if args.only_1m:
        import bench_large
        with torch.cuda.stream(stream):
            r, ex = bench_large.icp_iteration_roofline(pkg, ctx, args, *peaks(), stream=stream)
        print(json.dumps({"roofline": r, **ex}))
        return

    lo, hi = shard(args.pairs, world, rank)
    P = hi - lo
    b, _ = pkg.synth.modelnet_batch(P, n_full=N_FULL, first=lo)
    b = {k: b[k] for k in ("full_s", "full_t")}           # raw clouds only: the library runs AIVS itself
    host = {k: torch.from_numpy(v).pin_memory() for k, v in b.items()}
    hnp = {k: v.numpy() for k, v in host.items()}
    dev = {k: v.to("cuda:%d" % local, non_blocking=False) for k, v in host.items()}
    itemsize = pkg.RESULT_DTYPE.itemsize
    d_res = torch.zeros(P * itemsize, dtype=torch.uint8, device="cuda:%d" % local)
    h_res = torch.zeros(P * itemsize, dtype=torch.uint8).pin_memory()
    res_np = h_res.numpy().view(pkg.RESULT_DTYPE)
    caps = (0, 0, hnp["full_s"].shape[1], hnp["full_t"].shape[1])
    ptrs = (None, None, dev["full_s"].data_ptr(), dev["full_t"].data_ptr())
    h2d = sum(v.numel() * 8 for v in host.values())
    d2h = P * itemsize

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        ctx.register_batch_device(P, caps, ptrs, d_res.data_ptr(), step=STEP, max_iter=MAX_ITER)

    def step_e2e():
        ctx.register_batch(None, None, hnp["full_s"], hnp["full_t"], step=STEP, max_iter=MAX_ITER, results=res_np)

    # ---- value: inputs resident in HBM, CUDA events on the launching stream, max over ranks
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms_dev = e0.elapsed_time(e1) / args.steps
    launches = ctx.launch_count() - l0
    # per-stage device times: one extra, untimed step on a single lane (in the timed steps the chunks of the batch
    # run on two internal streams and their stages overlap, so their spans would not add up to the step)
    prev_lanes = os.environ.get("KSS_LANES")
    os.environ["KSS_LANES"] = "1"
    step_device()                          # grows the single-lane buffers once, outside the stage timers
    barrier()
    ctx.set_timing(True)
    step_device()
    barrier()
    stages = {pkg.STAGES[i]: ctx.stage_ms(i)[0] for i in range(len(pkg.STAGES))}
    ctx.set_timing(False)
    if prev_lanes is None:
        del os.environ["KSS_LANES"]
    else:
        os.environ["KSS_LANES"] = prev_lanes

    # ---- e2e: host pinned buffers in, host results out, copies inside the timed region
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    ms_e2e = 1000.0 * (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda:%d" % local)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])
        tl = torch.tensor([launches], dtype=torch.int64, device="cuda:%d" % local)
        dist.all_reduce(tl)
        launches = int(tl[0])

    # ---- config 5 on every rank (hypotheses sharded over the ranks inside the library); the other configurations at N = 1
    configs = {}
    if args.config == "all":
        from oracle import okss as _ok
        if rank == 0:
            _ok.build()
        barrier()
        c5 = bench_configs.leg_c5(pkg, ctx, _ok, world, rank, barrier=barrier)
        if rank == 0:
            configs["c5"] = c5
            if world == 1:
                configs["c1"] = bench_configs.leg_c1(pkg, ctx, _ok)
                configs["c2"] = bench_configs.leg_c2(pkg, ctx, _ok)
                if not args.no_1m:
                    configs["c4pipe"] = bench_configs.leg_c4pipe(pkg, ctx, _ok, n=args.points_1m)

    # ---- N=1 only: the 1M-point full-resolution ICP iteration (roofline) and the CPU baseline
    roofline = None
    extra = {}
    cpu_baseline = None
    order_stat = None
    if rank == 0:
        hbm_peak, which = peaks()
        try:
            import bench_large
            if not args.no_1m and world == 1:
                with torch.cuda.stream(stream):
                    roofline, extra = bench_large.icp_iteration_roofline(pkg, ctx, args, hbm_peak, which, stream=stream)
        except ImportError:
            roofline = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import okss
            okss.build()
            n, cores = cpu_sample_pairs(args, okss)
            t0 = time.perf_counter()
            serial, _ = okss.register_batch(None, None, hnp["full_s"][:n], hnp["full_t"][:n], step=STEP,
                                            max_iter=MAX_ITER, sum_order=okss.SUM_SERIAL, method=okss.NN_KDTREE, threads=0)
            dt = time.perf_counter() - t0
            n1 = min(8, n)                                        # the faithful single-thread figure (the reference is serial)
            t1 = time.perf_counter()
            okss.register_batch(None, None, hnp["full_s"][:n1], hnp["full_t"][:n1], step=STEP, max_iter=MAX_ITER,
                                sum_order=okss.SUM_SERIAL, method=okss.NN_KDTREE, threads=1)
            dt1 = time.perf_counter() - t1
            cpu_baseline = {"value": n / dt, "unit": "registrations/s", "cores": cores, "kind": "port",
                            "sample": "first %d of the %d pairs, oracle port (AIVS + kd-tree NN, serial sums), %d threads, %.1f s"
                                      % (n, args.pairs, cores, dt),
                            "single_thread": {"value": n1 / dt1, "unit": "registrations/s", "cores": 1,
                                              "sample": "first %d pairs, one host thread, %.1f s" % (n1, dt1)}}
            order_stat = bench_configs.order_statistic(okss, hnp["full_s"], hnp["full_t"], serial[:min(n, 128)])
        # parity spot check of the timed output (not timed): first pair against the oracle
        line = {"metric": METRIC, "value": args.pairs / (ms_dev / 1000.0), "unit": "registrations/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+f64",
                "data": "synthetic",
                "config": {"workload": "configs[2] ModelNet40-shape batch", "pairs": args.pairs, "points": N_FULL,
                           "pNumber": N_FULL // 2, "hypotheses": 729, "icp_max_iter": MAX_ITER,
                           "simplification": "AIVS on the device, inside the timed step",
                           "parallelism": "pairs sharded in contiguous blocks, %d per rank" % ((args.pairs + world - 1) // world),
                           "l2": "inputs (%.0f MB/step/rank) and the sweep scratch exceed the 126 MB L2" % (h2d / 1e6)},
                "e2e": {"value": args.pairs / (ms_e2e / 1000.0), "unit": "registrations/s",
                        "h2d_bytes_per_step": h2d * world if world > 1 else h2d,
                        "d2h_bytes_per_step": d2h * world if world > 1 else d2h, "ms_per_step": ms_e2e},
                "gpu_launches": launches, "clocks": clocks, "stage_ms_single_lane_rank0": stages,
                "roofline": roofline, "cpu_baseline": cpu_baseline, "configs": configs}
        line.update(extra)
        line["c3_kernels"] = c3_kernel_rooflines(stages, res_np, args.pairs if world == 1 else P, clocks)
        r0 = res_np[0]
        line["result_sample"] = {"pair0_rmse": float(r0["rmse"]), "pair0_winner": int(r0["winner"]),
                                 "mean_icp_runs": float(res_np["n_icp_runs"].mean()),
                                 "mean_icp_iters": float(res_np["total_icp_iters"].mean()),
                                 "mean_minima": float(res_np["n_minima"].mean()),
                                 "multi_fraction": float(res_np["branch_multi"].mean()),
                                 "overflow_pairs": int(res_np["overflow"].sum()),
                                 "summation_order_serial_vs_canon256": order_stat}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
