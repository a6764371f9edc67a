#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/.  Run in the build container only
(it reads the reference's shipped data under /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

1. fixtures_pairs.npz  -- three of the reference's own registration pairs
   (PS_AIS_Simplification/data/registration/<Model>.{wlop,gird}, known rotation in transfer.txt),
   decimated by a fixed stride so the file stays small, float32.
2. golden_oracle.npz   -- what the CPU oracle (oracle/kss_oracle.cpp) computes on those inputs:
   MiddleAlign, the 9x9x9 sweep grid, minima, judge / final ICP, metrics, the AIVS index lists and the
   registration of the AIVS-simplified clouds.  The oracle is the only
   CPU restatement available (the reference needs PCL 1.8.1 / FLANN / Eigen and cannot run here),
   so these vectors pin the ORACLE against regressions and pin the GPU path against the oracle;
   the link to the reference itself is the known-answer rotation of transfer.txt checked in
   tests/test_oracle_cpu.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import okss  # noqa: E402

REF = "/root/reference/PS_AIS_Simplification/data/registration/"
MODELS = {"Bunny": ("x", 1.1), "Horse": ("y", 1.15), "Dog": ("z", 1.1)}     # transfer.txt


def load(path):
    with open(path) as f:
        n = int(f.readline())
        return np.loadtxt(f, max_rows=n).astype(np.float32)


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    pairs, gold = {}, {}
    for name, (axis, ang) in MODELS.items():
        w = load(REF + name + ".wlop"); g = load(REF + name + ".gird")
        src = w[::4][:1200]                          # source: WLOP resample (unrotated)
        tgt = g[:: max(1, len(g) // 1000)][:1200]    # target: grid resample, rotated by transfer.txt
        pairs[name + "_src"] = src; pairs[name + "_tgt"] = tgt
        s = src.astype(np.float64); t = tgt.astype(np.float64)
        a7, al = okss.middle_align(s, t)
        sw = okss.sweep(al, t, 8.0, okss.SCORE_AVE, okss.NN_KDTREE)
        reg = okss.register(s, t, s, t, step=8.0, max_iter=1000, sum_order=okss.SUM_CANON256, method=okss.NN_KDTREE)
        ser = okss.register(s, t, s, t, step=8.0, max_iter=1000, sum_order=okss.SUM_SERIAL, method=okss.NN_KDTREE)
        gold[name + "_align7"] = a7
        gold[name + "_value"] = sw["value"]
        gold[name + "_best_index"] = sw["best_index"]
        gold[name + "_minima"] = sw["minima"]
        for k in ("winner", "n_minima", "branch_multi", "judge_fitness", "final_fitness", "judge_iters", "final_iters",
                  "total_icp_iters", "T", "mse", "rmse", "mae", "used_angle"):
            gold[name + "_canon_" + k] = np.asarray(reg[k])
            gold[name + "_serial_" + k] = np.asarray(ser[k])
        gold[name + "_axis_angle"] = np.array([{"x": 0, "y": 1, "z": 2}[axis], ang])
        # AIVS simplification of both clouds with the pNumber rule, and the registration behind it
        pn = min(len(s), len(t)) // 2
        sim_s, idx_s = okss.aivs_simplify(s, pn); sim_t, idx_t = okss.aivs_simplify(t, pn)
        raw = okss.register(sim_s, sim_t, s, t, step=8.0, max_iter=1000, sum_order=okss.SUM_CANON256, method=okss.NN_KDTREE)
        gold[name + "_aivs_src_idx"] = idx_s; gold[name + "_aivs_tgt_idx"] = idx_t
        for k in ("winner", "n_minima", "final_fitness", "T", "mse", "rmse", "mae"):
            gold[name + "_raw_" + k] = np.asarray(raw[k])
    np.savez_compressed(os.path.join(out, "fixtures_pairs.npz"), **pairs)
    np.savez_compressed(os.path.join(out, "golden_oracle.npz"), **gold)
    print("wrote", {k: v.shape for k, v in pairs.items()})


if __name__ == "__main__":
    main()
