#!/usr/bin/env python
"""Full-size fixtures from the reference's shipped data (run in the build container only; /root/reference does
not exist on the GPU box):

    python tests/golden/make_fullsize.py

1. fullsize_pairs.npz  -- ALL 10 pairs of PS_AIS_Simplification/data/registration/<Model>.{wlop,gird} at their
   native size (wlop 4999-5000 points = source, unrotated; gird 1041-43871 points = target, rotated by the
   axis / angle of transfer.txt), float32 as the text files hold them.
2. golden_fullsize.npz -- the CPU oracle on those pairs through the RAW path of Main_KSS_ICP.cpp:79-88
   (pNumber = min(|S|, |T|) / 2 capped at 2000, AIVS of both clouds, sweep, ICP runs, final apply, PCR_QM), CANON256
   sums, plus the known rotation of transfer.txt per model.  The C1 pair of BASELINE.json (Armadillo.gird, 10 000
   points, random similarity) is generated from the Armadillo fixture by kss_icp_b200.synth.c1_pair.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
from oracle import okss  # noqa: E402

REF = "/root/reference/PS_AIS_Simplification/data/registration/"
# transfer.txt:1-10
MODELS = {"ant": ("x", 1.56), "Cat": ("y", 1.56), "Dog": ("z", 1.1), "Girl": ("x", 1.1), "hand": ("x", 1.1),
          "woodMan": ("y", 1.1), "Angel": ("z", 1.1), "Armadillo": ("x", 1.1), "Bunny": ("x", 1.1), "Horse": ("y", 1.1)}
KEYS = ("winner", "n_minima", "branch_multi", "judge_fitness", "final_fitness", "judge_iters", "final_iters",
        "total_icp_iters", "n_icp_runs", "T", "mse", "rmse", "mae", "used_angle", "align")


def load(path):
    with open(path) as f:
        n = int(f.readline())
        return np.loadtxt(f, max_rows=n).astype(np.float32)


def raw_register(s, t):
    pn = min(min(len(s), len(t)) // 2, 2000)                       # KSS_ICP.hpp:57-66
    sim_s, _ = okss.aivs_simplify(s, pn); sim_t, _ = okss.aivs_simplify(t, pn)
    return okss.register(sim_s, sim_t, s, t, step=8.0, max_iter=1000, sum_order=okss.SUM_CANON256, method=okss.NN_KDTREE)


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    pairs, gold = {}, {}
    for name, (axis, ang) in MODELS.items():
        w = load(REF + name + ".wlop"); g = load(REF + name + ".gird")
        pairs[name + "_wlop"] = w; pairs[name + "_gird"] = g
        t0 = time.time()
        r = raw_register(w.astype(np.float64), g.astype(np.float64))
        for k in KEYS:
            gold[name + "_" + k] = np.asarray(r[k])
        gold[name + "_axis_angle"] = np.array([{"x": 0, "y": 1, "z": 2}[axis], ang])
        print("%-10s src %5d tgt %5d  winner %3d of %3d  rmse %.5f  (%.1f s)" % (name, len(w), len(g), r["winner"], r["n_minima"], r["rmse"], time.time() - t0))
    # BASELINE.json configs[0]: the 10k Armadillo pair with a random similarity, raw path
    pkg = entry.load_package()
    c1 = pkg.synth.c1_pair(pairs["Armadillo_gird"])
    r = raw_register(c1["full_s"], c1["full_t"])
    for k in KEYS:
        gold["c1_" + k] = np.asarray(r[k])
    print("c1: winner %d of %d rmse %.5f" % (r["winner"], r["n_minima"], r["rmse"]))
    np.savez_compressed(os.path.join(out, "fullsize_pairs.npz"), **pairs)
    np.savez_compressed(os.path.join(out, "golden_fullsize.npz"), **gold)


if __name__ == "__main__":
    main()
