"""Golden nearest-neighbour vectors from REAL FLANN code: OpenCV's bundled copy of FLANN (cv2.flann_Index,
algorithm 4 = FLANN_INDEX_KDTREE_SINGLE, the exact single kd-tree that pcl::KdTreeFLANN instantiates as
flann::KDTreeSingleIndex; L2 distance accumulated x, y, z in fp32 like flann::L2_Simple), exact search
(checks = -1, eps = 0).  Run in the build container:  python tests/golden/make_flann_golden.py
Writes tests/golden/flann_nn.npz (queries, targets, FLANN's index and squared distance per query).

What these vectors pin: the squared-distance arithmetic of the correspondence search bit for bit, and the index for
every query whose minimum is attained once.  Among fp32-equal minima FLANN returns whichever its traversal meets
first (cases "dups" and "lattice" hold such ties on purpose); the oracle and the GPU return the lowest index --
the documented exception of BASELINE.json's north_star."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import __graft_entry__ as entry  # noqa: E402


def flann_nn(q, t):
    idx = cv2.flann_Index(np.ascontiguousarray(t, np.float32), dict(algorithm=4, leaf_max_size=10, reorder=True, dim=3))
    i, d = idx.knnSearch(np.ascontiguousarray(q, np.float32), 1, params=dict(checks=-1, eps=0.0, sorted=True))
    return i.ravel().astype(np.int32), d.ravel().astype(np.float32)


def cases():
    synth = entry.load_package().synth
    rng = np.random.default_rng(20261019)
    t = rng.normal(size=(5000, 3)).astype(np.float32); q = rng.normal(size=(3000, 3)).astype(np.float32)
    yield "random", q, t
    t = rng.normal(size=(6000, 3)).astype(np.float32); t[3000:3020] = t[5:25]
    q = rng.normal(size=(3000, 3)).astype(np.float32); q[:20] = t[5:25]
    yield "dups", q, t
    g = np.stack(np.meshgrid(np.arange(16.0), np.arange(16.0), np.arange(16.0)), -1).reshape(-1, 3) / 16.0
    t = g[rng.permutation(len(g))].astype(np.float32); q = (g[::2] + 1.0 / 32.0).astype(np.float32)
    yield "lattice", q, t
    p = synth.scan_pair(1, 6000)
    yield "scan", p["full_s"].astype(np.float32), p["full_t"].astype(np.float32)
    p = synth.modelnet_pair(11, n_full=4000)
    yield "cad", p["full_s"].astype(np.float32), p["full_t"].astype(np.float32)


if __name__ == "__main__":
    out = {"opencv_version": np.array(cv2.__version__)}
    for name, q, t in cases():
        i, d = flann_nn(q, t)
        out[name + "_q"] = q; out[name + "_t"] = t; out[name + "_idx"] = i; out[name + "_d2"] = d
        print(name, len(q), len(t))
    np.savez_compressed(os.path.join(HERE, "flann_nn.npz"), **out)
