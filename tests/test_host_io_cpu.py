"""CPU test of the host-side text I/O (kss-icp_b200/host/xyzIO.hpp): the reference's count-prefixed cloud format
(save_PointCloud, Main_KSS_ICP.cpp:49-59; the shipped .wlop/.gird files) written byte for byte and read back."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "kss-icp_b200", "host")

PROG = r'''
#include <cstdio>
#include <iostream>
#include "xyzIO.hpp"
int main(int argc, char** argv) {
    std::vector<std::vector<double>> a = Load_XYZ(argv[1]);
    Save_XYZ(a, argv[2]);                       // the reference's writer
    Save_XYZ(a, argv[2]);                       // append mode: a second block, like running the reference twice
    std::vector<std::vector<double>> b = Load_XYZ(argv[2]);
    std::printf("%zu %zu\n", a.size(), b.size());
    return 0;
}
'''


def test_xyz_round_trip_and_reference_format(tmp_path):
    src = tmp_path / "io.cpp"; src.write_text(PROG)
    exe = tmp_path / "io"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", HOST, str(src), "-o", str(exe)])
    pts = np.array([[0.5, -1.25, 3.0], [1e-3, 2.0, -7.125], [123456.0, 0.1, 0.333333]])
    with open(tmp_path / "in.gird", "w") as f:                      # as shipped: count line, rows, blank line
        f.write("3\n")
        for p in pts:
            f.write("%g %g %g\n" % tuple(p))
        f.write("\n")
    out = subprocess.run([str(exe), str(tmp_path / "in.gird"), str(tmp_path / "out.xyz")], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == ["3", "3"]
    text = (tmp_path / "out.xyz").read_text()
    block = "3\n" + "".join("%g %g %g\n" % tuple(p) for p in pts) + "\n"   # default ostream precision == %g
    assert text == block + block
    # plain rows without a count line (.xyz / .txt flavour)
    (tmp_path / "plain.txt").write_text("".join("%g %g %g\n" % tuple(p) for p in pts))
    out = subprocess.run([str(exe), str(tmp_path / "plain.txt"), str(tmp_path / "out2.xyz")], capture_output=True, text=True)
    assert out.stdout.split() == ["3", "3"]
    # a missing file loads as an empty cloud, like the reference (no exception)
    out = subprocess.run([str(exe), str(tmp_path / "nope.xyz"), str(tmp_path / "out3.xyz")], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == ["0", "0"]


def test_transfer_pc_generator(tmp_path):
    """the reference's test-pair generator (transferPC.hpp): TransferPC_cli writes <model>.wlop / <model>.gird in the
    count-prefixed format; the gird copy is the model rotated with the axis convention of initRegistration_Transfer,
    scaled about its centroid and translated -- and the rotation it applies to the reference's own Bunny.wlop is the one
    transfer.txt documents for Bunny.gird (the two shipped clouds then lie on top of each other)"""
    import subprocess
    import numpy as np
    import __graft_entry__ as entry
    pkg = entry.load_package()
    exe = os.path.join(ROOT, "kss-icp_b200", "host", "build", "TransferPC_cli")
    if not os.path.exists(exe):
        pytest.skip("TransferPC_cli not built")
    fx = np.load(os.path.join(ROOT, "tests", "golden", "fullsize_pairs.npz"))
    w = fx["Bunny_wlop"].astype(np.float64)
    model = tmp_path / "bunny.xyz"
    with open(model, "w") as f:
        f.write("%d\n" % len(w))
        for q in w:
            f.write("%.9g %.9g %.9g\n" % (q[0], q[1], q[2]))
        f.write("\n")
    out = subprocess.run([exe, str(model), "1", "1.1", "1.5", "0.25", "100000", "1e-9"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    def load(p):
        t = open(p).read().split()
        return np.array([float(x) for x in t[1:1 + 3 * int(t[0])]]).reshape(-1, 3)
    wl, gd = load(tmp_path / "bunny.wlop"), load(tmp_path / "bunny.gird")
    assert np.allclose(wl, w, rtol=1e-5)                                       # 6 significant digits in the text format
    exp = pkg.synth.transfer_pc(w, 1, 1.1, 1.5, 0.25)
    assert len(gd) == len(w) and np.allclose(gd, exp, rtol=2e-5, atol=2e-6)    # gridCell 1e-9: every point its own voxel
    # known answer from the reference's data: Bunny.gird is Bunny rotated about x by 1.1 (transfer.txt:9)
    from scipy.spatial import cKDTree
    g = fx["Bunny_gird"].astype(np.float64)
    before = cKDTree(g).query(w)[0].mean()
    after = cKDTree(g).query(pkg.synth.transfer_pc(w, 1, 1.1))[0].mean()
    assert after < 0.2 * before and after < 0.03
