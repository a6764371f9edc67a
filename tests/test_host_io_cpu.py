"""CPU test of the host-side text I/O (kss-icp_b200/host/xyzIO.hpp): the reference's count-prefixed cloud format
(save_PointCloud, Main_KSS_ICP.cpp:49-59; the shipped .wlop/.gird files) written byte for byte and read back."""
import os
import subprocess

import numpy as np

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
