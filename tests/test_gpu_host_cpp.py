"""The reference's UNMODIFIED Main_KSS_ICP.cpp (compiled by kss-icp_b200/host/Makefile against this repo's
same-named KSS_ICP.hpp / initRegistrationKSS.hpp / registrationMeasure.hpp) and the batch driver, run as
real executables on the GPU and checked against the Python binding of the same C ABI."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "kss-icp_b200", "host", "build")


def _write_ply(path, pts):
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                "element face 0\nproperty list uchar int vertex_indices\nend_header\n" % len(pts))
        for p in pts:
            f.write("%.9g %.9g %.9g\n" % (np.float32(p[0]), np.float32(p[1]), np.float32(p[2])))


def _expected(ctx, okss, s, t):
    """KSSICP_init + KSSICP_Registration + PCR_QM with the oracle's AIVS in front of the C ABI's registration"""
    pn = min(min(len(s), len(t)) // 2, 2000)
    return ctx.register(okss.aivs_simplify(s, pn)[0], okss.aivs_simplify(t, pn)[0], s, t)


def test_unmodified_reference_main(ctx, okss, pkg, tmp_path):
    exe = os.path.join(BIN, "Main_KSS_ICP")
    if not os.path.exists(exe):
        pytest.skip("Main_KSS_ICP was not built (reference tree absent at build time)")
    p = pkg.synth.modelnet_pair(61, n_full=1500)
    d = tmp_path / "E:" / "chen_database" / "_Registration" / "_MiddleResult"     # Main_KSS_ICP.cpp:67-71
    d.mkdir(parents=True)
    _write_ply(d / "centuarPart.ply", p["full_s"]); _write_ply(d / "centuar.ply", p["full_t"])
    out = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"Registration Measure:MSE: (\S+) RMSE: (\S+) MAE: (\S+)", out.stdout)
    assert m, out.stdout[-2000:]
    exp = _expected(ctx, okss, p["full_s"], p["full_t"])
    got = [float(x) for x in m.groups()]
    assert np.allclose(got, [exp["mse"], exp["rmse"], exp["mae"]], rtol=2e-5)       # 6 significant digits printed
    xyz = (d / "Registration.xyz").read_text().split()
    assert int(xyz[0]) == 1500 and len(xyz) == 1 + 3 * 1500                        # save_PointCloud (Main_KSS_ICP.cpp:49-59)


def test_batch_driver(ctx, okss, pkg, tmp_path):
    exe = os.path.join(BIN, "Main_KSS_List_b200")
    if not os.path.exists(exe):
        pytest.skip("batch driver not built")
    lines, exp = [], []
    for i, n in enumerate((700, 1100, 400)):
        p = pkg.synth.modelnet_pair(70 + i, n_full=n)
        a, b = tmp_path / ("s%d.ply" % i), tmp_path / ("t%d.ply" % i)
        _write_ply(a, p["full_s"]); _write_ply(b, p["full_t"])
        lines.append("%s %s" % (a, b))
        exp.append(_expected(ctx, okss, p["full_s"], p["full_t"]))
    (tmp_path / "list.txt").write_text("\n".join(lines) + "\n")
    out = subprocess.run([exe, str(tmp_path / "list.txt")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    rows = re.findall(r"pair (\d+) MSE: (\S+) RMSE: (\S+) MAE: (\S+) fitness: (\S+) hypotheses: (\d+) winner: (-?\d+)", out.stdout)
    assert len(rows) == 3
    for r, e in zip(rows, exp):
        assert np.allclose([float(r[1]), float(r[2]), float(r[3])], [e["mse"], e["rmse"], e["mae"]], rtol=2e-5)
        assert int(r[5]) == int(e["n_minima"]) and int(r[6]) == int(e["winner"])


def test_batch_driver_xyz_formats_and_aligned_output(ctx, okss, pkg, tmp_path):
    """the count-prefixed text format of the reference (.xyz as written by save_PointCloud, .wlop / .gird as shipped
    under data/registration): read by the batch driver, aligned clouds written back in the same format"""
    exe = os.path.join(BIN, "Main_KSS_List_b200")
    if not os.path.exists(exe):
        pytest.skip("batch driver not built")
    p = pkg.synth.modelnet_pair(81, n_full=900)
    def write(path, pts, count_line=True):
        with open(path, "w") as f:
            if count_line:
                f.write("%d\n" % len(pts))
            for q in pts:
                f.write("%.9g %.9g %.9g\n" % (np.float32(q[0]), np.float32(q[1]), np.float32(q[2])))
            f.write("\n")
    write(tmp_path / "a.wlop", p["full_s"]); write(tmp_path / "a.gird", p["full_t"])
    write(tmp_path / "b.xyz", p["full_s"], count_line=False); write(tmp_path / "b.txt", p["full_t"], count_line=False)
    (tmp_path / "list.txt").write_text("%s %s\n%s %s\n" % (tmp_path / "a.wlop", tmp_path / "a.gird", tmp_path / "b.xyz", tmp_path / "b.txt"))
    out_dir = tmp_path / "out"; out_dir.mkdir()
    out = subprocess.run([exe, str(tmp_path / "list.txt"), "8", "1000", str(out_dir)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    rows = re.findall(r"pair (\d+) MSE: (\S+) RMSE: (\S+) MAE: (\S+)", out.stdout)
    assert len(rows) == 2 and rows[0][1:] == rows[1][1:]                 # same clouds through both flavours
    exp, pa = ctx.register_batch(None, None, p["full_s"][None], p["full_t"][None], want_points=True)
    assert np.allclose([float(x) for x in rows[0][1:]], [exp[0]["mse"], exp[0]["rmse"], exp[0]["mae"]], rtol=2e-5)
    txt = (out_dir / "0Align.xyz").read_text().split()
    assert int(txt[0]) == 900 and len(txt) == 1 + 3 * 900
    got = np.array([float(x) for x in txt[1:]]).reshape(900, 3)
    assert np.allclose(got, pa[0], rtol=1e-5, atol=1e-6)                 # default ostream precision: 6 significant digits


def test_released_command_line(ctx, okss, pkg, tmp_path):
    """EXE/Readme.txt: `KSS-ICP.exe PointSource.ply PointTarget.ply`, result written as .xyz"""
    exe = os.path.join(BIN, "KSS_ICP_cli")
    if not os.path.exists(exe):
        pytest.skip("KSS_ICP_cli not built")
    p = pkg.synth.modelnet_pair(91, n_full=800)
    _write_ply(tmp_path / "s.ply", p["full_s"]); _write_ply(tmp_path / "t.ply", p["full_t"])
    out = subprocess.run([exe, str(tmp_path / "s.ply"), str(tmp_path / "t.ply"), str(tmp_path / "r.xyz")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"Registration Measure:MSE: (\S+) RMSE: (\S+) MAE: (\S+)", out.stdout)
    assert m, out.stdout[-2000:]
    exp = _expected(ctx, okss, p["full_s"], p["full_t"])
    assert np.allclose([float(x) for x in m.groups()], [exp["mse"], exp["rmse"], exp["mae"]], rtol=2e-5)
    xyz = (tmp_path / "r.xyz").read_text().split()
    assert int(xyz[0]) == 800 and len(xyz) == 1 + 3 * 800


@pytest.mark.parametrize("mode", ["pairs", "hyp"])
def test_batch_driver_multi_gpu(ctx, okss, pkg, tmp_path, mode):
    """Main_KSS_List_b200 --gpus 2: the C++ host drives the loop of Main_KSS_List.cpp:128-179 on several GPUs, pairs in
    contiguous blocks (kss_register_batch_multi) or hypotheses sharded (NCCL inside the library): same output as one GPU"""
    import torch
    exe = os.path.join(BIN, "Main_KSS_List_b200")
    if not os.path.exists(exe) or torch.cuda.device_count() < 2:
        pytest.skip("needs the batch driver and 2 GPUs")
    lines = []
    for i, n in enumerate((700, 1100, 400, 900, 650)):
        p = pkg.synth.modelnet_pair(170 + i, n_full=n)
        a, b = tmp_path / ("s%d.ply" % i), tmp_path / ("t%d.ply" % i)
        _write_ply(a, p["full_s"]); _write_ply(b, p["full_t"])
        lines.append("%s %s" % (a, b))
    (tmp_path / "list.txt").write_text("\n".join(lines) + "\n")
    one = subprocess.run([exe, str(tmp_path / "list.txt")], capture_output=True, text=True, timeout=300)
    two = subprocess.run([exe, "--gpus", "2"] + (["--hyp-shard"] if mode == "hyp" else []) + [str(tmp_path / "list.txt")],
                         capture_output=True, text=True, timeout=300)
    assert one.returncode == 0 and two.returncode == 0, (one.stderr[-1000:], two.stderr[-1000:])
    pat = r"pair \d+ MSE: .*"
    assert re.findall(pat, one.stdout) == re.findall(pat, two.stdout) and len(re.findall(pat, one.stdout)) == 5
    assert "2 GPUs" in two.stdout
