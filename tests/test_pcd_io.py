"""the facade's .pcd reader/writer (rspcl::io::loadPCDFile / savePCDFileBinary, main.cpp:79-87) against the
numpy reader/writer of tools/gen_scene.py -- binary files as the synthetic datasets use them and an ASCII file with the
header shape of the reference's examples/visualizer/example.pcd (`.PCD v.7`, rgb written as a float VALUE).  CPU only."""
import os
import subprocess

import numpy as np

import gen_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def checksum(pts):
    w = np.stack([pts["x"].view(np.uint32), pts["y"].view(np.uint32), pts["z"].view(np.uint32), pts["rgba"]], 1).reshape(-1)
    s = 0
    for v in w.tolist():
        s = (s * 1000003 + v) % (1 << 64)
    return s


def build(tmp_path):
    exe = tmp_path / "pcd_roundtrip"
    lib = os.path.join(ROOT, "realsense-pointcloud_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++14", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "pcd_roundtrip.cpp"),
                           "-L" + lib, "-lrspcl_b200", "-Wl,-rpath," + lib])
    return str(exe)


def test_binary_organized_roundtrip(tmp_path):
    exe = build(tmp_path)
    full, _ = gen_scene.make_sweep(9, 1)
    frame = np.ascontiguousarray(full[0].reshape(480, 640)[100:148, 200:264]).reshape(-1)  # a 64 x 48 window
    fr = [frame]
    src, dst = str(tmp_path / "a.pcd"), str(tmp_path / "b")  # main.cpp:87 writes the merged cloud without a suffix
    gen_scene.write_pcd(src, fr[0], 64, 48)
    out = subprocess.run([exe, src, dst], capture_output=True, text=True, check=True).stdout.split()
    assert [int(out[0]), int(out[1]), int(out[2])] == [64 * 48, 64, 48]
    assert int(out[3]) == checksum(fr[0])
    back, w, h = gen_scene.read_pcd(dst)
    assert (w, h) == (64, 48) and np.array_equal(back, fr[0])


def test_ascii_file_like_the_reference_example(tmp_path):
    exe = build(tmp_path)
    rng = np.random.default_rng(4)
    n = 213
    xyz = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    rgba = rng.integers(0, 2 ** 24, n).astype(np.uint32)  # alpha 0: such values are normal floats and survive %g printing
    lines = ["# .PCD v.7 - Point Cloud Data file format", "VERSION .7", "FIELDS x y z rgb", "SIZE 4 4 4 4", "TYPE F F F F",
             "COUNT 1 1 1 1", "WIDTH %d" % n, "HEIGHT 1", "VIEWPOINT 0 0 0 1 0 0 0", "POINTS %d" % n, "DATA ascii"]
    for p, c in zip(xyz, rgba):
        lines.append("%.9g %.9g %.9g %.9g" % (p[0], p[1], p[2], np.array([c], np.uint32).view(np.float32)[0]))
    src, dst = str(tmp_path / "ex.pcd"), str(tmp_path / "ex_bin.pcd")
    open(src, "w").write("\n".join(lines) + "\n")
    out = subprocess.run([exe, src, dst], capture_output=True, text=True, check=True).stdout.split()
    assert [int(out[0]), int(out[1]), int(out[2])] == [n, n, 1]
    back, w, h = gen_scene.read_pcd(dst)
    assert np.array_equal(back["x"], xyz[:, 0]) and np.array_equal(back["y"], xyz[:, 1]) and np.array_equal(back["z"], xyz[:, 2])
    assert np.array_equal(back["rgba"], rgba)


def test_missing_file_raises(tmp_path):
    exe = build(tmp_path)
    r = subprocess.run([exe, str(tmp_path / "nope.pcd"), str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr


def test_ascii_file_with_unsigned_rgb_like_exampleTemp(tmp_path):
    """examples/visualizer/exampleTemp.pcd declares `TYPE F F F U` and writes rgb as the packed unsigned integer
    (e.g. 4281353262 = 0xFF30A42E); PCL stores those bits directly.  A reader that parses the token as a float value
    yields wrong colours."""
    exe = build(tmp_path)
    rng = np.random.default_rng(6)
    n = 10
    xyz = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    rgba = rng.integers(0, 2 ** 24, n).astype(np.uint32) | np.uint32(0xFF000000)   # alpha 255: NaN patterns as floats
    rgba[0] = 4281353262
    lines = ["VERSION .7", "FIELDS x y z rgb", "SIZE 4 4 4 4", "TYPE F F F U", "COUNT 1 1 1 1", "WIDTH %d" % n, "HEIGHT 1",
             "VIEWPOINT 0 0 0 1 0 0 0", "POINTS %d" % n, "DATA ascii"]
    for p, c in zip(xyz, rgba):
        lines.append("%.9g %.9g %.9g %d" % (p[0], p[1], p[2], int(c)))
    src, dst = str(tmp_path / "u.pcd"), str(tmp_path / "u_bin.pcd")
    open(src, "w").write("\n".join(lines) + "\n")
    subprocess.run([exe, src, dst], capture_output=True, text=True, check=True)
    back, w, h = gen_scene.read_pcd(dst)
    assert np.array_equal(back["rgba"], rgba)
    assert np.array_equal(back["x"], xyz[:, 0])


def test_unsupported_headers_are_rejected(tmp_path):
    exe = build(tmp_path)
    hdr = ["VERSION .7", "FIELDS x y z rgb", "SIZE 8 8 8 4", "TYPE F F F U", "COUNT 1 1 1 1", "WIDTH 1", "HEIGHT 1",
           "POINTS 1", "DATA ascii", "0 0 0 1"]
    src = str(tmp_path / "bad.pcd")
    open(src, "w").write("\n".join(hdr) + "\n")
    r = subprocess.run([exe, src, str(tmp_path / "o.pcd")], capture_output=True, text=True)
    assert r.returncode != 0 and "SIZE" in r.stderr
