"""IMU initial-guess front end of the C++ facade: the complementary filter of rotation_estimator.hpp:22-79
restated in numpy float32 here, against the facade's RotationEstimator replaying the same synthetic trace (CPU only)."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = np.float32


def filter_ref(trace, frame_ts, alpha=F(0.98)):
    theta = np.zeros(3, F)
    first, last_ts, out, k = True, 0.0, [], 0
    for t in frame_ts:
        while k < len(trace) and trace[k][0] <= t:
            ts, kind, v = trace[k][0], trace[k][1], np.asarray(trace[k][2:], F)
            k += 1
            if kind == 0:
                if first:
                    last_ts = ts
                    continue
                dt = F((ts - last_ts) / 1000.0)
                last_ts = ts
                theta = theta + np.array([-(v[2] * dt), -(v[1] * dt), v[0] * dt], F)
            else:
                az = F(np.arctan2(np.float64(v[1]), np.float64(v[2])))
                ax = F(np.arctan2(np.float64(v[0]), np.sqrt(np.float64(F(v[1] * v[1] + v[2] * v[2])))))
                if first:
                    first = False
                    theta = np.array([ax, F(np.pi), az], F)
                else:
                    theta = np.array([theta[0] * alpha + ax * (F(1) - alpha), theta[1], theta[2] * alpha + az * (F(1) - alpha)], F)
        out.append(theta.copy())
    return np.array(out)


def rot(axis, a):
    c, s = np.cos(a), np.sin(a)
    i, j = [(1, 2), (0, 2), (0, 1)][axis]
    R = np.eye(4)
    R[i, i], R[j, j] = c, c
    if axis == 1:
        R[i, j], R[j, i] = s, -s
    else:
        R[i, j], R[j, i] = -s, s
    return R


def test_rotation_estimator_replay(tmp_path):
    exe = tmp_path / "imu_trace"
    lib = os.path.join(ROOT, "realsense-pointcloud_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++14", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "imu_trace.cpp"),
                           "-L" + lib, "-lrspcl_b200", "-Wl,-rpath," + lib])
    rng = np.random.default_rng(7)
    trace, t = [], 0.0
    for i in range(600):  # 200 Hz gyro, 62.5 Hz accel, a slow yaw of ~ -30 deg per second plus noise (D435i-like rates)
        t += 5.0
        trace.append((t, 0, *(rng.normal(0, 0.02, 3) + np.array([0.01, 0.5236, -0.02]))))
        if i % 3 == 0:
            trace.append((t + 0.1, 1, *(rng.normal(0, 0.05, 3) + np.array([0.3, -9.7, 0.8]))))
    frame_ts = [20.0, 1000.0, 2000.0, 3000.0]
    text = "%d\n" % len(trace) + "".join("%.3f %d %.9g %.9g %.9g\n" % (a, b, F(c), F(d), F(e)) for a, b, c, d, e in trace)
    text += "%d\n" % len(frame_ts) + " ".join("%.3f" % v for v in frame_ts) + "\n"
    out = subprocess.run([str(exe)], input=text, capture_output=True, text=True, check=True).stdout.split("\n")
    got = np.array([[float(x) for x in ln.split()] for ln in out[:4]])
    ref = filter_ref([(a, b, F(c), F(d), F(e)) for a, b, c, d, e in trace], frame_ts)
    assert np.allclose(got, ref, rtol=0, atol=2e-6), (got, ref)
    assert abs(got[0, 1] - np.pi) < 0.02  # yaw convention: PI at the first accelerometer sample (+ 15 ms of gyro)
    assert 1.3 < got[0, 1] - got[-1, 1] < 1.8  # ~0.52 rad/s of yaw integrated over ~3 s (sign: theta.y -= gyro.y dt)
    # ICP scheme guess (icp:86-92): R_z(theta.x) R_y(-theta.y) R_x(theta.z) of the angles relative to frame 0
    rel = got[-1] - got[0]
    G = rot(2, rel[0]) @ rot(1, -rel[1]) @ rot(0, rel[2])
    gm = np.array([float(x) for x in out[4].split()]).reshape(4, 4)
    assert np.allclose(gm, G, atol=1e-5)
