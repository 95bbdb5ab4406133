#!/usr/bin/env python
"""BASELINE configs[4] sanity: 4096 independent (small, random) pairs through one rspcl_icp_align call on cuda:0; every 97th
pair is checked against the oracle.  Run from the repo root on a B200 box."""
import sys, time
sys.path[:0] = ["oracle", "tools", "realsense-pointcloud_b200"]
import numpy as np, gen_scene, orc, rspcl_b200 as R
ctx = R.Context(0)
rng = np.random.default_rng(43)
F = np.float32
def rand_cloud(n, scale):
    p = np.zeros(n, R.POINT)
    p["x"], p["y"], p["z"] = (rng.uniform(-scale, scale, (3, n))).astype(F)
    p["rgba"] = 0xFF000000
    return p
S = 4096
base = [rand_cloud(260 + 7 * k, 0.15) for k in range(16)]
srcs, tgts, Ts = [], [], []
for s in range(S):
    t = base[s % 16]
    T = np.eye(4); T[:3, :3] = gen_scene.rot_axis(rng.normal(size=3), rng.uniform(-0.01, 0.01)); T[:3, 3] = rng.uniform(-0.004, 0.004, 3)
    srcs.append(orc.transform(t[::2], np.linalg.inv(T))); tgts.append(t); Ts.append(T)
kw = dict(max_iterations=10, max_corr_dist=0.03, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300, mse_threshold_absolute=-1.0)
d_s, d_t = ctx.upload(srcs), ctx.upload(tgts)
R.icp_align(ctx, d_s, d_t, R.icp_params(**kw), want_aligned=False)
t0 = time.perf_counter()
res, _, _ = R.icp_align(ctx, d_s, d_t, R.icp_params(**kw), want_aligned=False)
dt = time.perf_counter() - t0
bad = 0
for s in range(0, S, 97):
    o = orc.icp_align(srcs[s], tgts[s], orc.icp_params(**kw))
    D = np.linalg.inv(res[s]["T"].astype(np.float64)) @ o["T"].astype(np.float64)
    if np.abs(D - np.eye(4)).max() > 1e-4 or res[s]["n_corr"] != o["n_corr"]: bad += 1
print("4096 pairs: %.1f ms, %d of %d sampled pairs differ from the oracle, all iterations 10: %s" % (1e3 * dt, bad, len(range(0, S, 97)), all(r["iterations"] == 10 for r in res)))
import ctypes as C
prm = R.icp_params(**kw)
resc = (R.IcpResult * S)()
for s in range(S):
    resc[s].prev_mse = R.DBL_MAX
ctx.profile_reset(); ctx.profile(True)
t0 = time.perf_counter()
ctx.check(R.lib().rspcl_icp_align(ctx.h, d_s.h, d_t.h, C.byref(prm), None, resc, None, None))
t1 = time.perf_counter()
ctx.profile(False)
print("C call alone: %.1f ms; k_icp_persist %.2f ms" % (1e3 * (t1 - t0), ctx.profile_get("k_icp_persist")["ms"]))
