"""A/B helper: time k_icp_persist on the bench workload for several RSPCL_PERSIST_BUDGET values (CTAs per wave)."""
import os, sys
sys.path[:0] = ["tools", "realsense-pointcloud_b200"]
import numpy as np, gen_scene, rspcl_b200 as R
ctx = R.Context(0)
F = 65
frames, _ = gen_scene.make_sweep(2, F)
d = ctx.upload(list(frames), 640, 480)
g = np.eye(4); g[:3, :3] = gen_scene.rot_y(-0.523599)
icp = R.icp_params(max_iterations=50, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300, mse_threshold_absolute=-1.0)
si = np.arange(1, F, dtype=np.int32); ti = np.arange(0, F - 1, dtype=np.int32)
out = ctx.cloud(F - 1, 640 * 480)
for b in sys.argv[1:]:
    os.environ["RSPCL_PERSIST_BUDGET"] = b
    for _ in range(2):
        R.register_pairs(ctx, d, si, ti, R.COARSE_ICP, icp=icp, guess=g, out_transformed=out)
    ctx.profile_reset(); ctx.profile(True)
    ctx.timer_start()
    for _ in range(5):
        R.register_pairs(ctx, d, si, ti, R.COARSE_ICP, icp=icp, guess=g, out_transformed=out)
    ms = ctx.timer_stop() / 5
    ctx.profile(False)
    k = ctx.profile_get("k_icp_persist")
    print("budget %s: step %.3f ms, k_icp_persist %.3f ms" % (b, ms, k["ms"] / max(k["launches"], 1)))
