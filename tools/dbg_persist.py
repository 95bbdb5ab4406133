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
R.register_pairs(ctx, d, si, ti, R.COARSE_ICP, icp=icp, guess=g, out_transformed=out)
os.environ["RSPCL_PERSIST_DBG"] = "1"
R.register_pairs(ctx, d, si, ti, R.COARSE_ICP, icp=icp, guess=g, out_transformed=out)
