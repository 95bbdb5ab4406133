#!/usr/bin/env python
"""BASELINE configs[4], first half: a batch of N independent 640x480 frame-pair registrations through ONE rspcl_register_pairs
call on cuda:0 (the wave-splitting path of k_icp_persist: more pairs than SMs).  The batch is `blocks` copies of a 65-frame
synthetic sweep (every copy has its own device frames, edges, voxel clouds and pair states; nothing is shared between the
copies), so every block must reproduce block 0.  Usage: python tools/batch_frames.py [blocks=64] [forced]   (forced = bench.py's 50 + 50 forced iterations; default = the
reference's convergence criteria)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("tools", "realsense-pointcloud_b200"):
    sys.path.insert(0, os.path.join(ROOT, sub))


def main():
    import gen_scene
    import rspcl_b200 as R
    blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    W, H, F = 640, 480, 65
    ctx = R.Context(0)
    fr, _ = gen_scene.make_sweep(7, F)
    one = np.concatenate(list(fr))
    d_frames = ctx.cloud(blocks * F, W * H)
    host = np.tile(one, blocks)
    counts = np.full(blocks * F, W * H, np.int32)
    import ctypes as C
    d_frames.upload_raw(host.ctypes.data_as(C.c_void_p), counts, W, H, R.LAYOUT_PCD16)
    ctx.sync()
    del host
    n_pairs = blocks * (F - 1)
    d_out = ctx.cloud(n_pairs, W * H)
    src = np.concatenate([b * F + np.arange(1, F) for b in range(blocks)]).astype(np.int32)
    tgt = (src - 1).astype(np.int32)
    guess = np.eye(4)
    guess[:3, :3] = gen_scene.rot_y(-0.523599)
    forced = len(sys.argv) > 2 and sys.argv[2] == "forced"
    icp = R.icp_params(max_iterations=50, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300,
                       mse_threshold_absolute=-1.0) if forced else R.icp_params()
    d_small = ctx.cloud(F - 1, W * H)
    small = R.register_pairs(ctx, d_frames, src[:F - 1], tgt[:F - 1], R.COARSE_ICP, icp=icp, guess=guess, out_transformed=d_small)  # warm-up
    ref = [np.array(small[k].T_fine) for k in range(F - 1)]
    ctx.profile_reset()
    ctx.profile(True)
    ctx.timer_start()
    res = R.register_pairs(ctx, d_frames, src, tgt, R.COARSE_ICP, icp=icp, guess=guess, out_transformed=d_out)
    ms = ctx.timer_stop()
    ctx.profile(False)
    worst, bitwise = 0.0, 0
    for k in range(n_pairs):
        d = float(np.abs(np.array(res[k].T_fine) - ref[k % (F - 1)]).max())
        worst = max(worst, d)
        bitwise += d == 0.0
    out = {"pairs": n_pairs, "iterations": "forced 50 + 50" if forced else "reference criteria", "mean_fine_iterations": float(np.mean([r.fine_iterations for r in res])), "mean_coarse_iterations": float(np.mean([r.coarse_iterations for r in res])), "frames_resident": blocks * F, "ms": ms, "pairs_per_s": n_pairs / (ms * 1e-3),
           "k_icp_persist_ms": ctx.profile_get("k_icp_persist")["ms"], "k_icp_persist_launches": ctx.profile_get("k_icp_persist")["launches"],
           "converged": int(sum(int(r.converged) for r in res)),
           "max_abs_T_diff_vs_block0": worst, "pairs_bitwise_equal_to_block0": int(bitwise)}
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
