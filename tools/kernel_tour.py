#!/usr/bin/env python
"""One warm pass over every kernel of the hot path, bracketed by cudaProfilerStart/Stop, for a single
`ncu --profile-from-start off --metrics ...` capture (profiles/rNN_kernel_table.md is made from its CSV by
tools/ncu_table.py).  Stages: 640x480 pairwise ICP pipeline (8 pairs), NDT-coarse pipeline, 1280x720 NDT with 0.05 m
voxels (BASELINE configs[2]), one large-cloud ICP on the global-memory grid path (configs[4] shape, 2 M points),
fitness / nearest, transform, concat, crop."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("tools", "realsense-pointcloud_b200"):
    sys.path.insert(0, os.path.join(ROOT, sub))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=9)
    ap.add_argument("--big", type=int, default=2_000_000)
    ap.add_argument("--iters", type=int, default=50)
    a = ap.parse_args()
    import torch
    import gen_scene
    import rspcl_b200 as R

    ctx = R.Context(0)
    W, H = 640, 480
    frames, _ = gen_scene.make_sweep(2, a.frames)
    d_frames = ctx.upload(list(frames), W, H)
    guess = np.eye(4)
    guess[:3, :3] = gen_scene.rot_y(-0.523599)
    forced = dict(max_iterations=a.iters, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300, mse_threshold_absolute=-1.0)
    icp = R.icp_params(**forced)
    si = np.arange(1, a.frames, dtype=np.int32)
    ti = np.arange(0, a.frames - 1, dtype=np.int32)
    d_out = ctx.cloud(a.frames - 1, W * H)

    W2, H2 = 1280, 720
    fr2, _ = gen_scene.make_sweep(3, 2, W2, H2, noise_scale=0.2)
    d_fr2 = ctx.upload(list(fr2), W2, H2)
    ndt005 = R.ndt_params(resolution=0.05)

    tgt = gen_scene.sample_room_surface(5, a.big)
    src = gen_scene.sample_room_surface(6, a.big)
    d_tgt, d_src = ctx.upload([tgt]), ctx.upload([src])
    big_prm = R.icp_params(max_iterations=3, max_corr_dist=0.004, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300,
                           mse_threshold_absolute=-1.0)

    def tour():
        R.register_pairs(ctx, d_frames, si, ti, R.COARSE_ICP, icp=icp, guess=guess, out_transformed=d_out)
        R.register_pairs(ctx, d_frames, si[:2], ti[:2], R.COARSE_NDT, icp=R.icp_params(), ndt=R.ndt_params(), guess=guess,
                         out_transformed=None)
        e2 = R.edge_extract(ctx, d_fr2)
        v2 = R.voxel_approx(ctx, e2)
        dl = v2.download()
        s2, t2 = ctx.upload([dl[1]]), ctx.upload([dl[0]])
        R.ndt_align(ctx, s2, t2, ndt005, guess=guess, want_aligned=True)
        R.icp_align(ctx, d_src, d_tgt, big_prm, want_aligned=False)
        R.fitness(ctx, s2, t2)
        R.nearest(ctx, s2, t2)
        R.transform(ctx, d_fr2, guess)
        R.concat(ctx, s2, t2)
        R.crop35(ctx, d_frames)

    tour()  # warm-up: pool growth, module load
    ctx.sync()
    rt = torch.cuda.cudart()
    rt.cudaProfilerStart()
    tour()
    ctx.sync()
    rt.cudaProfilerStop()
    print("tour done")
    ctx.close()


if __name__ == "__main__":
    main()
