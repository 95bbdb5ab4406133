#!/usr/bin/env python
"""BASELINE configs[2]: edge-based NDT on one 1280x720 pair (921,600 points per frame), 0.05 m voxels, step 0.1, eps 0.01.
Prints ms per NDT iteration and per derivative evaluation (CUDA events on the library stream, per-kernel breakdown)."""
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
    ctx = R.Context(0)
    W2, H2 = 1280, 720
    fr, Tgt = gen_scene.make_sweep(3, 2, W2, H2, noise_scale=0.2)
    d = ctx.upload(list(fr), W2, H2)
    guess = np.eye(4)
    guess[:3, :3] = gen_scene.rot_y(-0.523599)
    prm = R.ndt_params(resolution=0.05)
    e = R.edge_extract(ctx, d)
    v = R.voxel_approx(ctx, e).download()
    src, tgt = ctx.upload([v[1]]), ctx.upload([v[0]])
    out = {}
    for label, p in (("res_0.05", prm), ("res_1.0_reference_literal", R.ndt_params())):
        R.ndt_align(ctx, src, tgt, p, guess=guess, want_aligned=False)  # warm-up
        ctx.profile_reset()
        ctx.profile(True)
        ctx.timer_start()
        res, _ = R.ndt_align(ctx, src, tgt, p, guess=guess, want_aligned=False)
        ms = ctx.timer_stop()
        ctx.profile(False)
        r = res[0]
        k = {n: ctx.profile_get(n) for n in ("k_ndt_eval", "k_ndt_control", "ndt_voxel_build")}
        out[label] = {"points_src": int(len(v[1])), "points_tgt": int(len(v[0])), "iterations": r["iterations"],
                      "derivative_evals": r["n_derivative_evals"], "converged": r["converged"], "ms_total": ms,
                      "ms_per_iteration": ms / max(r["iterations"], 1),
                      "ms_per_derivative_eval": k["k_ndt_eval"]["ms"] / max(k["k_ndt_eval"]["launches"], 1),
                      "ms_per_control_step": k["k_ndt_control"]["ms"] / max(k["k_ndt_control"]["launches"], 1),
                      "ms_voxel_build": k["ndt_voxel_build"]["ms"]}
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
