#!/usr/bin/env python
"""Generates tests/golden/rspcl_golden.npz from the CPU oracle on seeded synthetic inputs.

The reference ships no golden vectors and cannot be built or imported here, so these fixtures pin the
ORACLE's behaviour (regression across compilers / machines) and give the GPU tests a fixed target that does not depend
on the oracle being rebuilt identically on the GPU box.  Re-run after any intentional oracle change:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for sub in ("oracle", "tools"):
    sys.path.insert(0, os.path.join(ROOT, sub))
import gen_scene  # noqa: E402
import orc  # noqa: E402

W, H = 640, 480
CW, CH, R0, C0 = 160, 96, 200, 380   # crop of frame 0 used as a small organized cloud


def crop(frame):
    return frame.reshape(H, W)[R0:R0 + CH, C0:C0 + CW].reshape(-1).copy()


def main():
    fr, T = gen_scene.make_sweep(7, 2)
    small = [crop(fr[0]), crop(fr[1])]
    out = {"seed": 7, "crop": np.array([CW, CH, R0, C0])}
    masks = []
    for k, c in enumerate(small):
        m, near = orc.canny(c, CW, CH)
        assert near == 0
        masks.append(np.packbits(m > 0))
    out["edge_mask_bits"] = np.stack(masks)
    e0, idx0 = orc.extract_edges(small[0], CW, CH)
    out["edge_idx0"] = idx0
    v0 = orc.approx_voxel(e0)
    out["voxel0"] = v0.view(np.uint32).reshape(-1, 4)
    # full-frame edge clouds -> voxel -> one ICP pair (reference settings and 10 forced iterations) and one NDT align
    ef = [orc.approx_voxel(orc.extract_edges(f, W, H)[0]) for f in fr]
    out["n_voxel_full"] = np.array([len(e) for e in ef])
    guess = np.eye(4)
    guess[:3, :3] = gen_scene.rot_y(-0.523599)
    r = orc.icp_align(ef[1], ef[0], orc.icp_params(), guess=guess, want_first_corr=True)
    out["icp_ref_T"] = r["T"]
    out["icp_ref_ncorr"] = np.array([r["n_corr"], r["iterations"], r["state"]])
    out["icp_ref_first_corr"] = r["first_corr"].astype(np.int32)
    kw = dict(max_iterations=10, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300, mse_threshold_absolute=-1.0)
    r10 = orc.icp_align(ef[1], ef[0], orc.icp_params(**kw), guess=guess)
    out["icp_10_T"] = r10["T"]
    n = orc.ndt_align(ef[1], ef[0], orc.ndt_params(), guess=guess)
    out["ndt_T"] = n["T"]
    out["ndt_meta"] = np.array([n["iterations"], n["n_derivative_evals"], n["n_hessian_evals"]])
    out["ndt_score"] = np.array([n["score"]])
    vox = orc.NdtGrid(ef[0], orc.ndt_params()).voxels()
    out["ndt_vox_ijk"] = vox["ijk"]
    out["ndt_vox_npts"] = vox["npts"]
    out["ndt_vox_mean"] = vox["mean"]
    out["gt_pair"] = gen_scene.pairwise_gt(T, 1)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rspcl_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
