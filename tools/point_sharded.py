#!/usr/bin/env python
"""Point-sharded ICP / NDT over N GPUs of one node (BASELINE configs[4]: one huge cloud, per-GPU partial sums combined by
an NCCL all-reduce over NVLink).  Launch:  python -m torch.distributed.run --nproc-per-node N tools/point_sharded.py ...

The target (replicated on every GPU) and the source (sharded) are independent uniform samples of the same room surface;
the source is moved by a known rigid transform the registration has to recover.  Prints one JSON line on rank 0 with the
time per iteration (CUDA events on the library stream, max over ranks), the HBM-roofline fraction of the correspondence
kernel (32 B per source point per iteration) and the agreement between ranks / with a single-GPU run."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("tools", "realsense-pointcloud_b200"):
    sys.path.insert(0, os.path.join(ROOT, sub))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=4_000_000, help="points in the source AND in the target cloud")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--gate", type=float, default=0.004, help="max correspondence distance [m]")
    ap.add_argument("--method", default="icp", choices=["icp", "ndt"])
    ap.add_argument("--check-single", action="store_true", help="rank 0 also aligns the un-sharded source for comparison")
    ap.add_argument("--seed", type=int, default=5)
    a = ap.parse_args()

    import torch
    import torch.distributed as dist
    import gen_scene
    import rspcl_b200 as R

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = R.Context(local)
    if world > 1:
        uid = torch.from_numpy(R.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
        dist.broadcast(uid, 0)
        R.comm_init(ctx, world, rank, uid.cpu().numpy())

    tgt = gen_scene.sample_room_surface(a.seed, a.points)
    src0 = gen_scene.sample_room_surface(a.seed + 1, a.points)
    T = np.eye(4)
    T[:3, :3] = gen_scene.rot_axis([0.3, 1.0, 0.2], 0.0004)
    T[:3, 3] = [0.0006, -0.0004, 0.0005]
    Ti = np.linalg.inv(T)
    xyz = np.stack([src0["x"], src0["y"], src0["z"]], 1).astype(np.float64) @ Ti[:3, :3].T + Ti[:3, 3]
    src = src0.copy()
    src["x"], src["y"], src["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    lo, hi = rank * a.points // world, (rank + 1) * a.points // world
    d_tgt = ctx.upload([tgt])
    d_src = ctx.upload([src[lo:hi]])

    if a.method == "icp":
        prm = R.icp_params(max_iterations=a.iters, max_corr_dist=a.gate, transformation_epsilon=-1.0,
                           euclidean_fitness_epsilon=-1e300, mse_threshold_absolute=-1.0)
        run = lambda c, s, t: R.icp_align_sharded(c, s, t, prm)[0][0]
    else:
        prm = R.ndt_params(resolution=0.25, max_iterations=a.iters)
        run = lambda c, s, t: R.ndt_align_sharded(c, s, t, prm)[0][0]

    run(ctx, d_src, d_tgt)  # warm-up (pool growth, NCCL channels)
    if world > 1:
        dist.barrier()
    ctx.profile_reset()
    ctx.profile(True)
    ctx.timer_start()
    res = run(ctx, d_src, d_tgt)
    ms = ctx.timer_stop()
    ctx.profile(False)
    if a.method == "icp":
        ks, kr = ctx.profile_get("k_icp_stream"), ctx.profile_get("k_icp_rescan")
        kern = ks if ks["launches"] else ctx.profile_get("k_icp_step")
        rescan_ms = kr["ms"]
    else:
        kern = ctx.profile_get("k_ndt_eval")
        rescan_ms = 0.0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    Tall = torch.from_numpy(res["T"].astype(np.float64)).cuda().reshape(1, 16)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(Tall) for _ in range(world)]
        dist.all_gather(gathered, Tall)
        spread = max(float((g - Tall).abs().max()) for g in gathered)
    else:
        spread = 0.0
    out = None
    if rank == 0:
        peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
        iters = res["iterations"] if a.method == "icp" else res["n_derivative_evals"]
        ach = 32.0 * kern["units"] / (kern["ms"] / 1e3) / 1e9 if kern["ms"] > 0 and a.method == "icp" else None
        err = np.abs(res["T"].astype(np.float64) - T).max()
        out = {"config": "point-sharded %s, %d-point source sharded over %d GPU(s), %d-point target replicated, gate %.4f m"
                         % (a.method, a.points, world, a.points, a.gate),
               "n_gpus": world, "iterations": int(iters), "ms_total": float(t.item()), "ms_per_iteration": float(t.item()) / max(iters, 1),
               "kernel_ms_per_launch": kern["ms"] / max(kern["launches"], 1),
               "rescan_ms_per_launch": rescan_ms / max(kern["launches"], 1),
               "roofline": {"kernel": "k_icp_stream" if ks["launches"] else "k_icp_step", "bound": "hbm", "achieved": ach, "peak": peak,
                            "unit": "GB/s", "frac": (ach / peak) if ach else None,
                            "note": "32 B algorithmic bytes per source point per launch, average over all iterations of the align "
                                    "(the first one rescans every point); real traffic of the streaming pass is 60 B per point"}
               if a.method == "icp" else None,
               "max_abs_T_error_vs_ground_truth": float(err), "max_T_spread_over_ranks": spread,
               "n_corr": res.get("n_corr"), "allreduce_bytes_per_iteration": 17 * 8 if a.method == "icp" else 28 * 8}
        if a.check_single:
            c1 = R.Context(local)
            r1 = (R.icp_align(c1, c1.upload([src]), c1.upload([tgt]), prm, want_aligned=False)[0][0] if a.method == "icp"
                  else R.ndt_align(c1, c1.upload([src]), c1.upload([tgt]), prm, want_aligned=False)[0][0])
            out["max_abs_T_diff_vs_single_gpu"] = float(np.abs(r1["T"].astype(np.float64) - res["T"].astype(np.float64)).max())
            if a.method == "icp":
                out["n_corr_single_gpu"] = r1["n_corr"]
            c1.close()
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        R.comm_destroy(ctx)
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
