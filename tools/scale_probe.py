#!/usr/bin/env python
"""Per-rank diagnosis of the pair-sharded step under torchrun: step time per rank (device events and host wall clock),
host sync latency, the GPU's PCI bus / NUMA node and the CPUs this rank may run on.  One JSON line per rank."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("tools", "realsense-pointcloud_b200"):
    sys.path.insert(0, os.path.join(ROOT, sub))


def main():
    import torch
    import torch.distributed as dist
    import gen_scene
    import rspcl_b200 as R
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = R.Context(local)
    F = int(os.environ.get("PROBE_FRAMES", "33"))
    frames, _ = gen_scene.make_sweep(2 + 1000 * rank, F)
    d = ctx.upload(list(frames), 640, 480)
    g = np.eye(4)
    g[:3, :3] = gen_scene.rot_y(-0.523599)
    icp = R.icp_params(max_iterations=50, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300, mse_threshold_absolute=-1.0)
    si = np.arange(1, F, dtype=np.int32)
    ti = np.arange(0, F - 1, dtype=np.int32)
    out = ctx.cloud(F - 1, 640 * 480)
    step = lambda: R.register_pairs(ctx, d, si, ti, R.COARSE_ICP, icp=icp, guess=g, out_transformed=out)
    for _ in range(3):
        step()
    res = {}
    for phase in ("alone" if world == 1 else "staggered", "together"):
        if world > 1:
            dist.barrier()
        if phase == "staggered":  # one rank at a time: is the slowdown contention or placement?
            ms_dev = None
            for r in range(world):
                if r == rank:
                    ctx.timer_start()
                    t0 = time.perf_counter()
                    for _ in range(5):
                        step()
                    ms_dev = ctx.timer_stop() / 5
                    ms_host = 1e3 * (time.perf_counter() - t0) / 5
                dist.barrier()
        else:
            ctx.timer_start()
            t0 = time.perf_counter()
            for _ in range(5):
                step()
            ms_dev = ctx.timer_stop() / 5
            ms_host = 1e3 * (time.perf_counter() - t0) / 5
        res[phase] = [round(ms_dev, 3), round(ms_host, 3)]
    # sync latency: 200 x (tiny op + sync)
    t0 = time.perf_counter()
    for _ in range(200):
        d.counts()
    sync_us = 1e6 * (time.perf_counter() - t0) / 200
    ctx.profile_reset()
    ctx.profile(True)
    step()
    ctx.profile(False)
    kp = ctx.profile_get("k_icp_persist")
    pr = torch.cuda.get_device_properties(local)
    bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    try:
        node = open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip()
        lcpu = open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read().strip()
    except Exception as e:
        node, lcpu = "?", type(e).__name__
    print(json.dumps({"rank": rank, "bus": bus, "gpu_numa_node": node, "gpu_local_cpus": lcpu,
                      "allowed_cpus": len(os.sched_getaffinity(0)), "step_ms_dev_host": res, "counts_roundtrip_us": round(sync_us, 1),
                      "k_icp_persist_ms": round(kp["ms"], 3), "load1": os.getloadavg()[0]}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
