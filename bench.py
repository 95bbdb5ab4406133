#!/usr/bin/env python
"""bench.py -- frame-pair registrations/s @640x480 (BASELINE.json metric) on N B200s of one node.

A step = one pass of the registration hot path over one synthetic sweep per GPU: F organized 640x480 RGB-D frames
(307,200 points each) -> F-1 pairwise edge-based ICP registrations (frame k onto frame k-1, BASELINE configs[1]
settings: -30 deg initial guess, 50 forced coarse + 50 forced fine iterations, 1 cm approximate voxel filter,
1 cm correspondence gate) including the full-cloud transformPointCloud of every source frame.

  value  : device-resident throughput (frames already in HBM as device clouds; the gray plane of the Canny input is
           rebuilt inside every step), CUDA events on the library stream.
  e2e    : the same step through the C ABI with HOST buffers inside the timed region: pinned frames uploaded, transformed
           full clouds downloaded.  Headline leg = the 16-byte `.pcd` row layout (x y z rgb: what the reference's
           --registration mode reads from and writes to disk, main.cpp:81,87, and what the CPU reference arm holds);
           `e2e_pcl32` = the same with pcl::PointXYZRGB's 32-byte in-memory layout (12 of every 32 bytes are padding).
  roofline: the ICP correspondence+reduction kernel (k_icp_persist: all iterations of a batch of pairs in one launch;
           k_icp_stream on the global-memory path), algorithmic bytes = 32 B per source point per iteration,
           timed with CUDA events bracketing each launch in one extra, untimed-for-`value` step.
  ndt    : (rank 0, N=1) BASELINE configs[2] on the side: one 1280x720 pair, 0.05 m voxels -> ms per NDT iteration and per
           derivative evaluation (the metric also names "ms per ICP/NDT iteration").
  configs0: (rank 0, N=1) BASELINE configs[0], the reference's literal `rs-pcl --registration` path: 3 frames, NDT coarse +
           ICP fine with the reference's own epsilons, accumulating target (rspcl_register_sequence) next to the oracle.
  point_sharded: BASELINE configs[4]: one 50 M-point pair, source sharded over the N ranks, 17 fp64 partial sums per
           iteration all-reduced over NVLink; ms per iteration, roofline fraction, spread over ranks, difference to 1 GPU.
  cpu_baseline: the oracle (CPU port of the reference's PCL path) on a bounded sample of the same sweep, 1 thread.
  --impl reference: the oracle with all host threads (one pair per thread) -- the reference arm.

Multi-GPU: independent sweeps, one per rank (pair-sharded, no data-path collective) -> "scaling": "weak".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for sub in ("tools", "realsense-pointcloud_b200"):
    sys.path.insert(0, os.path.join(ROOT, sub))

W, H = 640, 480
NPX = W * H
RADS = -0.523599  # icp_edge_based_registration.hpp:135


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 50; 10 for --impl reference, whose step is ~7 s of CPU)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=65, help="frames per GPU sweep (pairs = frames - 1); 65 = a 64-pair sweep (configs[3])")
    ap.add_argument("--iters", type=int, default=50, help="forced ICP iterations per stage (configs[1]: 50)")
    ap.add_argument("--coarse", default="icp", choices=["icp", "ndt"])
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--cpu-pairs", type=int, default=24, help="pairs in the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ndt", action="store_true", help="skip the side measurement of configs[2] (NDT, 1280x720)")
    ap.add_argument("--no-configs0", action="store_true", help="skip the literal rs-pcl --registration leg (configs[0])")
    ap.add_argument("--no-sharded", action="store_true", help="skip the point-sharded leg (configs[4])")
    ap.add_argument("--no-batch", action="store_true", help="skip the 4096-pair batch leg (configs[4])")
    ap.add_argument("--distinct-sweeps", action="store_true", help="N > 1: a different synthetic sweep (seed + 1000 rank) per rank instead of the same one")
    ap.add_argument("--batch-pairs", type=int, default=4096)
    ap.add_argument("--sharded-points", type=int, default=50_000_000, help="points of the source AND of the target of the point-sharded leg")
    ap.add_argument("--sharded-iters", type=int, default=30)
    ap.add_argument("--sharded-gate", type=float, default=0.002)
    ap.add_argument("--e2e-contexts", type=int, default=4)
    ap.add_argument("--e2e-chunks", type=int, default=4)
    ap.add_argument("--e2e-lock", default="nosync", choices=["sync", "nosync", "none"],
                    help="serialise transfers per direction (sync: hold the lock until the copy completed)")
    ap.add_argument("--e2e-trace", action="store_true", help="print per-chunk phase timestamps of the timed e2e pipeline to stderr")
    a = ap.parse_args()
    if a.steps is None:
        a.steps = 10 if a.impl == "reference" else 50
    return a


def forced_kw(iters):
    return dict(max_iterations=iters, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300,
                mse_threshold_absolute=-1.0)


def guess_matrix():
    import gen_scene
    g = np.eye(4)
    g[:3, :3] = gen_scene.rot_y(RADS)
    return g


def pose_err(A, B):
    D = np.linalg.inv(np.asarray(A, np.float64)) @ np.asarray(B, np.float64)
    sk = np.linalg.norm([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]]) / 2
    return float(np.arctan2(sk, (np.trace(D[:3, :3]) - 1) / 2)), float(np.linalg.norm(D[:3, 3]))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device, all_devices=None):
        self.device = device
        self.all = list(all_devices) if all_devices else [device]  # N > 1: rank 0 watches every GPU of the job
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(str(d) for d in self.all), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, per = [], [], set(), {}
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                idx, clk = int(f[0]), float(f[1])
                mx.append(float(f[2]))
            except ValueError:
                continue
            per.setdefault(idx, []).append(clk)
            if idx == self.device:
                sm.append(clk)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if len(per) > 1:
            out["per_gpu_sm_mhz"] = {str(k): float(np.median(v)) for k, v in sorted(per.items())}
        return out


# ---------------------------------------------------------------------------------------------- CPU (oracle) legs
def cpu_sweep_pairs(frames, n_pairs, iters, coarse, threads):
    """Oracle restatement of the same step on `n_pairs` pairs; returns (seconds, transforms).  TEST-INFRA use of
    oracle/ permitted here (cpu_baseline leg / reference arm only)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    from concurrent.futures import ThreadPoolExecutor
    prm = orc.icp_params(**forced_kw(iters))
    ndt = orc.ndt_params()
    guess = guess_matrix()
    orc.lib()

    def feat(k):
        e, _ = orc.extract_edges(frames[k], W, H)
        return orc.approx_voxel(e)

    def reg(i, feats):
        src, tgt = feats[i + 1], feats[i]
        if coarse == "ndt":
            c = orc.ndt_align(src, tgt, ndt, guess=guess)
        else:
            c = orc.icp_align(src, tgt, prm, guess=guess)
        f = orc.icp_align(c["aligned"], tgt, prm)
        if f["converged"]:
            orc.transform(orc.transform(frames[i + 1], c["T"]), f["T"])
        return f["T"].astype(np.float64) @ c["T"].astype(np.float64)

    t0 = time.perf_counter()
    if threads <= 1:
        feats = [feat(k) for k in range(n_pairs + 1)]
        Ts = [reg(i, feats) for i in range(n_pairs)]
    else:
        with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL inside the oracle calls
            feats = list(ex.map(feat, range(n_pairs + 1)))
            Ts = list(ex.map(lambda i: reg(i, feats), range(n_pairs)))
    return time.perf_counter() - t0, Ts


import contextlib


@contextlib.contextmanager
def stdout_to_stderr():
    """File-descriptor level: native libraries (NCCL's version banner) must not write into the one-JSON-line stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(a, rank, world):
    if rank != 0:
        return
    import gen_scene
    threads = host_threads()
    n_pairs = a.frames - 1  # the GPU arm's pairs per step, whatever the host (same config on both arms)
    frames, _ = gen_scene.make_sweep(a.seed, min(n_pairs + 1, 17))
    # bounded sample: reuse the generated frames cyclically so generation stays short
    idx = [k % len(frames) for k in range(n_pairs + 1)]
    fr = [frames[k] for k in idx]
    for _ in range(a.warmup):
        cpu_sweep_pairs(fr, min(n_pairs, threads), a.iters, a.coarse, threads)
    total = 0.0
    for _ in range(a.steps):
        dt, _ = cpu_sweep_pairs(fr, n_pairs, a.iters, a.coarse, threads)
        total += dt
    val = a.steps * n_pairs / total
    sample = ("%d pairs/step x %d steps, %d threads (one pair per thread), oracle port of the PCL path; the sweep reuses %d "
              "generated frames cyclically" % (n_pairs, a.steps, threads, len(frames)))
    print(json.dumps({
        "impl": "reference", "metric": "frame-pair registrations/sec @640x480", "value": val, "unit": "pairs/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, n_pairs + 1),
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(a, frames):
    return {"workload": "configs[1] batched: pairwise edge-based ICP over a %d-frame 640x480 sweep per GPU "
                        "(%d pairs/step), %d forced coarse + %d forced fine iterations, -30deg guess, 1cm voxel, "
                        "1cm gate, full-cloud transform" % (frames, frames - 1, a.iters, a.iters),
            "frames_per_gpu": frames, "pairs_per_step_per_gpu": frames - 1, "points_per_frame": NPX,
            "coarse": a.coarse, "icp_iterations": [a.iters, a.iters],
            "l2": "inputs larger than L2 (%.0f MB of frames per step vs 126 MB L2)" % (frames * NPX * 16 / 1e6),
            "parallelism": "pair-sharded x%d (one independent sweep per rank, no collective)" % a.gpus,
            "per_rank_data": ("a different synthetic sweep per rank (seed + 1000 rank)" if a.distinct_sweeps else
                              "the same synthetic sweep on every rank: fixed work per GPU (weak scaling)")}


def bind_to_gpu_numa_node(torch, dev):
    """Best effort: run this rank's host threads (and first-touch its pinned buffers) on the CPUs next to its GPU, so that
    N ranks do not all stream their frames through one socket's memory controller.  Returns a short description."""
    try:
        pr = torch.cuda.get_device_properties(dev)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        txt = open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        if os.environ.get("BENCH_PIN") == "1":  # diagnostic: a private pair of CPUs per rank
            world = int(os.environ.get("WORLD_SIZE", "1"))
            rank = int(os.environ.get("LOCAL_RANK", "0"))
            al = sorted(allowed)
            per = max(1, len(al) // max(world, 1))
            mine = set(al[rank * per:(rank + 1) * per])
            if mine:
                os.sched_setaffinity(0, mine)
                return "pinned to CPUs %s (BENCH_PIN)" % sorted(mine)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return "bound to %d of %d allowed CPUs local to %s" % (len(use), len(allowed), bus)
        return "no binding (%d local CPUs allowed of %d)" % (len(use), len(allowed))
    except Exception as e:  # sysfs layout differs / not permitted: run unbound
        return "no binding (%s)" % type(e).__name__


# ---------------------------------------------------------------------------------------------- side legs
def pcie_ceiling(torch, dist, world):
    """What the box's host<->device path gives when every rank moves data at once: 256 MiB pinned buffers, H2D and D2H on
    two streams at the same time, all ranks started together.  Returns (this rank's GB/s per direction, aggregate GB/s
    per direction over the ranks) -- the denominator of the e2e figure, which is transfer-bound."""
    n = 256 << 20
    h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    both()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        both()
        both()
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / 2)
    mine = n / 1e9 / best
    agg = mine
    if dist is not None:
        t = torch.tensor([mine], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        agg = float(t.item())
    return mine, agg


def ndt_config2_leg(ctx, R, gen_scene, guess, hbm_peak):
    """BASELINE configs[2]: edge-based NDT on one 1280x720 pair (921,600 points per frame), 0.05 m voxels."""
    W2, H2 = 1280, 720
    fr2, T2 = gen_scene.make_sweep(3, 2, W2, H2, noise_scale=0.2)
    d2 = ctx.upload(list(fr2), W2, H2)
    v2 = R.voxel_approx(ctx, R.edge_extract(ctx, d2)).download()
    s2, t2 = ctx.upload([v2[1]]), ctx.upload([v2[0]])
    p2 = R.ndt_params(resolution=0.05)
    R.ndt_align(ctx, s2, t2, p2, guess=guess, want_aligned=False)
    ms2 = 1e30
    for _ in range(3):  # the align as a caller sees it (no profiling scopes inside), best of three
        ctx.timer_start()
        r2, _ = R.ndt_align(ctx, s2, t2, p2, guess=guess, want_aligned=False)
        ms2 = min(ms2, ctx.timer_stop())
    ctx.profile_reset()
    ctx.profile(True)  # per-kernel breakdown from one more, profiled align
    R.ndt_align(ctx, s2, t2, p2, guess=guess, want_aligned=False)
    ctx.profile(False)
    ke, kc = ctx.profile_get("k_ndt_eval"), ctx.profile_get("k_ndt_control")
    kp = ctx.profile_get("k_ndt_persist")
    kb = ctx.profile_get("ndt_voxel_build")
    ang2, tr2 = pose_err(r2[0]["T"], gen_scene.pairwise_gt(T2, 1))
    leg = {"workload": "configs[2]: edge-based NDT, one 1280x720 pair (921,600 points per frame), 0.05 m voxels, step 0.1, eps 0.01",
           "edge_points_src_tgt": [int(len(v2[1])), int(len(v2[0]))], "iterations": r2[0]["iterations"],
           "derivative_evals": r2[0]["n_derivative_evals"], "converged": bool(r2[0]["converged"]),
           "ms_align": ms2, "ms_per_ndt_iteration": ms2 / max(r2[0]["iterations"], 1),
           "ms_voxel_build": kb["ms"], "err_vs_ground_truth_rad_m": [ang2, tr2]}
    evals = max(r2[0]["n_derivative_evals"] + r2[0]["n_hessian_evals"], 1)
    if kp["launches"]:
        leg["ms_per_derivative_eval"] = kp["ms"] / evals
        leg["kernel"] = "k_ndt_persist (all evaluations of the align in one launch)"
        # units = (source point, neighbour voxel) pairs the evaluations touched, in gradient-evaluation equivalents
        # (~135 fp64 operations per pair, ~456 when the Hessian is evaluated too: counted from ndt_point_eval in
        # csrc/ndt.cu); bytes per pair: 16 B source point + 72 B voxel (mean + inverse covariance)
        pairs = kp["units"]
        flops = pairs * 135.0
        fp64_peak = 148 * 64 * 2 * 1.965e9 / 1e12  # B200: 64 fp64 FMA / clk / SM
        leg["roofline"] = {"bound": "fp64 (latency / launch-bound at ~6 k points: see note)", "point_voxel_pairs": pairs,
                           "gflops": flops / (kp["ms"] / 1e3) / 1e9 if kp["ms"] > 0 else None,
                           "frac_fp64": (flops / (kp["ms"] / 1e3) / 1e12) / fp64_peak if kp["ms"] > 0 else None,
                           "fp64_peak_tflops_nominal": fp64_peak,
                           "hbm_frac": (pairs * 88.0 / (kp["ms"] / 1e3) / 1e9) / hbm_peak if kp["ms"] > 0 else None,
                           "note": "one 6 k-point pair occupies one cluster: the align is bound by the dependent chain of the "
                                   "Newton / More-Thuente evaluations, not by the FP64 pipe or HBM; the pipe fraction is what a "
                                   "batch of pairs multiplies"}
    else:
        leg["ms_per_derivative_eval"] = ke["ms"] / max(ke["launches"], 1)
        leg["ms_per_control_step"] = kc["ms"] / max(kc["launches"], 1)
    return leg


def configs0_leg(ctx, R, gen_scene, a):
    """BASELINE configs[0]: `rs-pcl --registration` on 3 synthetic 640x480 frames with the reference's literal settings --
    NDT coarse (eps 0.01, step 0.1, 1 m voxels, <= 50 iterations) + ICP fine (eps 1 / 1000: one iteration), -30 deg per
    frame, accumulating target -- through rspcl_register_sequence, next to the oracle's line-by-line scheme on one thread."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    fr3, T3 = gen_scene.make_sweep(1, 3)
    n = len(fr3)
    guesses = np.stack([np.eye(4)] * n)
    acc = np.float32(0)
    for k in range(1, n):
        acc = np.float32(acc + np.float32(RADS))
        guesses[k][:3, :3] = gen_scene.rot_y(float(acc))
    h_in = ctx.pinned(n * NPX * 16)
    h_in.view(R.POINT)[:] = np.concatenate(list(fr3))
    h_out = ctx.pinned(n * NPX * 16)
    cnt = np.full(n, NPX, np.int32)
    d_fr = ctx.cloud(n, NPX)
    d_out = ctx.cloud(1, n * NPX)
    import ctypes as C
    L = R.lib()
    oc = np.zeros(1, np.int32)

    def up():
        d_fr.upload_raw(h_in.ctypes.data_as(C.c_void_p), cnt, W, H, R.LAYOUT_PCD16)

    def run():
        d_fr.invalidate_gray()
        return R.register_sequence(ctx, d_fr, guesses, R.COARSE_NDT, out_global=d_out)[0]

    def down():
        ctx.check(L.rspcl_cloud_download(ctx.h, d_out.h, h_out.ctypes.data_as(C.c_void_p), R.LAYOUT_PCD16, C.c_longlong(n * NPX),
                                         oc.ctypes.data_as(C.c_void_p)))

    up()
    res = run()
    reps = 10
    ctx.timer_start()
    for _ in range(reps):
        res = run()
    ms_dev = ctx.timer_stop() / reps
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        up()
        run()
        down()
    ms_e2e = 1e3 * (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    o = orc.scheme_edge(fr3.reshape(-1), W, H, "ndt")
    cpu_s = time.perf_counter() - t0
    errs_o, errs_gt = [], []
    for k in range(1, n):
        T = R.c_to_mat(res[k].T_fine).astype(np.float64) @ R.c_to_mat(res[k].T_coarse).astype(np.float64)
        errs_o.append(pose_err(T, o["T"][k]))
        errs_gt.append(pose_err(T, T3[k]))
    return {"workload": "configs[0]: rs-pcl --registration, 3 frames 640x480, NDT coarse + ICP fine, reference-literal settings, "
                        "-30 deg per frame, accumulating target (rspcl_register_sequence)",
            "registrations": n - 1, "accepted": [int(r.converged) for r in res],
            "ndt_iterations": [int(r.coarse_iterations) for r in res[1:]],
            "gpu_ms_device_resident": ms_dev, "gpu_ms_e2e_host_pcd16": ms_e2e,
            "pairs_per_s_device_resident": (n - 1) / (ms_dev / 1e3), "pairs_per_s_e2e": (n - 1) / (ms_e2e / 1e3),
            "cpu_baseline": {"value": (n - 1) / cpu_s, "unit": "pairs/s", "cores": 1, "kind": "port",
                             "sample": "the same 3 frames, oracle scheme (orc_scheme_edge), single thread, %.2f s" % cpu_s},
            "speedup_e2e_vs_cpu_1_thread": cpu_s * 1e3 / ms_e2e,
            "max_err_vs_oracle_rad_m_frame1": [float(errs_o[0][0]), float(errs_o[0][1])],
            "max_err_vs_oracle_rad_m_later_frames": [float(max(e[0] for e in errs_o[1:])), float(max(e[1] for e in errs_o[1:]))],
            "max_err_vs_ground_truth_rad_m": [float(max(e[0] for e in errs_gt)), float(max(e[1] for e in errs_gt))],
            "note": "frames >= 2 of a chained NDT differ from the oracle's own chain at the method's accuracy (ulp-level chaos of "
                    "the iteration count, tests/test_gpu_facade.py); every stage matches the oracle to 1e-4 on identical inputs"}


def point_sharded_leg(ctx, R, gen_scene, torch, dist, rank, world, dev, a, hbm_peak):
    """BASELINE configs[4]: one pair of --sharded-points points each (uniform samples of the room surface); the source is
    split over the ranks, the target replicated; per iteration the 17 fp64 partial sums are all-reduced on the library
    stream and every rank runs the identical solve.  Rank 0 then aligns the WHOLE source alone for the 1-vs-N check."""
    P, iters = a.sharded_points, a.sharded_iters
    if world > 1:
        uid = torch.from_numpy(R.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
        dist.broadcast(uid, 0)
        # (NCCL prints its version banner to stdout when the library creates its communicator: keep stdout = the JSON line)
        with stdout_to_stderr():
            R.comm_init(ctx, world, rank, uid.cpu().numpy())
    tgt = gen_scene.sample_room_surface(a.seed + 77, P)
    Tm = np.eye(4)
    Tm[:3, :3] = gen_scene.rot_axis([0.3, 1.0, 0.2], 0.0004)
    Tm[:3, 3] = [0.0006, -0.0004, 0.0005]
    Ti = np.linalg.inv(Tm)

    def shard(r):  # source shard r: its own seeded sample, moved by the inverse of the transform the align must recover
        lo, hi = r * P // world, (r + 1) * P // world
        s0 = gen_scene.sample_room_surface(a.seed + 78 + r, hi - lo)
        xyz = np.stack([s0["x"], s0["y"], s0["z"]], 1).astype(np.float64) @ Ti[:3, :3].T + Ti[:3, 3]
        s0["x"], s0["y"], s0["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
        return s0

    d_tgt = ctx.upload([tgt])
    d_src = ctx.upload([shard(rank)])
    prm = R.icp_params(max_iterations=iters, max_corr_dist=a.sharded_gate, transformation_epsilon=-1.0,
                       euclidean_fitness_epsilon=-1e300, mse_threshold_absolute=-1.0)
    R.icp_align_sharded(ctx, d_src, d_tgt, prm)  # warm-up (pool growth, NCCL channels)
    if dist is not None:
        dist.barrier()
    ctx.profile_reset()
    ctx.profile(True)
    ctx.timer_start()
    res = R.icp_align_sharded(ctx, d_src, d_tgt, prm)[0][0]
    ms = ctx.timer_stop()
    ctx.profile(False)
    ks, kr, ka = ctx.profile_get("k_icp_stream"), ctx.profile_get("k_icp_rescan"), ctx.profile_get("allreduce")
    kf = ctx.profile_get("k_icp_stream_first")
    ksol = ctx.profile_get("k_icp_solve")
    # the same align with the partial sums going through ncclAllReduce instead of the peer-memory exchange
    nccl_cmp = None
    if dist is not None:
        os.environ["RSPCL_PEER_XCHG"] = "0"
        R.icp_align_sharded(ctx, d_src, d_tgt, prm)
        dist.barrier()
        ctx.profile_reset()
        ctx.profile(True)
        ctx.timer_start()
        res_n = R.icp_align_sharded(ctx, d_src, d_tgt, prm)[0][0]
        ms_n = ctx.timer_stop()
        ctx.profile(False)
        ka_n, ksol_n = ctx.profile_get("allreduce"), ctx.profile_get("k_icp_solve")
        del os.environ["RSPCL_PEER_XCHG"]
        tn = torch.tensor([ms_n], dtype=torch.float64, device="cuda")
        dist.all_reduce(tn, op=dist.ReduceOp.MAX)
        nccl_cmp = {"ms_per_iteration": float(tn.item()) / max(res_n["iterations"], 1),
                    "sum_allreduce_solve_us_per_iteration": 1e3 * ksol_n["ms"] / max(ksol_n["launches"], 1),
                    "ncclAllReduce_us_per_iteration": 1e3 * ka_n["ms"] / max(ka_n["launches"], 1) if ka_n["launches"] else None,
                    "same_transform_as_peer_path": bool(np.array_equal(res_n["T"], res["T"]))}
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    Tall = torch.from_numpy(res["T"].astype(np.float64)).cuda().reshape(1, 16)
    spread = 0.0
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(Tall) for _ in range(world)]
        dist.all_gather(gathered, Tall)
        spread = max(float((g - Tall).abs().max()) for g in gathered)
    out = None
    if rank == 0:
        ms_all = float(t.item())
        # algorithmic bytes: 32 B per source point per iteration, all ranks' points, against N x the HBM peak
        ach = 32.0 * P * res["iterations"] / (ms_all / 1e3) / 1e9
        ach_stream = 32.0 * ks["units"] / (ks["ms"] / 1e3) / 1e9 if ks["ms"] > 0 else None
        out = {"workload": "configs[4]: point-sharded ICP, %d-point source sharded over %d GPU(s), %d-point target replicated, "
                           "gate %.4f m, %d forced iterations" % (P, world, P, a.sharded_gate, iters),
               "n_gpus": world, "iterations": int(res["iterations"]), "ms_total": ms_all,
               "ms_per_iteration": ms_all / max(res["iterations"], 1),
               "stream_kernel_ms_per_launch": ks["ms"] / max(ks["launches"], 1),
               "first_iteration_full_query_ms": kf["ms"],
               "rescan_ms_per_launch": kr["ms"] / max(kr["launches"], 1),
               "exchange": ("one-shot peer-memory all-reduce over NVLink fused with the solve (k_icp_solve_peer)" if world > 1 and not ka["launches"]
                            else ("ncclAllReduce" if world > 1 else "none (1 GPU)")),
               "sum_exchange_solve_us_per_iteration": 1e3 * ksol["ms"] / max(ksol["launches"], 1),
               "allreduce_bytes_per_iteration": 17 * 8,
               "nccl_allreduce_variant": nccl_cmp,
               "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak * world, "unit": "GB/s", "frac": ach / (hbm_peak * world),
                            "stream_kernel_frac_this_rank": (ach_stream / hbm_peak) if ach_stream else None,
                            "note": "whole iteration (streaming pass + exact re-queries + solve + all-reduce), 32 algorithmic "
                                    "bytes per source point, all ranks, against N x the measured HBM peak"},
               "n_corr": int(res["n_corr"]), "max_T_spread_over_ranks": spread,
               "max_abs_T_error_vs_ground_truth": float(np.abs(res["T"].astype(np.float64) - Tm).max())}
        if world > 1:
            c1 = R.Context(dev)
            whole = np.concatenate([shard(r) for r in range(world)])
            r1 = R.icp_align(c1, c1.upload([whole]), c1.upload([tgt]), prm, want_aligned=False)[0][0]
            out["diff_vs_1gpu_max_abs_T"] = float(np.abs(r1["T"].astype(np.float64) - res["T"].astype(np.float64)).max())
            out["n_corr_1gpu"] = int(r1["n_corr"])
            c1.close()
    if dist is not None:
        dist.barrier()
        R.comm_destroy(ctx)
    return out


def batch_leg(ctx, R, frames, guess, icp, ndt, coarse, a, W, H, peak):
    """BASELINE configs[4], first half: `--batch-pairs` (4096) independent frame-pair registrations in ONE rspcl_register_pairs
    call (more pairs than SMs: every pair is one CTA of k_icp_persist, waves of pairs per SM).  The batch is copies of the
    step's 65-frame sweep -- every copy with its own device frames, edges, voxel clouds and pair states -- so every block
    must reproduce block 0."""
    import ctypes as C
    F = len(frames)
    blocks = max(1, a.batch_pairs // (F - 1))
    n_pairs = blocks * (F - 1)
    one = np.concatenate(list(frames))
    host = np.tile(one, blocks)
    d_frames = ctx.cloud(blocks * F, W * H)
    counts = np.full(blocks * F, W * H, np.int32)
    d_frames.upload_raw(host.ctypes.data_as(C.c_void_p), counts, W, H, R.LAYOUT_PCD16)
    ctx.sync()
    del host
    d_out = ctx.cloud(n_pairs, W * H)
    src = np.concatenate([b * F + np.arange(1, F) for b in range(blocks)]).astype(np.int32)
    tgt = (src - 1).astype(np.int32)
    R.register_pairs(ctx, d_frames, src, tgt, coarse, icp=icp, ndt=ndt, guess=guess, out_transformed=d_out)  # warm-up (scratch pool)
    times = []
    for _ in range(3):  # best of three: the call allocates ~60 GB of stream-ordered scratch and its host side jitters
        ctx.profile_reset()
        ctx.profile(True)
        ctx.timer_start()
        d_frames.invalidate_gray()
        res = R.register_pairs(ctx, d_frames, src, tgt, coarse, icp=icp, ndt=ndt, guess=guess, out_transformed=d_out)
        times.append(ctx.timer_stop())
        ctx.profile(False)
    ms = min(times)
    kp = ctx.profile_get("k_icp_persist")
    ref = [np.array(res[k].T_fine) for k in range(F - 1)]
    worst, bitwise = 0.0, 0
    for k in range(n_pairs):
        d = float(np.abs(np.array(res[k].T_fine) - ref[k % (F - 1)]).max())
        worst = max(worst, d)
        bitwise += d == 0.0
    achieved = 32.0 * kp["units"] / (kp["ms"] / 1e3) / 1e9 if kp["ms"] > 0 else 0.0
    return {"workload": "configs[4]: %d independent 640x480 frame pairs in one call (%d resident frames), same stages and forced "
                        "iteration counts as the step" % (n_pairs, blocks * F),
            "pairs": n_pairs, "ms": ms, "ms_all_runs": times, "pairs_per_s": n_pairs / (ms * 1e-3), "k_icp_persist_ms": kp["ms"],
            "k_icp_persist_launches": kp["launches"],
            "k_icp_persist_roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak},
            "pairs_converged": int(sum(int(r.converged) for r in res)),
            "max_abs_T_diff_vs_block0": worst, "pairs_bitwise_equal_to_block0": int(bitwise),
            "note": "one CTA per pair (no cluster exchange, no idle SMs): the per-pair cost of k_icp_persist drops against the "
                    "64-pair step, where 64 pairs have to be spread over 148 SMs as clusters"}


# ---------------------------------------------------------------------------------------------- GPU arm
def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return

    import torch
    import gen_scene
    import rspcl_b200 as R

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    numa = bind_to_gpu_numa_node(torch, dev)
    ctx = R.Context(dev)

    F = a.frames
    n_pairs = F - 1
    # weak scaling = fixed work per GPU: every rank registers the SAME synthetic sweep unless --distinct-sweeps is given
    # (different scene content costs 1.98 - 2.73 ms per step, and the slowest rank is what gets reported: measured at N = 8)
    frames, Tgt = gen_scene.make_sweep(a.seed + (1000 * rank if a.distinct_sweeps else 0), F)
    guess = guess_matrix()
    icp = R.icp_params(**forced_kw(a.iters))
    ndt = R.ndt_params()
    coarse = R.COARSE_NDT if a.coarse == "ndt" else R.COARSE_ICP
    src_idx = np.arange(1, F, dtype=np.int32)
    tgt_idx = np.arange(0, F - 1, dtype=np.int32)

    # pinned host frames in both host layouts + pinned result buffers
    LAY = {"pcd16": (R.LAYOUT_PCD16, R.POINT, 16), "pcl32": (R.LAYOUT_PCL32, R.PCL32, 32)}
    h_in, h_out = {}, {}
    for name, (_, dt, esz) in LAY.items():
        h_in[name] = ctx.pinned(F * NPX * esz)
        h_out[name] = ctx.pinned(n_pairs * NPX * esz)
    h_in["pcd16"].view(R.POINT)[:] = np.concatenate(list(frames))
    h_in["pcl32"].view(R.PCL32)[:] = np.concatenate([R.to_pcl32(f) for f in frames])
    counts = np.full(F, NPX, np.int32)
    d_frames = ctx.cloud(F, NPX)
    d_out = ctx.cloud(n_pairs, NPX)
    out_counts = np.zeros(n_pairs, np.int32)
    L = R.lib()
    import ctypes as C

    def upload():
        d_frames.upload_raw(h_in["pcd16"].ctypes.data_as(C.c_void_p), counts, W, H, R.LAYOUT_PCD16)

    def step():
        d_frames.invalidate_gray()  # the (r+g+b)/3 plane of the Canny input is rebuilt inside the step
        return R.register_pairs(ctx, d_frames, src_idx, tgt_idx, coarse, icp=icp, ndt=ndt, guess=guess, out_transformed=d_out)

    def download():
        ctx.check(L.rspcl_cloud_download(ctx.h, d_out.h, h_out["pcd16"].ctypes.data_as(C.c_void_p), R.LAYOUT_PCD16,
                                         C.c_longlong(n_pairs * NPX), out_counts.ctypes.data_as(C.c_void_p)))

    def barrier():
        if dist is not None:
            dist.barrier()
        ctx.sync()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- correctness guard: every pair converges and lands near the ground truth
    upload()
    res = step()
    n_conv = sum(int(r.converged) for r in res)
    errs = []
    for i in range(n_pairs):
        T = R.c_to_mat(res[i].T_fine).astype(np.float64) @ R.c_to_mat(res[i].T_coarse).astype(np.float64)
        errs.append(pose_err(T, gen_scene.pairwise_gt(Tgt, i + 1)))
    max_ang, max_tr = max(e[0] for e in errs), max(e[1] for e in errs)
    if n_conv != n_pairs or max_ang > 0.01 or max_tr > 0.02:
        raise SystemExit("bench sanity check failed: converged %d/%d, max err %.4g rad %.4g m" % (n_conv, n_pairs, max_ang, max_tr))
    mean_src = float(np.mean([r.n_src for r in res]))

    # ---- device-resident throughput
    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(dev, range(world) if world > 1 else None)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches()
    ctx.timer_start()
    for _ in range(a.steps):
        step()
    ms = ctx.timer_stop()
    l1 = ctx.launches()
    barrier()
    ms_rank = [ms]
    if dist is not None:
        t_all = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(t_all, torch.tensor([ms], dtype=torch.float64, device="cuda"))
        ms_rank = [float(t.item()) for t in t_all]
    ms = max_over_ranks(ms)
    value = world * n_pairs * a.steps / (ms / 1e3)

    # ---- roofline of the dominant kernel, from per-launch CUDA events in one extra profiled step (taken right after the
    # timed region: the e2e legs below run several contexts side by side on this GPU, which is a different regime)
    ctx.profile_reset()
    ctx.profile(True)
    ctx.timer_start()
    step()
    ms_prof = ctx.timer_stop()
    ctx.profile(False)
    prof_snapshot = {k: ctx.profile_get(k) for k in ("k_icp_persist", "k_icp_step", "k_icp_stream", "k_icp_rescan", "k_icp_solve", "grid_build",
                                                     "k_canny_nms", "edge_hysteresis_compact", "k_approx_voxel", "k_transform2",
                                                     "k_ndt_eval", "ndt_voxel_build")}

    # ---- end to end through the C ABI with host buffers: the sweep is cut into chunks of pairs that are pipelined over
    # several contexts (one stream + one host thread each), so chunk c+1's H2D, chunk c's kernels and chunk c-1's D2H
    # overlap (PCIe is full duplex).  Every frame still crosses PCIe (a chunk re-uploads its one boundary frame).
    from concurrent.futures import ThreadPoolExecutor
    import contextlib
    n_ctx = a.e2e_contexts
    n_chunks = max(d for d in range(1, a.e2e_chunks + 1) if n_pairs % d == 0)  # equal chunks keep the batch shape fixed
    cp = n_pairs // n_chunks
    chunks = [(c * cp, (c + 1) * cp) for c in range(n_chunks)]
    workers = []
    for w in range(n_ctx):
        cw = ctx if w == 0 else R.Context(dev)
        workers.append({"ctx": cw, "frames": cw.cloud(cp + 1, NPX), "out": cw.cloud(cp, NPX), "oc": np.zeros(cp, np.int32)})
    cnts = np.full(cp + 1, NPX, np.int32)
    si = np.arange(1, cp + 1, dtype=np.int32)
    ti = np.arange(0, cp, dtype=np.int32)
    h2d_lock, d2h_lock = (contextlib.nullcontext(), contextlib.nullcontext()) if a.e2e_lock == "none" else (threading.Lock(), threading.Lock())
    pool = ThreadPoolExecutor(n_ctx)
    trace = []

    def e2e_leg(name, xyz_mode=None):
        """One timed e2e region in host layout `name`; xyz_mode (pcl32 only): download x y z 1 only (0: 2-D DMA, 1: kernel
        stores into the pinned buffer), the caller's buffer keeping its colours."""
        lay, dt, esz = LAY[name]
        in_rows = h_in[name].view(dt).reshape(F, NPX)
        out_rows = h_out[name].view(dt).reshape(n_pairs, NPX)

        def run_chunks(w, steps, timed):
            W_ = workers[w]
            cw = W_["ctx"]
            if timed:
                cw.timer_start()
            for st_i in range(steps):
                for ci in range(w, n_chunks, n_ctx):
                    lo, hi = chunks[ci]
                    # one transfer per direction at a time: the contexts fall into a staggered pipeline (A computes while B
                    # uploads and C downloads) instead of moving in lockstep and sharing each PCIe direction
                    t0 = time.perf_counter()
                    with h2d_lock:
                        t1 = time.perf_counter()
                        W_["frames"].upload_raw(in_rows[lo:hi + 1].ctypes.data_as(C.c_void_p), cnts, W, H, lay)
                        if a.e2e_lock == "sync":
                            cw.sync()
                    t2 = time.perf_counter()
                    R.register_pairs(cw, W_["frames"], si, ti, coarse, icp=icp, ndt=ndt, guess=guess, out_transformed=W_["out"])
                    t3 = time.perf_counter()
                    with d2h_lock:
                        t4 = time.perf_counter()
                        if xyz_mode is None:
                            cw.check(L.rspcl_cloud_download(cw.h, W_["out"].h, out_rows[lo:hi].ctypes.data_as(C.c_void_p), lay,
                                                            C.c_longlong(cp * NPX), W_["oc"].ctypes.data_as(C.c_void_p)))
                        else:
                            cw.check(L.rspcl_cloud_download_xyz_pcl32(cw.h, W_["out"].h, out_rows[lo:hi].ctypes.data_as(C.c_void_p),
                                                                      C.c_longlong(cp * NPX), int(xyz_mode)))
                    t5 = time.perf_counter()
                    if timed and a.e2e_trace:
                        trace.append((name, w, st_i, ci, t0, t1, t2, t3, t4, t5))
            if timed:
                cw.timer_mark()

        list(pool.map(lambda w: run_chunks(w, max(1, a.warmup - 1), False), range(n_ctx)))
        barrier()
        list(pool.map(lambda w: run_chunks(w, a.steps, True), range(n_ctx)))
        ms_leg = max_over_ranks(R.timer_span([w["ctx"] for w in workers]))
        barrier()
        d2h = n_pairs * NPX * (esz if xyz_mode is None else 16) + n_pairs * 160
        return {"value": world * n_pairs * a.steps / (ms_leg / 1e3), "unit": "pairs/s", "layout": name,
                "h2d_bytes_per_step": (n_pairs + len(chunks)) * NPX * esz, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_leg / a.steps}, out_rows

    e2e16, rows16 = e2e_leg("pcd16")
    # the pipelined path must reproduce the single-context result
    chk = rows16[0].copy()
    download()
    ref0 = h_out["pcd16"].view(R.POINT).reshape(n_pairs, NPX)[0]
    # (bit equality is not guaranteed: the cluster sizes, hence the fp64 summation order, depend on the batch size)
    dmax = max(float(np.abs(chk[a_] - ref0[a_]).max()) for a_ in "xyz")
    assert dmax < 1e-5 and np.array_equal(chk["rgba"], ref0["rgba"]), "pipelined e2e result differs from the single-context result"
    e2e32, rows32 = e2e_leg("pcl32")
    assert np.array_equal(rows32[0]["rgba"], ref0["rgba"]) and max(float(np.abs(rows32[0][a_] - ref0[a_]).max()) for a_ in "xyz") < 1e-5
    # PCL32, x y z only: transformPointCloud changes 12 of the 32 bytes of a point; is a strided 16-of-32-byte read-back
    # (2-D DMA copy or a kernel storing over PCIe) faster than moving all 32?  Measured on one chunk, the faster way is kept.
    probe = None
    if rank == 0 and world == 1:
        cw, Wk = workers[0]["ctx"], workers[0]
        Wk["frames"].upload_raw(h_in["pcl32"].view(R.PCL32).reshape(F, NPX)[0:cp + 1].ctypes.data_as(C.c_void_p), cnts, W, H, R.LAYOUT_PCL32)
        R.register_pairs(cw, Wk["frames"], si, ti, coarse, icp=icp, ndt=ndt, guess=guess, out_transformed=Wk["out"])
        cw.sync()
        probe = {"points": cp * NPX}
        dst = rows32[0:cp]
        for label, fn in (("full_32B", lambda: cw.check(L.rspcl_cloud_download(cw.h, Wk["out"].h, dst.ctypes.data_as(C.c_void_p), R.LAYOUT_PCL32, C.c_longlong(cp * NPX), Wk["oc"].ctypes.data_as(C.c_void_p)))),
                          ("xyz_2d_dma", lambda: cw.check(L.rspcl_cloud_download_xyz_pcl32(cw.h, Wk["out"].h, dst.ctypes.data_as(C.c_void_p), C.c_longlong(cp * NPX), 0))),
                          ("xyz_kernel_store", lambda: cw.check(L.rspcl_cloud_download_xyz_pcl32(cw.h, Wk["out"].h, dst.ctypes.data_as(C.c_void_p), C.c_longlong(cp * NPX), 1)))):
            fn()
            t0 = time.perf_counter()
            for _ in range(3):
                fn()
            probe[label + "_ms"] = 1e3 * (time.perf_counter() - t0) / 3
        ok_xyz = np.array_equal(dst[0]["rgba"], ref0["rgba"]) and (dst[0]["w"] == 1.0).all()
        probe["xyz_paths_keep_colours"] = bool(ok_xyz)
    clocks = sampler.stop() if rank == 0 else None  # sampled over the timed regions (device-resident and e2e)
    pcie_rank, pcie_agg = pcie_ceiling(torch, dist, world)
    for leg in (e2e16, e2e32):
        bps = max(leg["h2d_bytes_per_step"], leg["d2h_bytes_per_step"]) * world / (leg["ms_per_step"] / 1e3) / 1e9
        leg["pcie"] = {"busier_direction_GBps_all_ranks": bps, "measured_ceiling_GBps_all_ranks_concurrent": pcie_agg,
                       "frac_of_ceiling": bps / pcie_agg if pcie_agg > 0 else None}
    if a.e2e_trace and trace:
        tz = min(t[4] for t in trace)
        for t in sorted(trace, key=lambda t: t[4]):
            sys.stderr.write("%s ctx %d step %d chunk %d: wait_h2d %.2f upload %.2f compute %.2f wait_d2h %.2f download %.2f  [start %.2f end %.2f ms]\n" % (
                t[0], t[1], t[2], t[3], 1e3 * (t[5] - t[4]), 1e3 * (t[6] - t[5]), 1e3 * (t[7] - t[6]), 1e3 * (t[8] - t[7]),
                1e3 * (t[9] - t[8]), 1e3 * (t[4] - tz), 1e3 * (t[9] - tz)))

    kern = prof_snapshot
    dom = "k_icp_persist" if kern["k_icp_persist"]["launches"] else ("k_icp_stream" if kern["k_icp_stream"]["launches"] else "k_icp_step")
    ki = kern[dom]
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = 32.0 * ki["units"] / (ki["ms"] / 1e3) / 1e9 if ki["ms"] > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", dom + "_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": 32.0 * ki["units"] / max(ki["launches"], 1),
                "avg_launch_us": 1e3 * ki["ms"] / max(ki["launches"], 1), "launches_per_step": ki["launches"],
                "share_of_step": ki["ms"] / ms_prof if ms_prof > 0 else None,
                "units": "source points x executed iterations (32 B each: 16 R source + 16 R matched target)",
                "note": "persistent kernel, one timed span = the side-by-side launches (one per cluster size) that run the coarse AND "
                        "the fine align of every pair: the target grid lives in shared memory and the source slice in registers, "
                        "so DRAM traffic is a few percent of the algorithmic bytes; the kernel is bound by the serial chain of one ICP "
                        "iteration (cache test, re-query, reduce, cluster exchange, solve) on the slowest pair, not by HBM (SURVEY "
                        "H3).  The HBM-bound regime (one 50 M-point pair, k_icp_stream) is the point_sharded leg."}

    # ---- side measurement (rank 0, N=1): BASELINE configs[2], edge-based NDT on one 1280x720 pair, 0.05 m voxels
    ndt_leg = None
    if rank == 0 and world == 1 and not a.no_ndt:
        ndt_leg = ndt_config2_leg(ctx, R, gen_scene, guess, peak)

    # ---- side measurement (rank 0, N=1): BASELINE configs[0], the literal rs-pcl --registration path on 3 frames
    cfg0 = None
    if rank == 0 and world == 1 and not a.no_configs0:
        cfg0 = configs0_leg(ctx, R, gen_scene, a)

    # ---- BASELINE configs[4], first half (rank 0, N=1): 4096 pairs in one call
    batch = None
    if rank == 0 and world == 1 and not a.no_batch:
        batch = batch_leg(ctx, R, frames, guess, icp, ndt, coarse, a, W, H, peak)

    # ---- BASELINE configs[4]: one huge pair, source points sharded over the ranks, partial sums all-reduced
    sharded = None
    if not a.no_sharded:
        sharded = point_sharded_leg(ctx, R, gen_scene, torch, dist, rank, world, dev, a, peak)

    # ---- CPU baseline (rank 0, N=1 only): oracle, one thread, bounded sample of the same sweep
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        m = min(a.cpu_pairs, n_pairs)
        dt, Tcpu = cpu_sweep_pairs(list(frames[:m + 1]), m, a.iters, a.coarse, 1)
        worst = 0.0
        for i in range(m):
            T = R.c_to_mat(res[i].T_fine).astype(np.float64) @ R.c_to_mat(res[i].T_coarse).astype(np.float64)
            ang, tr = pose_err(T, Tcpu[i])
            worst = max(worst, ang, tr)
        cpu = {"value": m / dt, "unit": "pairs/s", "cores": 1, "kind": "port",
               "sample": "first %d pairs of the same sweep, single thread (the reference path is single-threaded), %.1f s; "
                         "max |GPU - CPU| transform difference %.2e (rad or m)" % (m, dt, worst),
               "host_threads_available": host_threads()}

    if rank == 0:
        e2e16.update({"pipeline": "%d chunks over %d contexts (streams), pinned host buffers" % (len(chunks), n_ctx), "host_numa": numa,
                      "note": "host layout = the 16-byte .pcd row (x y z rgb) the reference's --registration mode reads and writes"})
        e2e32.update({"note": "host layout = pcl::PointXYZRGB in memory (32 B per point, 12 of them padding)",
                      "download_probe_one_chunk": probe})
        out = {
            "metric": "frame-pair registrations/sec @640x480", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "ms_per_step_per_rank": [m / a.steps for m in ms_rank],
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, F),
            "clocks": clocks,
            "e2e": e2e16,
            "e2e_pcl32": e2e32,
            "gpu_launches": int(l1 - l0),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "kernels_ms_per_step": {k: v["ms"] for k, v in kern.items() if v["launches"]},
            "ndt": ndt_leg,
            "configs0": cfg0,
            "batch4096": batch,
            "point_sharded": sharded,
            "ms_per_icp_iteration": (kern[dom]["ms"] + kern["k_icp_solve"]["ms"]) / (2.0 * a.iters if dom == "k_icp_persist" else max(kern[dom]["launches"], 1)),
            "check": {"pairs_converged": n_conv, "max_err_vs_ground_truth": [max_ang, max_tr],
                      "mean_source_edge_points": mean_src},
        }
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
