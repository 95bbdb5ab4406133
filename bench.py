#!/usr/bin/env python
"""bench.py -- frame-pair registrations/s @640x480 (BASELINE.json metric) on N B200s of one node.

A step = one pass of the registration hot path over one synthetic sweep per GPU: F organized 640x480 RGB-D frames
(307,200 points each) -> F-1 pairwise edge-based ICP registrations (frame k onto frame k-1, BASELINE configs[1]
settings: -30 deg initial guess, 50 forced coarse + 50 forced fine iterations, 1 cm approximate voxel filter,
1 cm correspondence gate) including the full-cloud transformPointCloud of every source frame.

  value  : device-resident throughput (frames already in HBM as device clouds), CUDA events on the library stream.
  e2e    : the same step through the C ABI with HOST buffers: pinned pcl::PointXYZRGB (32 B/pt) frames uploaded and the
           transformed full clouds downloaded inside the timed region.
  roofline: the ICP correspondence+reduction kernel (k_icp_persist: all iterations of a batch of pairs in one launch;
           k_icp_stream on the global-memory path), algorithmic bytes = 32 B per source point per iteration,
           timed with CUDA events bracketing each launch in one extra, untimed-for-`value` step.
  ndt    : (rank 0, N=1) BASELINE configs[2] on the side: one 1280x720 pair, 0.05 m voxels -> ms per NDT iteration and per
           derivative evaluation (the metric also names "ms per ICP/NDT iteration").
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=65, help="frames per GPU sweep (pairs = frames - 1); 65 = a 64-pair sweep (configs[3])")
    ap.add_argument("--iters", type=int, default=50, help="forced ICP iterations per stage (configs[1]: 50)")
    ap.add_argument("--coarse", default="icp", choices=["icp", "ndt"])
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--cpu-pairs", type=int, default=24, help="pairs in the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ndt", action="store_true", help="skip the side measurement of configs[2] (NDT, 1280x720)")
    ap.add_argument("--e2e-contexts", type=int, default=4)
    ap.add_argument("--e2e-chunks", type=int, default=4)
    ap.add_argument("--e2e-lock", default="nosync", choices=["sync", "nosync", "none"],
                    help="serialise transfers per direction (sync: hold the lock until the copy completed)")
    ap.add_argument("--e2e-trace", action="store_true", help="print per-chunk phase timestamps of the timed e2e pipeline to stderr")
    return ap.parse_args()


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

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
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
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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
    n_pairs = max(8, min(2 * threads, 64))
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
    sample = "%d pairs/step x %d steps, %d threads (one pair per thread), oracle port of the PCL path" % (n_pairs, a.steps, threads)
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
            "parallelism": "pair-sharded x%d (independent sweeps, no collective)" % a.gpus}


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
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return "bound to %d of %d allowed CPUs local to %s" % (len(use), len(allowed), bus)
        return "no binding (%d local CPUs allowed of %d)" % (len(use), len(allowed))
    except Exception as e:  # sysfs layout differs / not permitted: run unbound
        return "no binding (%s)" % type(e).__name__


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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    numa = bind_to_gpu_numa_node(torch, dev)
    ctx = R.Context(dev)

    F = a.frames
    n_pairs = F - 1
    frames, Tgt = gen_scene.make_sweep(a.seed + 1000 * rank, F)
    guess = guess_matrix()
    icp = R.icp_params(**forced_kw(a.iters))
    ndt = R.ndt_params()
    coarse = R.COARSE_NDT if a.coarse == "ndt" else R.COARSE_ICP
    src_idx = np.arange(1, F, dtype=np.int32)
    tgt_idx = np.arange(0, F - 1, dtype=np.int32)

    # pinned host frames in pcl::PointXYZRGB layout (what the reference holds in memory) + pinned result buffer
    h_in = ctx.pinned(F * NPX * 32)
    h_in.view(R.PCL32)[:] = np.concatenate([R.to_pcl32(f) for f in frames])
    h_out = ctx.pinned(n_pairs * NPX * 32)
    counts = np.full(F, NPX, np.int32)
    d_frames = ctx.cloud(F, NPX)
    d_out = ctx.cloud(n_pairs, NPX)
    out_counts = np.zeros(n_pairs, np.int32)
    L = R.lib()
    import ctypes as C

    def upload():
        d_frames.upload_raw(h_in.ctypes.data_as(C.c_void_p), counts, W, H, R.LAYOUT_PCL32)

    def step():
        return R.register_pairs(ctx, d_frames, src_idx, tgt_idx, coarse, icp=icp, ndt=ndt, guess=guess, out_transformed=d_out)

    def download():
        ctx.check(L.rspcl_cloud_download(ctx.h, d_out.h, h_out.ctypes.data_as(C.c_void_p), R.LAYOUT_PCL32,
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
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches()
    ctx.timer_start()
    for _ in range(a.steps):
        step()
    ms = ctx.timer_stop()
    l1 = ctx.launches()
    barrier()
    ms = max_over_ranks(ms)
    value = world * n_pairs * a.steps / (ms / 1e3)

    # ---- end to end through the C ABI with host buffers: the sweep is cut into chunks of pairs that are pipelined over
    # several contexts (one stream + one host thread each), so chunk c+1's H2D, chunk c's kernels and chunk c-1's D2H
    # overlap (PCIe is full duplex).  Every frame still crosses PCIe (a chunk re-uploads its one boundary frame).
    from concurrent.futures import ThreadPoolExecutor
    n_ctx = a.e2e_contexts
    n_chunks = max(d for d in range(1, a.e2e_chunks + 1) if n_pairs % d == 0)  # equal chunks keep the batch shape fixed
    cp = n_pairs // n_chunks
    chunks = [(c * cp, (c + 1) * cp) for c in range(n_chunks)]
    workers = []
    for w in range(n_ctx):
        cw = ctx if w == 0 else R.Context(dev)
        workers.append({"ctx": cw, "frames": cw.cloud(cp + 1, NPX), "out": cw.cloud(cp, NPX), "oc": np.zeros(cp, np.int32)})
    in_rows = h_in.view(R.PCL32).reshape(F, NPX)
    out_rows = h_out.view(R.PCL32).reshape(n_pairs, NPX)
    cnts = np.full(cp + 1, NPX, np.int32)
    si = np.arange(1, cp + 1, dtype=np.int32)
    ti = np.arange(0, cp, dtype=np.int32)

    def run_chunks(w, steps, timed):
        """Worker w streams its chunks for `steps` consecutive steps (no barrier between steps: the e2e timed region
        is one continuous pipeline of steps x chunks, bracketed once on both sides)."""
        W_ = workers[w]
        cw = W_["ctx"]
        if timed:
            cw.timer_start()
        for st_i in range(steps):
            for ci in range(w, n_chunks, n_ctx):
                lo, hi = chunks[ci]
                # frames lo .. hi: pair i registers frame i+1 onto frame i
                # one transfer per direction at a time: the contexts fall into a staggered pipeline (A computes while B
                # uploads and C downloads) instead of moving in lockstep and sharing each PCIe direction
                t0 = time.perf_counter()
                with h2d_lock:
                    t1 = time.perf_counter()
                    W_["frames"].upload_raw(in_rows[lo:hi + 1].ctypes.data_as(C.c_void_p), cnts, W, H, R.LAYOUT_PCL32)
                    if a.e2e_lock == "sync":
                        cw.sync()
                t2 = time.perf_counter()
                R.register_pairs(cw, W_["frames"], si, ti, coarse, icp=icp, ndt=ndt, guess=guess, out_transformed=W_["out"])
                t3 = time.perf_counter()
                with d2h_lock:
                    t4 = time.perf_counter()
                    cw.check(L.rspcl_cloud_download(cw.h, W_["out"].h, out_rows[lo:hi].ctypes.data_as(C.c_void_p),
                                                    R.LAYOUT_PCL32, C.c_longlong(cp * NPX), W_["oc"].ctypes.data_as(C.c_void_p)))
                t5 = time.perf_counter()
                if timed and a.e2e_trace:
                    trace.append((w, st_i, ci, t0, t1, t2, t3, t4, t5))
        if timed:
            cw.timer_mark()

    import contextlib
    h2d_lock, d2h_lock = (contextlib.nullcontext(), contextlib.nullcontext()) if a.e2e_lock == "none" else (threading.Lock(), threading.Lock())
    trace = []
    pool = ThreadPoolExecutor(n_ctx)
    list(pool.map(lambda w: run_chunks(w, max(1, a.warmup - 1), False), range(n_ctx)))
    barrier()
    list(pool.map(lambda w: run_chunks(w, a.steps, True), range(n_ctx)))
    ms_e2e = max_over_ranks(R.timer_span([w["ctx"] for w in workers]))
    clocks = sampler.stop() if rank == 0 else None  # sampled over both timed regions (device-resident and e2e)
    if a.e2e_trace and trace:
        tz = min(t[3] for t in trace)
        for t in sorted(trace, key=lambda t: t[3]):
            sys.stderr.write("ctx %d step %d chunk %d: wait_h2d %.2f upload %.2f compute %.2f wait_d2h %.2f download %.2f  [start %.2f end %.2f ms]\n" % (
                t[0], t[1], t[2], 1e3 * (t[4] - t[3]), 1e3 * (t[5] - t[4]), 1e3 * (t[6] - t[5]), 1e3 * (t[7] - t[6]),
                1e3 * (t[8] - t[7]), 1e3 * (t[3] - tz), 1e3 * (t[8] - tz)))
    barrier()
    e2e = world * n_pairs * a.steps / (ms_e2e / 1e3)
    # the pipelined path must reproduce the single-context result
    chk = R.from_pcl32(out_rows[0])
    download()
    ref0 = R.from_pcl32(h_out.view(R.PCL32).reshape(n_pairs, NPX)[0])
    # (bit equality is not guaranteed: the cluster size, hence the fp64 summation order, depends on the batch size)
    dmax = max(float(np.abs(chk[a_] - ref0[a_]).max()) for a_ in "xyz")
    assert dmax < 1e-5 and np.array_equal(chk["rgba"], ref0["rgba"]), "pipelined e2e result differs from the single-context result"

    # ---- roofline of the dominant kernel, from per-launch CUDA events in one extra profiled step
    ctx.profile_reset()
    ctx.profile(True)
    ctx.timer_start()
    step()
    ms_prof = ctx.timer_stop()
    ctx.profile(False)
    kern = {k: ctx.profile_get(k) for k in ("k_icp_persist", "k_icp_step", "k_icp_stream", "k_icp_rescan", "k_icp_solve", "grid_build", "k_canny_nms",
                                            "edge_hysteresis_compact", "k_approx_voxel", "k_transform2", "k_ndt_eval",
                                            "ndt_voxel_build")}
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
                "note": "persistent kernel: the target grid lives in shared memory and the working cloud in L2, so DRAM traffic "
                        "is ~2 % of the algorithmic bytes; the kernel is bound by the serial chain of one ICP iteration (stream, "
                        "re-query, reduce, cluster exchange, solve) on the slowest pair of the batch, not by HBM (SURVEY H3). "
                        "The HBM-bound regime (one 25 M-point pair, k_icp_stream: 4.5 TB/s of DRAM traffic, 27 % in algorithmic "
                        "bytes) is in profiles/r01_final_summary.md"}

    # ---- side measurement (rank 0, N=1): BASELINE configs[2], edge-based NDT on one 1280x720 pair, 0.05 m voxels
    ndt_leg = None
    if rank == 0 and world == 1 and not a.no_ndt:
        W2, H2 = 1280, 720
        fr2, T2 = gen_scene.make_sweep(3, 2, W2, H2, noise_scale=0.2)
        d2 = ctx.upload(list(fr2), W2, H2)
        v2 = R.voxel_approx(ctx, R.edge_extract(ctx, d2)).download()
        s2, t2 = ctx.upload([v2[1]]), ctx.upload([v2[0]])
        p2 = R.ndt_params(resolution=0.05)
        R.ndt_align(ctx, s2, t2, p2, guess=guess, want_aligned=False)
        ctx.profile_reset()
        ctx.profile(True)
        ctx.timer_start()
        r2, _ = R.ndt_align(ctx, s2, t2, p2, guess=guess, want_aligned=False)
        ms2 = ctx.timer_stop()
        ctx.profile(False)
        ke, kc = ctx.profile_get("k_ndt_eval"), ctx.profile_get("k_ndt_control")
        ang2, tr2 = pose_err(r2[0]["T"], gen_scene.pairwise_gt(T2, 1))
        ndt_leg = {"workload": "configs[2]: edge-based NDT, one 1280x720 pair (921,600 points per frame), 0.05 m voxels, step 0.1, eps 0.01",
                   "edge_points_src_tgt": [int(len(v2[1])), int(len(v2[0]))], "iterations": r2[0]["iterations"],
                   "derivative_evals": r2[0]["n_derivative_evals"], "converged": bool(r2[0]["converged"]),
                   "ms_align": ms2, "ms_per_ndt_iteration": ms2 / max(r2[0]["iterations"], 1),
                   "ms_per_derivative_eval": ke["ms"] / max(ke["launches"], 1),
                   "ms_per_control_step": kc["ms"] / max(kc["launches"], 1),
                   "err_vs_ground_truth_rad_m": [ang2, tr2]}

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
        out = {
            "metric": "frame-pair registrations/sec @640x480", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, F),
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": (n_pairs + len(chunks)) * NPX * 32,
                    "d2h_bytes_per_step": n_pairs * NPX * 32 + n_pairs * 160, "ms_per_step": ms_e2e / a.steps,
                    "pipeline": "%d chunks over %d contexts (streams), pinned host buffers" % (len(chunks), n_ctx),
                    "host_numa": numa},
            "gpu_launches": int(l1 - l0),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "kernels_ms_per_step": {k: v["ms"] for k, v in kern.items() if v["launches"]},
            "ndt": ndt_leg,
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
