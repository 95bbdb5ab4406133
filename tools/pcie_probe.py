#!/usr/bin/env python
"""Measured host<->device bandwidth of this box (pinned memory, cudaMemcpyAsync, CUDA events): H2D alone, D2H alone and
both directions at once on two streams.  The e2e figure of bench.py moves 32 B per point in each direction, so these
are the denominators of its PCIe roofline.  Prints one JSON line."""
import json

import torch


def main():
    n = 512 << 20
    h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=5):
        best = 1e30
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            s1.synchronize()
            s2.synchronize()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    for f in (h2d, d2h, both):
        f()
    torch.cuda.synchronize()
    import time

    def wall(fn, reps=5):
        best = 1e30
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best

    gb = n / 1e9
    out = {"bytes": n, "h2d_GBps": gb / wall(h2d), "d2h_GBps": gb / wall(d2h)}
    t = wall(both)
    out["bidir_each_GBps"] = gb / t
    out["bidir_total_GBps"] = 2 * gb / t
    print(json.dumps(out))


if __name__ == "__main__":
    main()
