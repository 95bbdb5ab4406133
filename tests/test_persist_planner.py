"""Host logic of the persistent ICP kernel's launch plan (icp.cu plan_clusters), on the CPU: every pair gets a cluster of
1..8 CTAs such that one wave fits the SM budget, the slowest slice is as short as the budget allows, and slices stay
register-resident whenever that is possible at all."""
import ctypes as C

import numpy as np

import rspcl_b200 as R


def plan(counts, budget, weights=None):
    L = R.lib()
    L.rspcl_debug_plan_clusters.restype = C.c_int
    cnt = np.ascontiguousarray(counts, np.int32)
    out = np.zeros(len(cnt), np.int32)
    w = np.ascontiguousarray(weights, np.float64) if weights is not None else None
    chunk = L.rspcl_debug_plan_clusters(cnt.ctypes.data_as(C.c_void_p), len(cnt), C.c_double(budget),
                                        w.ctypes.data_as(C.c_void_p) if w is not None else None, out.ctypes.data_as(C.c_void_p))
    return out, chunk


def cost(n, c, chunk):
    ln = -(-n // c)
    return 2 * ln if ln > chunk else ln


def test_bench_sweep_distribution_fills_the_chip_with_resident_slices():
    # voxel-filtered edge counts of the 64 source frames of bench.py's default sweep (seed 2)
    counts = np.array([4500, 4566, 4590, 4591, 4640, 4647, 4670, 4671, 4703, 4716, 4737, 4741, 4767, 4776, 4820, 4842, 5040, 5063,
                       5123, 5132, 5163, 6211, 6234, 6240, 6245, 6250, 6348, 7267, 7352, 7376, 7426, 7449, 7708, 7767, 7834, 7857,
                       7918, 7919, 7954, 7957, 8010, 8059, 8175, 8513, 8703, 8729, 8750, 8778, 8866, 10232, 10432, 10458, 10460,
                       10594, 11056, 11141, 11381, 11382, 11581, 11593, 11654, 11778, 12048, 12090])
    cl, chunk = plan(counts, 146)
    assert cl.sum() <= 146 and cl.min() >= 1 and cl.max() <= 8
    assert all(-(-n // c) <= chunk for n, c in zip(counts, cl))       # nobody streams
    assert cl.sum() >= 140                                            # and the wave is (nearly) full
    # the plan is a minimax: no pair could take a CTA from another without raising the slowest slice
    worst = max(cost(n, c, chunk) for n, c in zip(counts, cl))
    for i in range(len(counts)):
        if cl[i] > 1:
            assert cost(counts[i], cl[i] - 1, chunk) >= worst * 0.999 or cl.sum() < 146


def test_single_pair_small_clouds_and_oversubscription():
    cl, chunk = plan([7000], 146)
    assert cl[0] == 8                                # one pair: as many CTAs as a cluster allows
    cl, _ = plan([150] * 100, 146)
    assert (cl == 1).all()                           # tiny clouds: more CTAs would only add exchange latency
    cl, _ = plan([9000] * 400, 146)
    assert (cl == 1).all()                           # more pairs than SMs: one CTA each (the caller cuts waves first)
    cl, chunk = plan([60000, 500], 146)
    assert cl[0] == 8 and cl[1] == 1                 # a slice above the register capacity still gets the maximum


def test_cluster_weights_from_the_occupancy_query_are_respected():
    counts = [11000] * 40
    w = [1, 2, 3.08, 4.48, 5, 6.2, 7, 8.7]           # what clusters of 3 / 4 / ... effectively occupy on 148 SMs
    cl, chunk = plan(counts, 146, w)
    assert sum(w[c - 1] for c in cl) <= 146 + 1e-9
    cl_unw, _ = plan(counts, 146)
    assert cl.sum() <= cl_unw.sum()


def test_optimistic_budget_keeps_the_bench_plan_on_the_chip():
    """The planner starts at SMs + 2 (icp.cu): with the measured occupancy weights (clusters of 3 charged 3.2 SMs) the plan for
    the bench sweep must still be at most 148 CTAs -- what ran resident in one wave on the B200 (1.25 ms against 1.38 ms at
    SMs - 2); the start-time feedback covers chips where it does not."""
    counts = np.array([4500, 4566, 4590, 4591, 4640, 4647, 4670, 4671, 4703, 4716, 4737, 4741, 4767, 4776, 4820, 4842, 5040, 5063,
                       5123, 5132, 5163, 6211, 6234, 6240, 6245, 6250, 6348, 7267, 7352, 7376, 7426, 7449, 7708, 7767, 7834, 7857,
                       7918, 7919, 7954, 7957, 8010, 8059, 8175, 8513, 8703, 8729, 8750, 8778, 8866, 10232, 10432, 10458, 10460,
                       10594, 11056, 11141, 11381, 11382, 11581, 11593, 11654, 11778, 12048, 12090])
    w = [1, 2, 148 / 46.0, 148 / 34.0, 148 / 27.0, 148 / 22.0, 148 / 18.0, 148 / 16.0]
    cl, chunk = plan(counts, 150, w)
    assert cl.sum() <= 148
    assert sum(w[c - 1] for c in cl) <= 150 + 1e-9
    worst150 = max(cost(n, c, chunk) for n, c in zip(counts, cl))
    cl146, _ = plan(counts, 146, w)
    worst146 = max(cost(n, c, chunk) for n, c in zip(counts, cl146))
    assert worst150 <= worst146 and cl.sum() >= cl146.sum()
