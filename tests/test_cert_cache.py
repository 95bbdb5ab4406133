"""The certified nearest-neighbour cache of the ICP kernels (icp_persist.cuh phase A/B, icp.cu k_icp_stream/k_icp_rescan)
restated in numpy float32 and checked against a brute-force search at every iteration: the cached decision must be the
exact nearest neighbour inside the gate (lowest index on ties) for every point, whatever the motion, and unmatched points
must stay unmatched exactly as long as the brute force says so.  CPU only; the device code is covered by the GPU parity
tests (first-iteration indices, n_corr and transforms against the oracle)."""
import numpy as np

F = np.float32


def dist2(p, t):  # FLANN L2_Simple order, float32
    d = (p[0] - t[:, 0]).astype(F)
    r = (d * d).astype(F)
    d = (p[1] - t[:, 1]).astype(F)
    r = (r + (d * d).astype(F)).astype(F)
    d = (p[2] - t[:, 2]).astype(F)
    return (r + (d * d).astype(F)).astype(F)


def cells_of_ball(p, rr, inv_cs):
    lo = np.floor((p - rr).astype(F) * inv_cs).astype(np.int64)
    hi = np.floor((p + rr).astype(F) * inv_cs).astype(np.int64)
    return lo, hi


def brute(p, tgt, gate2):
    d = dist2(p, tgt)
    j = int(np.argmin(d))  # argmin returns the lowest index on ties
    return (j, d[j]) if not (np.float64(d[j]) > gate2) else (-1, d[j])


def run(seed, n_t, n_s, gate, iters, motion0):
    rng = np.random.default_rng(seed)
    tgt = rng.uniform(0, 1, (n_t, 3)).astype(F)
    # a few exact duplicates and mirrored pairs so that exact distance ties do occur
    tgt[1] = tgt[0]
    src = (tgt[rng.integers(0, n_t, n_s)] + rng.normal(0, gate * 0.6, (n_s, 3))).astype(F)
    src[:20] = rng.uniform(1.5, 2.0, (20, 3)).astype(F)  # far from everything: must stay unmatched
    gate2 = np.float64(gate) ** 2
    r = F(gate * 1.01)
    cs = F(gate * 4.1)
    inv_cs = F(1.0) / cs
    rmax = F(0.485) / inv_cs
    slack = F(0.5) * r
    tcell = np.floor(tgt * inv_cs).astype(np.int64)
    cached = np.full(n_s, -1)
    lb = np.zeros(n_s, F)
    n_rescan = []
    for it in range(iters):
        ang = motion0 * 0.7 ** it
        c, s = np.cos(ang), np.sin(ang)
        R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        tr = rng.normal(0, motion0 * 0.7 ** it, 3)
        new = (src.astype(np.float64) @ R.T + tr).astype(F)
        moved = np.sqrt(((new - src).astype(F) ** 2).sum(1, dtype=F)).astype(F)
        lb = (lb - (moved * F(1.00001) + F(1e-9))).astype(F)
        src = new
        resc = 0
        for i in range(n_s):
            p = src[i]
            kp = cached[i]
            if kp >= 0:
                bd = dist2(p, tgt[kp:kp + 1])[0]
                s1 = F(np.sqrt(bd) * F(1.0001) + F(1e-7))
                valid = s1 < lb[i]
            else:
                valid = lb[i] > r
            if not valid:
                resc += 1
                rr = rmax
                if kp >= 0 and s1 <= rmax:
                    rr = min(rmax, F(s1 + slack))
                else:
                    kp = -1
                lo, hi = cells_of_ball(p, rr, inv_cs)
                inside = np.all((tcell >= lo) & (tcell <= hi), axis=1)
                idx = np.nonzero(inside)[0]
                if kp >= 0 and not inside[kp]:
                    idx = np.sort(np.append(idx, kp))  # the incumbent is merged even if its cell is outside the ball
                if len(idx):
                    d = dist2(p, tgt[idx])
                    o = np.lexsort((idx, d))  # distance, then lowest index
                    kp, bd = int(idx[o[0]]), d[o[0]]
                    d2nd = d[o[1]] if len(idx) > 1 else F(np.inf)
                else:
                    kp, bd, d2nd = -1, F(np.inf), F(np.inf)
                edge = F(rr - F(2e-7) * (np.abs(p).sum(dtype=F) + rr))
                lb[i] = F(min(F(np.sqrt(d2nd)), edge) * F(0.9999))
                cached[i] = kp
            decided = kp if (kp >= 0 and not (np.float64(bd) > gate2)) else -1
            want, _ = brute(p, tgt, gate2)
            assert decided == want, (seed, it, i, decided, want, valid)
        n_rescan.append(resc)
    return n_rescan


def test_cached_decision_equals_brute_force_every_iteration():
    resc = run(seed=3, n_t=2500, n_s=500, gate=0.02, iters=14, motion0=0.01)
    assert resc[0] == 500  # nothing cached at the start
    assert resc[-1] < 150  # converged: only points with two nearly equidistant targets are looked at again (dense random target)


def test_large_motion_and_dense_target():
    # motion of the order of the gate between iterations and ~10 target points inside every gate ball
    run(seed=5, n_t=6000, n_s=250, gate=0.06, iters=8, motion0=0.03)
