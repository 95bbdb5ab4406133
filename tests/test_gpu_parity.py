"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): edge masks, voxel keys/outputs and NN correspondence indices bit-exact; final 4x4
transforms within 1e-4 rad / 1e-4 m; fitness within 1e-5 relative."""
import numpy as np
import pytest

import gen_scene
import orc
import rspcl_b200 as R

pytestmark = pytest.mark.gpu
F = np.float32
W, H = 640, 480


@pytest.fixture(scope="module")
def ctx():
    c = R.Context(0)
    yield c
    c.close()


def rand_cloud(rng, n, scale=1.0):
    p = np.zeros(n, R.POINT)
    p["x"], p["y"], p["z"] = (rng.uniform(-scale, scale, (3, n))).astype(F)
    p["rgba"] = rng.integers(0, 2**24, n).astype(np.uint32) | np.uint32(0xFF000000)
    return p


def rigid(rng, ang=0.2, tr=0.1):
    T = np.eye(4)
    T[:3, :3] = gen_scene.rot_axis(rng.normal(size=3), rng.uniform(-ang, ang))
    T[:3, 3] = rng.uniform(-tr, tr, 3)
    return T


def pose_err(A, B):
    D = np.linalg.inv(np.asarray(A, np.float64)) @ np.asarray(B, np.float64)
    # atan2 of the skew part: well conditioned for tiny angles (acos(1 - 1e-8) would already read 1.4e-4 rad)
    sk = np.linalg.norm([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]]) / 2
    return np.arctan2(sk, (np.trace(D[:3, :3]) - 1) / 2), np.linalg.norm(D[:3, 3])


def forced(n_iter):
    kw = dict(max_iterations=n_iter, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300, mse_threshold_absolute=-1.0)
    return R.icp_params(**kw), orc.icp_params(**kw)


# ------------------------------------------------------------------ clouds, layouts, transform, concat, crop
def test_upload_download_layouts_ragged_and_empty(ctx):
    rng = np.random.default_rng(1)
    segs = [rand_cloud(rng, n) for n in (0, 1, 257, 5000, 33)]
    c = ctx.upload(segs)
    assert c.counts().tolist() == [0, 1, 257, 5000, 33]
    for a, b in zip(c.download(), segs):
        assert np.array_equal(a, b)
    c32 = ctx.upload([R.to_pcl32(s) for s in segs], layout=R.LAYOUT_PCL32)
    for a, b in zip(c32.download(), segs):
        assert np.array_equal(a, b)
    back32 = c32.download(R.LAYOUT_PCL32)
    for a, b in zip(back32, segs):
        assert np.array_equal(R.from_pcl32(a), b) and (a["w"] == 1.0).all()


def test_transform_bit_exact_and_in_place(ctx):
    rng = np.random.default_rng(2)
    segs = [rand_cloud(rng, n, 3.0) for n in (1000, 7, 40000)]
    segs[0]["x"][5] = np.nan  # non-finite points are copied, not moved
    Ts = np.stack([rigid(rng) for _ in segs]).astype(F)
    c = ctx.upload(segs)
    out = R.transform(ctx, c, Ts).download()
    for s, T, o in zip(segs, Ts, out):
        assert np.array_equal(o.view(np.uint32), orc.transform(s, T).view(np.uint32))
    R.transform(ctx, c, Ts[0], out=c)  # broadcast + in place (icp:117)
    for s, o in zip(segs, c.download()):
        assert np.array_equal(o.view(np.uint32), orc.transform(s, Ts[0]).view(np.uint32))


def test_concat_order(ctx):
    rng = np.random.default_rng(3)
    a = [rand_cloud(rng, n) for n in (10, 0, 300)]
    b = [rand_cloud(rng, n) for n in (5, 4, 0)]
    out = R.concat(ctx, ctx.upload(a), ctx.upload(b)).download()
    for x, y, o in zip(a, b, out):
        assert np.array_equal(o, np.concatenate([x, y]))


def test_crop35(ctx, pair2):
    fr, _ = pair2
    c = ctx.upload([fr[0], fr[1]], W, H)
    out = R.crop35(ctx, c)
    assert out.dims() == (384, 288)
    for k, o in enumerate(out.download()):
        exp, ow, oh = orc.crop35(fr[k], W, H)
        assert np.array_equal(o, exp)


# ------------------------------------------------------------------ edges
def test_edge_mask_and_cloud_bit_exact_scene(ctx, sweep3):
    fr, _ = sweep3
    c = ctx.upload(list(fr), W, H, layout=R.LAYOUT_PCD16)
    edges, mask = R.edge_extract(ctx, c, want_mask=True)
    got = edges.download()
    for k in range(len(fr)):
        m, near = orc.canny(fr[k], W, H)
        assert near == 0, "oracle flags %d pixels within 1e-3 deg of a direction bin edge" % near
        assert np.array_equal(mask[k], m), "frame %d: %d mask pixels differ" % (k, (mask[k] != m).sum())
        e, _ = orc.extract_edges(fr[k], W, H)
        assert np.array_equal(got[k], e)
        assert len(e) > 3000


@pytest.mark.parametrize("w,h", [(64, 48), (37, 23), (130, 17), (3, 3), (640, 480)])
def test_edge_mask_random_images(ctx, w, h):
    rng = np.random.default_rng(w * 100 + h)
    frames = []
    for _ in range(3):
        img = np.zeros((h, w), np.uint32)
        for _ in range(14):
            r0, c0 = rng.integers(0, h), rng.integers(0, w)
            img[r0:r0 + rng.integers(1, h + 1), c0:c0 + rng.integers(1, w + 1)] = rng.integers(0, 256)
        img = np.minimum(img + rng.integers(0, 6, (h, w)).astype(np.uint32), 255).astype(np.uint32)
        cl = rand_cloud(rng, w * h)
        cl["rgba"] = ((img << 16) | (img << 8) | img).ravel() | np.uint32(0xFF000000)
        frames.append(cl)
    c = ctx.upload(frames, w, h)
    edges, mask = R.edge_extract(ctx, c, want_mask=True)
    got = edges.download()
    for k, f in enumerate(frames):
        m, dbg = orc.canny(f, w, h, debug=True)
        if dbg["near_bin_edge"]:
            continue
        assert np.array_equal(mask[k], m), (k, (mask[k] != m).sum())
        assert np.array_equal(got[k], f[np.flatnonzero(m.ravel())])


def test_edge_mask_without_a_gray_plane_and_on_a_reused_handle(ctx, sweep3):
    """ADVICE r1 (stale gray plane) and the k_canny_nms<FROM_PTS> path: a batch whose cached (r+g+b)/3 plane is stale --
    explicitly invalidated, or because a device-side operation rewrote the colours of a handle that had been organized
    before -- must give the masks of its CURRENT colours."""
    fr, _ = sweep3
    c = ctx.upload(list(fr), W, H)
    c.invalidate_gray()
    edges, mask = R.edge_extract(ctx, c, want_mask=True)
    got = edges.download()
    for k in range(len(fr)):
        m, _ = orc.canny(fr[k], W, H)
        assert np.array_equal(mask[k], m), "frame %d: %d mask pixels differ" % (k, (mask[k] != m).sum())
        assert np.array_equal(got[k], orc.extract_edges(fr[k], W, H)[0])
    # reuse: `dst` was uploaded with frame 1 (plane valid for frame 1), then overwritten on the device with frame 0
    src0 = ctx.upload([fr[0]], W, H)
    dst = ctx.upload([fr[1]], W, H)
    R.edge_extract(ctx, dst)  # (uses dst's plane once, like a previous registration would have)
    R.transform(ctx, src0, np.eye(4), out=dst)
    _, mask2 = R.edge_extract(ctx, dst, want_mask=True)
    m0, _ = orc.canny(fr[0], W, H)
    assert np.array_equal(mask2[0], m0), "%d mask pixels differ (stale gray plane?)" % (mask2[0] != m0).sum()


@pytest.mark.parametrize("w,h", [(37, 23), (130, 17), (3, 3), (64, 48)])
def test_edge_mask_random_images_from_points(ctx, w, h):
    rng = np.random.default_rng(w * 131 + h)
    frames = []
    for _ in range(2):
        img = np.zeros((h, w), np.uint32)
        for _ in range(14):
            r0, c0 = rng.integers(0, h), rng.integers(0, w)
            img[r0:r0 + rng.integers(1, h + 1), c0:c0 + rng.integers(1, w + 1)] = rng.integers(0, 256)
        img = np.minimum(img + rng.integers(0, 6, (h, w)).astype(np.uint32), 255).astype(np.uint32)
        cl = rand_cloud(rng, w * h)
        # three different channels: the plane is (r + g + b) / 3 in integer arithmetic
        cl["rgba"] = ((img << 16) | (np.minimum(img + 1, 255) << 8) | np.maximum(img, 2) - 2).ravel() | np.uint32(0xFF000000)
        frames.append(cl)
    c = ctx.upload(frames, w, h)
    c.invalidate_gray()
    _, mask = R.edge_extract(ctx, c, want_mask=True)
    for k, f in enumerate(frames):
        m, dbg = orc.canny(f, w, h, debug=True)
        if dbg["near_bin_edge"]:
            continue
        assert np.array_equal(mask[k], m), (k, (mask[k] != m).sum())


def test_edge_thresholds_and_pcl32_input(ctx, pair2):
    fr, _ = pair2
    c = ctx.upload([R.to_pcl32(fr[0])], W, H, layout=R.LAYOUT_PCL32)
    for lo, hi in ((40.0, 100.0), (20.0, 60.0), (90.0, 200.0)):
        _, mask = R.edge_extract(ctx, c, lo, hi, want_mask=True)
        m, near = orc.canny(fr[0], W, H, lo, hi)
        assert np.array_equal(mask[0], m)


# ------------------------------------------------------------------ ApproximateVoxelGrid
def test_voxel_keys_and_output_bit_exact(ctx, sweep3):
    fr, _ = sweep3
    edges = [orc.extract_edges(f, W, H)[0] for f in fr]
    c = ctx.upload(edges)
    ijk, slot = R.voxel_keys(ctx, c)
    eijk = np.concatenate([orc.voxel_keys(e)[0] for e in edges])
    eslot = np.concatenate([orc.voxel_keys(e)[1] for e in edges])
    assert np.array_equal(ijk, eijk) and np.array_equal(slot, eslot)
    out = R.voxel_approx(ctx, c).download()
    for e, o in zip(edges, out):
        exp = orc.approx_voxel(e)
        assert len(o) == len(exp)
        assert np.array_equal(o.view(np.uint32), exp.view(np.uint32))
    # in place (icp:59-60)
    R.voxel_approx(ctx, c, in_place=True)
    for e, o in zip(edges, c.download()):
        assert np.array_equal(o.view(np.uint32), orc.approx_voxel(e).view(np.uint32))


def test_voxel_edge_cases(ctx):
    rng = np.random.default_rng(5)
    dense = rand_cloud(rng, 20000, 0.05)          # many collisions and revisits
    dense[:8000] = np.sort(dense[:8000], order=["z", "y", "x"])
    big_leaf = rand_cloud(rng, 50000, 2.0)        # IncrementalICP: default 1 m leaf, long runs
    segs = [dense, np.zeros(0, R.POINT), dense[:1], big_leaf]
    out = R.voxel_approx(ctx, ctx.upload(segs)).download()
    for s, o in zip(segs[:3], out[:3]):
        assert np.array_equal(o.view(np.uint32), orc.approx_voxel(s).view(np.uint32))
    out1 = R.voxel_approx(ctx, ctx.upload([big_leaf]), leaf=(1.0, 1.0, 1.0)).download()[0]
    assert np.array_equal(out1.view(np.uint32), orc.approx_voxel(big_leaf, (1.0, 1.0, 1.0)).view(np.uint32))


# ------------------------------------------------------------------ NN, fitness
def test_nearest_and_fitness(ctx):
    rng = np.random.default_rng(7)
    tg = [rand_cloud(rng, 3000), rand_cloud(rng, 17)]
    qs = [rand_cloud(rng, 1200), rand_cloud(rng, 600)]
    tg[0][100:120] = tg[0][50]
    qs[0][:20] = tg[0][100:120]
    ct, cq = ctx.upload(tg), ctx.upload(qs)
    idx, d2 = R.nearest(ctx, cq, ct)
    ei = np.concatenate([orc.nn(t, q, "brute")[0] for t, q in zip(tg, qs)])
    ed = np.concatenate([orc.nn(t, q, "brute")[1] for t, q in zip(tg, qs)])
    assert np.array_equal(idx, ei) and np.array_equal(d2, ed)
    fit = R.fitness(ctx, cq, ct)
    for k in range(2):
        e = orc.fitness(qs[k], tg[k])
        assert abs(fit[k] - e) <= 1e-5 * e
    thr = float(np.median(ed))
    fit = R.fitness(ctx, cq, ct, thr)
    for k in range(2):
        e = orc.fitness(qs[k], tg[k], thr)
        assert abs(fit[k] - e) <= 1e-5 * e
    assert (R.fitness(ctx, cq, ct, -1.0) == np.finfo(np.float64).max).all()


# ------------------------------------------------------------------ ICP
def edge_pair(fr):
    e0 = orc.approx_voxel(orc.extract_edges(fr[0], W, H)[0])
    e1 = orc.approx_voxel(orc.extract_edges(fr[1], W, H)[0])
    return e0, e1


def test_icp_first_iteration_correspondences_bit_exact(ctx, pair2):
    fr, _ = pair2
    tgt, src = edge_pair(fr)
    guess = np.eye(4)
    guess[:3, :3] = gen_scene.rot_y(-0.523599)
    res, aligned, fc = R.icp_align(ctx, ctx.upload([src]), ctx.upload([tgt]), R.icp_params(), guess=guess, want_first_corr=True)
    o = orc.icp_align(src, tgt, orc.icp_params(), guess=guess, want_first_corr=True)
    assert np.array_equal(fc, o["first_corr"])
    assert (fc >= 0).sum() > 500
    r = res[0]
    assert r["converged"] == o["converged"] and r["iterations"] == o["iterations"] == 1 and r["state"] == o["state"] == 2
    assert r["n_corr"] == o["n_corr"]
    ang, tr = pose_err(r["T"], o["T"])
    assert ang < 1e-4 and tr < 1e-4
    assert abs(r["mse"] - o["mse"]) <= 1e-5 * o["mse"]
    a = aligned.download()[0]
    assert np.array_equal(a.view(np.uint32), orc.transform(src, r["T"]).view(np.uint32))


def test_icp_50_iterations_config2(ctx, pair2):
    fr, Tgt = pair2
    tgt, src = edge_pair(fr)
    guess = np.eye(4)
    guess[:3, :3] = gen_scene.rot_y(-0.523599)
    gp, op = forced(50)
    res, _, _ = R.icp_align(ctx, ctx.upload([src]), ctx.upload([tgt]), gp, guess=guess, want_aligned=False)
    o = orc.icp_align(src, tgt, op, guess=guess)
    r = res[0]
    assert r["iterations"] == o["iterations"] == 50 and r["state"] == o["state"] == 1 and r["converged"]
    ang, tr = pose_err(r["T"], o["T"])
    assert ang < 1e-4 and tr < 1e-4, (ang, tr)
    assert abs(r["n_corr"] - o["n_corr"]) <= max(3, o["n_corr"] // 200)
    gt = gen_scene.pairwise_gt(Tgt, 1)
    ang, tr = pose_err(r["T"], gt)
    assert ang < 2e-3 and tr < 5e-3


def test_icp_batch_of_pairs_matches_single(ctx):
    rng = np.random.default_rng(11)
    srcs, tgts, Ts = [], [], []
    for k in range(6):
        t = rand_cloud(rng, 1500 + 300 * k, 0.4)
        T = rigid(rng, 0.003, 0.002)
        srcs.append(orc.transform(t[:: 2 + (k % 2)], np.linalg.inv(T)))
        tgts.append(t)
        Ts.append(T)
    gp, op = forced(20)
    res, aligned, _ = R.icp_align(ctx, ctx.upload(srcs), ctx.upload(tgts), gp)
    al = aligned.download()
    for k in range(6):
        o = orc.icp_align(srcs[k], tgts[k], op)
        ang, tr = pose_err(res[k]["T"], o["T"])
        assert ang < 1e-4 and tr < 1e-4
        assert res[k]["iterations"] == 20
        ang, tr = pose_err(res[k]["T"], Ts[k])
        assert ang < 1e-3 and tr < 1e-3  # ground truth: 20 point-to-point iterations are not fully converged
        assert np.array_equal(al[k].view(np.uint32), orc.transform(srcs[k], res[k]["T"]).view(np.uint32))


def test_icp_shared_target_mixed_convergence_and_no_correspondences(ctx):
    rng = np.random.default_rng(13)
    tgt = rand_cloud(rng, 4000, 0.5)
    s_ok = orc.transform(tgt[::2], np.linalg.inv(rigid(rng, 0.002, 0.001)))
    s_far = tgt[::3].copy()
    s_far["x"] += 10.0
    s_tiny = tgt[:2].copy()
    srcs = [s_ok, s_far, s_tiny, np.zeros(0, R.POINT)]
    res, _, _ = R.icp_align(ctx, ctx.upload(srcs), ctx.upload([tgt]), R.icp_params(), want_aligned=False)
    for k, s in enumerate(srcs):
        o = orc.icp_align(s, tgt, orc.icp_params())
        assert res[k]["converged"] == o["converged"] and res[k]["state"] == o["state"] and res[k]["iterations"] == o["iterations"], k
        assert res[k]["n_corr"] == o["n_corr"]
    assert res[0]["converged"] and not res[1]["converged"] and res[1]["state"] == 5


def test_icp_unbounded_default_distance_uses_exact_nn(ctx):
    rng = np.random.default_rng(17)
    tgt = rand_cloud(rng, 2500, 0.5)
    T = rigid(rng, 0.01, 0.01)
    src = orc.transform(tgt[::2], np.linalg.inv(T))
    kw = dict(max_iterations=15, max_corr_dist=float(np.sqrt(np.finfo(np.float64).max)), transformation_epsilon=-1.0,
              euclidean_fitness_epsilon=-1e300, mse_threshold_absolute=-1.0)
    res, _, fc = R.icp_align(ctx, ctx.upload([src]), ctx.upload([tgt]), R.icp_params(**kw), want_aligned=False, want_first_corr=True)
    o = orc.icp_align(src, tgt, orc.icp_params(**kw), want_first_corr=True)
    assert np.array_equal(fc, o["first_corr"])
    ang, tr = pose_err(res[0]["T"], o["T"])
    assert ang < 1e-4 and tr < 1e-4


def test_icp_prev_mse_persists_like_one_pcl_object(ctx):
    rng = np.random.default_rng(19)
    tgt = rand_cloud(rng, 3000, 0.4)
    src = orc.transform(tgt[::2], np.linalg.inv(rigid(rng, 0.002, 0.002)))
    kw = dict(max_iterations=40, transformation_epsilon=1e-14, euclidean_fitness_epsilon=1e-4)
    gp, op = R.icp_params(**kw), orc.icp_params(**kw)
    r1, _, _ = R.icp_align(ctx, ctx.upload([src]), ctx.upload([tgt]), gp, want_aligned=False)
    o1 = orc.icp_align(src, tgt, op)
    assert r1[0]["iterations"] == o1["iterations"] and r1[0]["state"] == o1["state"]
    r2, _, _ = R.icp_align(ctx, ctx.upload([src]), ctx.upload([tgt]), gp, prev_mse=r1[0]["prev_mse"], want_aligned=False)
    o2 = orc.icp_align(src, tgt, op, prev_mse=o1["prev_mse"])
    assert r2[0]["iterations"] == o2["iterations"] and r2[0]["state"] == o2["state"]


# ------------------------------------------------------------------ pairwise pipeline (configs 1/2 inputs)
def test_register_pairs_matches_oracle_pairwise(ctx, sweep3):
    fr, Tgt = sweep3
    frames = ctx.upload([R.to_pcl32(f) for f in fr], W, H, layout=R.LAYOUT_PCL32)
    guess = np.eye(4)
    guess[:3, :3] = gen_scene.rot_y(-0.523599)
    gp, op = forced(50)
    out = ctx.cloud(2, W * H)
    res = R.register_pairs(ctx, frames, [1, 2], [0, 1], R.COARSE_ICP, icp=gp, guess=guess, out_transformed=out)
    full = out.download()
    for i, (s, t) in enumerate(((1, 0), (2, 1))):
        o = orc.register_pair(fr[t], fr[s], W, H, "icp", icp=op, guess=guess, want_full=True)
        r = res[i]
        assert bool(r.converged) == o["converged"]
        assert r.coarse_iterations == 50 and r.fine_iterations == 50
        for key in ("T_coarse", "T_fine"):
            ang, tr = pose_err(R.c_to_mat(getattr(r, key)), o[key])
            assert ang < 1e-4 and tr < 1e-4, (i, key, ang, tr)
        exp = orc.transform(orc.transform(fr[s], R.c_to_mat(r.T_coarse)), R.c_to_mat(r.T_fine))
        assert np.array_equal(full[i].view(np.uint32), exp.view(np.uint32))
        d = np.abs(np.stack([full[i][a] - o["transformed_full"][a] for a in "xyz"]))
        assert d.max() < 5e-4
        Tgpu = R.c_to_mat(r.T_fine).astype(np.float64) @ R.c_to_mat(r.T_coarse).astype(np.float64)
        ang, tr = pose_err(Tgpu, gen_scene.pairwise_gt(Tgt, s))
        assert ang < 3e-3 and tr < 6e-3


def test_register_pairs_reference_literal_settings(ctx, sweep3):
    fr, _ = sweep3
    frames = ctx.upload(list(fr), W, H)
    guess = np.eye(4)
    guess[:3, :3] = gen_scene.rot_y(-0.523599)
    res = R.register_pairs(ctx, frames, [1, 2], [0, 1], R.COARSE_ICP, guess=guess)
    for i, (s, t) in enumerate(((1, 0), (2, 1))):
        o = orc.register_pair(fr[t], fr[s], W, H, "icp", guess=guess)
        assert res[i].coarse_iterations == 1 and res[i].fine_iterations == 1 and res[i].converged == 1
        for key in ("T_coarse", "T_fine"):
            ang, tr = pose_err(R.c_to_mat(getattr(res[i], key)), o[key])
            assert ang < 1e-4 and tr < 1e-4
        assert res[i].n_corr == o["fine"].n_corr


def test_frames_with_holes_and_nans(ctx, pair2):
    """librealsense emits (0,0,0) for missing depth and PCL clouds may carry NaNs: edges depend on rgb only,
    the voxel filter bins the zero points into one voxel, non-finite points never become correspondences."""
    fr, _ = pair2
    rng = np.random.default_rng(23)
    f = [fr[0].copy(), fr[1].copy()]
    for c in f:
        holes = rng.random(len(c)) < 0.05
        c["x"][holes] = 0.0
        c["y"][holes] = 0.0
        c["z"][holes] = 0.0
    nan_idx = rng.choice(len(f[1]), 2000, replace=False)
    f[1]["z"][nan_idx] = np.nan
    frames = ctx.upload(f, W, H)
    edges, mask = R.edge_extract(ctx, frames, want_mask=True)
    got = edges.download()
    ec = []
    for k in range(2):
        m, _ = orc.canny(f[k], W, H)
        assert np.array_equal(mask[k], m)
        e, _ = orc.extract_edges(f[k], W, H)
        assert np.array_equal(got[k].view(np.uint32), e.view(np.uint32))
        ec.append(e)
    # voxel filter on the finite part (NaN -> int conversion is undefined behaviour in the reference; the GPU and the
    # oracle both use the x86 result, checked here on the hole-only frame)
    v = R.voxel_approx(ctx, ctx.upload([ec[0]])).download()[0]
    assert np.array_equal(v.view(np.uint32), orc.approx_voxel(ec[0]).view(np.uint32))
    vn = R.voxel_approx(ctx, ctx.upload([ec[1]])).download()[0]
    on = orc.approx_voxel(ec[1])
    assert len(vn) == len(on)
    for a in "xyz":  # NaN centroids appear at the same places (their payload bits are not comparable); the rest is bit-equal
        nan_g, nan_o = np.isnan(vn[a]), np.isnan(on[a])
        assert np.array_equal(nan_g, nan_o)
        assert np.array_equal(vn[a][~nan_o].view(np.uint32), on[a][~nan_o].view(np.uint32))
    assert np.array_equal(vn["rgba"], on["rgba"]) and np.isnan(on["z"]).sum() > 0
    guess = np.eye(4)
    guess[:3, :3] = gen_scene.rot_y(-0.523599)
    res, aligned, fc = R.icp_align(ctx, ctx.upload([on]), ctx.upload([orc.approx_voxel(ec[0])]), R.icp_params(), guess=guess,
                                   want_first_corr=True)
    o = orc.icp_align(on, orc.approx_voxel(ec[0]), orc.icp_params(), guess=guess, want_first_corr=True)
    assert np.array_equal(fc, o["first_corr"])
    assert res[0]["n_corr"] == o["n_corr"] and res[0]["converged"] == o["converged"]
    ang, tr = pose_err(res[0]["T"], o["T"])
    assert ang < 1e-4 and tr < 1e-4


# ------------------------------------------------------------------ global-memory ICP path (certified-cache passes)
@pytest.fixture
def global_path(monkeypatch):
    """Force icp_align off the shared-memory persistent kernel (what clouds > 14336 target points take anyway)."""
    monkeypatch.setenv("RSPCL_ICP_PERSIST", "0")
    yield monkeypatch


def test_icp_global_path_matches_oracle_and_legacy(ctx, pair2, global_path):
    fr, _ = pair2
    tgt, src = edge_pair(fr)
    guess = np.eye(4)
    guess[:3, :3] = gen_scene.rot_y(-0.523599)
    gp, op = forced(30)
    o = orc.icp_align(src, tgt, op, guess=guess, want_first_corr=True)
    out = {}
    for cache in ("1", "0"):  # k_icp_stream + k_icp_rescan vs the per-iteration full scan (k_icp_step)
        global_path.setenv("RSPCL_ICP_CACHE", cache)
        res, _, fc = R.icp_align(ctx, ctx.upload([src]), ctx.upload([tgt]), gp, guess=guess, want_aligned=False, want_first_corr=True)
        assert np.array_equal(fc, o["first_corr"])
        r = res[0]
        assert r["iterations"] == o["iterations"] == 30
        ang, tr = pose_err(r["T"], o["T"])
        assert ang < 1e-4 and tr < 1e-4, (cache, ang, tr)
        assert abs(r["n_corr"] - o["n_corr"]) <= max(3, o["n_corr"] // 200)
        out[cache] = r
    # both GPU paths take the exact nearest neighbour every iteration: same correspondences, same sums up to fp64 order
    assert out["1"]["n_corr"] == out["0"]["n_corr"]
    assert np.abs(out["1"]["T"] - out["0"]["T"]).max() < 1e-6


def test_icp_global_path_batch_mixed_states(ctx, global_path):
    rng = np.random.default_rng(29)
    tgt = rand_cloud(rng, 4000, 0.5)
    s_ok = orc.transform(tgt[::2], np.linalg.inv(rigid(rng, 0.002, 0.001)))
    s_far = tgt[::3].copy()
    s_far["x"] += 10.0
    s_nan = s_ok.copy()
    s_nan["y"][::7] = np.nan
    srcs = [s_ok, s_far, s_nan, np.zeros(0, R.POINT)]
    gp, op = forced(12)
    res, aligned, fc = R.icp_align(ctx, ctx.upload(srcs), ctx.upload([tgt]), gp, want_first_corr=True)
    al = aligned.download()
    off = 0
    for k, s in enumerate(srcs):
        o = orc.icp_align(s, tgt, op, want_first_corr=True)
        assert np.array_equal(fc[off:off + len(s)], o["first_corr"]), k
        off += len(s)
        assert res[k]["converged"] == o["converged"] and res[k]["state"] == o["state"] and res[k]["iterations"] == o["iterations"], k
        assert res[k]["n_corr"] == o["n_corr"], k
        if o["converged"]:
            ang, tr = pose_err(res[k]["T"], o["T"])
            assert ang < 1e-4 and tr < 1e-4
        fin = np.isfinite(s["y"])
        assert np.array_equal(al[k][fin].view(np.uint32), orc.transform(s, res[k]["T"])[fin].view(np.uint32))


def test_icp_large_target_takes_global_path(ctx):
    """A target above the shared-memory capacity (14336 points) goes to the grid path without any env override."""
    rng = np.random.default_rng(31)
    tgt = rand_cloud(rng, 40000, 0.5)
    T = rigid(rng, 0.002, 0.002)
    src = orc.transform(tgt[::3], np.linalg.inv(T))
    kw = dict(max_iterations=15, max_corr_dist=0.02, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300,
              mse_threshold_absolute=-1.0)
    res, _, fc = R.icp_align(ctx, ctx.upload([src]), ctx.upload([tgt]), R.icp_params(**kw), want_aligned=False, want_first_corr=True)
    o = orc.icp_align(src, tgt, orc.icp_params(**kw), want_first_corr=True)
    assert np.array_equal(fc, o["first_corr"])
    assert res[0]["n_corr"] == o["n_corr"] and res[0]["iterations"] == 15
    ang, tr = pose_err(res[0]["T"], o["T"])
    assert ang < 1e-4 and tr < 1e-4
    ang, tr = pose_err(res[0]["T"], T)
    assert ang < 1e-3 and tr < 1e-3


def test_icp_batch_mixing_shared_memory_and_global_pairs(ctx):
    """One pair of the batch exceeds the shared-memory capacity: only that pair takes the global-memory path, the others
    keep the persistent kernel's result (a whole-batch fallback used to make an entire sweep 3x slower)."""
    rng = np.random.default_rng(37)
    srcs, tgts = [], []
    for k, nt in enumerate((4000, 20000, 3000)):
        t = rand_cloud(rng, nt, 0.5)
        T = rigid(rng, 0.002, 0.002)
        srcs.append(orc.transform(t[::3], np.linalg.inv(T)))
        tgts.append(t)
    kw = dict(max_iterations=12, max_corr_dist=0.02, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300,
              mse_threshold_absolute=-1.0)
    res, aligned, fc = R.icp_align(ctx, ctx.upload(srcs), ctx.upload(tgts), R.icp_params(**kw), want_first_corr=True)
    al = aligned.download()
    off = 0
    for k in range(3):
        o = orc.icp_align(srcs[k], tgts[k], orc.icp_params(**kw), want_first_corr=True)
        assert np.array_equal(fc[off:off + len(srcs[k])], o["first_corr"]), k
        off += len(srcs[k])
        assert res[k]["iterations"] == 12 and res[k]["n_corr"] == o["n_corr"], k
        ang, tr = pose_err(res[k]["T"], o["T"])
        assert ang < 1e-4 and tr < 1e-4, (k, ang, tr)
        assert np.array_equal(al[k].view(np.uint32), orc.transform(srcs[k], res[k]["T"]).view(np.uint32))


def test_icp_gate_as_large_as_the_scene(ctx):
    """A gate of the size of the cloud puts every target point into a handful of cells (degenerate for the shared-memory
    grid: thousands of candidates per cell); results must still match the oracle and the call must not crawl."""
    rng = np.random.default_rng(41)
    tgt = rand_cloud(rng, 3000, 0.2)
    T = rigid(rng, 0.02, 0.01)
    src = orc.transform(tgt[::2], np.linalg.inv(T))
    kw = dict(max_iterations=6, max_corr_dist=0.3, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300,
              mse_threshold_absolute=-1.0)
    res, _, fc = R.icp_align(ctx, ctx.upload([src]), ctx.upload([tgt]), R.icp_params(**kw), want_aligned=False, want_first_corr=True)
    o = orc.icp_align(src, tgt, orc.icp_params(**kw), want_first_corr=True)
    assert np.array_equal(fc, o["first_corr"])
    assert res[0]["n_corr"] == o["n_corr"] == len(src)
    ang, tr = pose_err(res[0]["T"], o["T"])
    assert ang < 1e-4 and tr < 1e-4


# ------------------------------------------------------------------ sequential schemes, device-resident (row f3)
@pytest.mark.parametrize("scheme", ["icp", "ndt"])
def test_register_sequence_matches_oracle_scheme(ctx, sweep3, scheme):
    """rspcl_register_sequence = the reference's accumulating-target loop (icp:71-124 / ndt:64-112) without leaving HBM:
    transforms against the oracle's line-by-line scheme, merged cloud order (icp:120) and the prepend order of the target
    (icp:119: new points first)."""
    fr, Tgt = sweep3
    frames = ctx.upload(list(fr), W, H)
    rads = np.float32(-0.523599)
    guesses = np.stack([np.eye(4)] * 3)
    acc = np.float32(0)
    for k in (1, 2):
        acc = np.float32(acc + rads)
        guesses[k][:3, :3] = gen_scene.rot_y(float(acc))
    res, merged, target = R.register_sequence(ctx, frames, guesses, R.COARSE_NDT if scheme == "ndt" else R.COARSE_ICP,
                                              want_target=True)
    o = orc.scheme_edge(fr.reshape(-1), W, H, scheme)
    assert [int(r.converged) for r in res] == o["accepted"].tolist() == [1, 1, 1]
    strict = (1,) if scheme == "ndt" else (1, 2)   # chained NDT: see test_gpu_facade.py / DESIGN section 5
    for k in (1, 2):
        T = R.c_to_mat(res[k].T_fine).astype(np.float64) @ R.c_to_mat(res[k].T_coarse).astype(np.float64)
        ang, tr = pose_err(T, o["T"][k])
        if k in strict:
            assert ang < 1e-4 and tr < 1e-4, (scheme, k, ang, tr)
        else:
            assert ang < 0.02 and tr < 0.03, (scheme, k, ang, tr)
        assert res[k].coarse_iterations >= 1 and res[k].fine_iterations == 1
    g = merged.download()[0]
    assert len(g) == len(o["global"]) == 3 * W * H
    assert np.array_equal(g[:W * H].view(np.uint32), fr[0].view(np.uint32))
    assert np.array_equal(g["rgba"], o["global"]["rgba"])
    n_strict = (max(strict) + 1) * W * H
    for a in "xyz":
        assert np.abs(g[a][:n_strict] - o["global"][a][:n_strict]).max() < 5e-4
    # the final edge target: frame 2's fine-aligned edges first, then frame 1's, then frame 0's voxel-filtered edges
    t = target.download()[0]
    v = [orc.approx_voxel(orc.extract_edges(f, W, H)[0]) for f in fr]
    assert len(t) == sum(len(x) for x in v)
    assert np.array_equal(t[len(v[2]) + len(v[1]):].view(np.uint32), v[0].view(np.uint32))
    for k, lo in ((2, 0), (1, len(v[2]))):
        exp = orc.transform(orc.transform(v[k], R.c_to_mat(res[k].T_coarse)), R.c_to_mat(res[k].T_fine))
        assert np.array_equal(t[lo:lo + len(v[k])].view(np.uint32), exp.view(np.uint32)), k


def test_edge_labels_depth_classes_and_rgb(ctx, pair2):
    """Row f2 (partial): the label image rs-pcl --edges shows -- NaN boundary / occluding / occluded from the depth channel
    (OrganizedEdgeBase) plus the RGB-Canny bit -- bit-exact against the oracle, on a frame with NaN holes, (0,0,0) holes
    and a depth step."""
    fr, _ = pair2
    rng = np.random.default_rng(61)
    f = [fr[0].copy(), fr[1].copy()]
    z = f[0]["z"].reshape(H, W)
    z[100:160, 200:300] = np.nan                      # a NaN hole with finite pixels across it
    z[300:340, 50:90] += 1.5                          # an occluded patch (farther than its surroundings)
    z[5:8, :] = np.nan                                # NaNs reaching the image border: nothing to find across them
    holes = rng.random(W * H) < 0.01
    f[1]["z"][holes] = 0.0
    frames = ctx.upload(f, W, H)
    lab = R.edge_labels(ctx, frames)
    for k in range(2):
        exp = orc.depth_edge_labels(f[k], W, H)
        m, _ = orc.canny(f[k], W, H)
        exp = exp | np.where(m > 0, 16, 0).astype(np.uint8)
        assert np.array_equal(lab[k], exp), (k, int((lab[k] != exp).sum()))
    assert (lab[0] & 1).sum() > 0 and (lab[0] & 2).sum() > 0 and (lab[0] & 4).sum() > 0 and (lab[0] & 16).sum() > 3000
