"""CPU tests pinning the oracle (oracle/) against known-answer constants and independent
implementations (numpy restatements, scipy cKDTree / ndimage, numpy SVD, finite differences).  The reference ships no
tests or golden vectors, so these are the only pins: parity is otherwise UNPINNED."""
import numpy as np
import pytest

import gen_scene
import orc

F = np.float32


def rand_cloud(rng, n, scale=1.0):
    p = np.zeros(n, orc.POINT)
    p["x"], p["y"], p["z"] = (rng.uniform(-scale, scale, (3, n))).astype(F)
    p["rgba"] = rng.integers(0, 2**24, n).astype(np.uint32) | np.uint32(0xFF000000)
    return p


def rigid(rng, ang=0.2, tr=0.1):
    T = np.eye(4)
    T[:3, :3] = gen_scene.rot_axis(rng.normal(size=3), rng.uniform(-ang, ang))
    T[:3, 3] = rng.uniform(-tr, tr, 3)
    return T


# ------------------------------------------------------------------ known-answer constants (Appendix B)
def test_gaussian_kernel_constants():
    k = orc.gaussian_kernel3()
    assert k[0, 0] == F(0.07511360943317413) and k[0, 1] == F(0.12384141236543655) and k[1, 1] == F(0.20417995750904083)
    assert np.array_equal(k, k.T) and np.array_equal(k, k[::-1, ::-1])


def test_ndt_gauss_constants():
    d1, d2 = orc.ndt_gauss_constants(1.0)
    assert abs(d1 - (-2.21722524404)) < 1e-10 and abs(d2 - 0.433123004704) < 1e-10
    d1, d2 = orc.ndt_gauss_constants(0.05)
    assert abs(d1 / -0.0010222046435 - 1) < 1e-6 and abs(d2 - 0.999597863459) < 1e-7  # resolution_ is a float in PCL


def test_inverse_leaf_is_exact_100():
    assert F(1.0) / F(0.01) == F(100.0)
    p = np.zeros(3, orc.POINT)
    p["x"] = [0.0149999, -0.0000001, 0.02]
    ijk, slot = orc.voxel_keys(p)
    assert ijk[:, 0].tolist() == [1, -1, 2]
    assert slot[0] == (1 * 7171) & 511 and slot[1] == ((-1 * 7171) & 0xFFFFFFFF) & 511


def test_reference_params():
    p = orc.icp_params()
    assert (p.max_iterations, p.max_corr_dist, p.transformation_epsilon, p.euclidean_fitness_epsilon) == (100, 0.01, 1.0, 1000.0)
    q = orc.ndt_params()
    assert (q.max_iterations, q.transformation_epsilon, q.step_size, q.resolution) == (50, 0.01, 0.1, 1.0)


# ------------------------------------------------------------------ Canny vs an independent numpy restatement
def numpy_canny(rgba, w, h, lo=40.0, hi=100.0):
    from scipy import ndimage
    r, g, b = (rgba >> 16) & 255, (rgba >> 8) & 255, rgba & 255
    gray = ((r.astype(np.int64) + g + b) // 3).astype(F).reshape(h, w)
    k = orc.gaussian_kernel3()

    def conv(img, kern):
        pad = np.pad(img, 1, mode="edge")
        acc = np.zeros_like(img)
        for kr in range(3):
            for kc in range(3):
                acc = (acc + (F(kern[kr, kc]) * pad[kr:kr + h, kc:kc + w]).astype(F)).astype(F)
        return acc

    blur = conv(gray, k)
    gx = conv(blur, np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], F))
    gy = conv(blur, np.array([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], F))
    mag = np.sqrt(((gx * gx).astype(F) + (gy * gy).astype(F)).astype(F)).astype(F)
    ang = (np.arctan2(gy.astype(np.float64), gx.astype(np.float64)).astype(F) * F(57.29578)).astype(F)
    d = np.full((h, w), 255, np.uint8)
    d[((ang <= 22.5) & (ang >= -22.5)) | (ang >= 157.5) | (ang <= -157.5)] = 0
    d[((ang > 22.5) & (ang < 67.5)) | ((ang < -112.5) & (ang > -157.5))] = 45
    d[((ang >= 67.5) & (ang <= 112.5)) | ((ang <= -67.5) & (ang >= -112.5))] = 90
    d[((ang > 112.5) & (ang < 157.5)) | ((ang < -22.5) & (ang > -67.5))] = 135
    mx = np.zeros((h, w), F)
    c = mag[1:-1, 1:-1]
    nb = {0: (mag[1:-1, :-2], mag[1:-1, 2:]), 45: (mag[:-2, :-2], mag[2:, 2:]), 90: (mag[:-2, 1:-1], mag[2:, 1:-1]),
          135: (mag[:-2, 2:], mag[2:, :-2])}
    keep = np.zeros_like(c, bool)
    for ang_bin, (a, b2) in nb.items():
        keep |= (d[1:-1, 1:-1] == ang_bin) & (c >= a) & (c >= b2)
    keep &= c >= lo
    mx[1:-1, 1:-1][keep] = c[keep]
    lab, n = ndimage.label(mx > 0, structure=np.ones((3, 3)))
    strong = np.unique(lab[mx >= hi])
    strong = strong[strong > 0]
    mask = np.isin(lab, strong)
    return mask, dict(blur=blur, gx=gx, gy=gy, mag=mag, dir=d, maxima=mx)


@pytest.mark.parametrize("w,h", [(16, 16), (64, 48), (37, 23)])
def test_canny_matches_numpy_restatement_random(w, h):
    rng = np.random.default_rng(w * 1000 + h)
    img = np.zeros((h, w), np.uint32)
    # blocky random image: strong and weak steps
    for _ in range(12):
        r0, c0 = rng.integers(0, h), rng.integers(0, w)
        img[r0:r0 + rng.integers(2, h), c0:c0 + rng.integers(2, w)] = rng.integers(0, 256)
    img = img + rng.integers(0, 6, (h, w)).astype(np.uint32)
    img = np.minimum(img, 255).astype(np.uint32)
    cloud = np.zeros(w * h, orc.POINT)
    cloud["rgba"] = ((img << 16) | (img << 8) | img).ravel()
    mask, dbg = orc.canny(cloud, w, h, debug=True)
    ref, rd = numpy_canny(cloud["rgba"], w, h)
    for k in ("blur", "gx", "gy", "mag", "maxima"):
        assert np.array_equal(dbg[k].reshape(h, w), rd[k]), k
    if dbg["near_bin_edge"] == 0:
        assert np.array_equal(dbg["dir"].reshape(h, w)[rd["mag"] > 0], rd["dir"][rd["mag"] > 0])
    assert np.array_equal(mask > 0, ref)
    assert mask.sum() > 0


def test_canny_vertical_step_known_answer():
    w, h = 16, 12
    img = np.zeros((h, w), np.uint32)
    img[:, 8:] = 200
    cloud = np.zeros(w * h, orc.POINT)
    cloud["rgba"] = ((img << 16) | (img << 8) | img).ravel()
    mask, dbg = orc.canny(cloud, w, h, debug=True)
    m = mask > 0
    # gradient is purely horizontal -> direction bin 0 -> NMS along columns; the two columns adjacent to the step
    # have equal magnitude (>= keeps both); rows 0 and h-1 are never maxima (interior only)
    assert not m[0].any() and not m[-1].any()
    assert np.array_equal(np.where(m[5])[0], [7, 8])
    assert (dbg["dir"].reshape(h, w)[5, 7:9] == 0).all()
    assert dbg["gy"].reshape(h, w)[5, 7] == 0


def test_canny_on_scene_and_extract_order(pair2):
    fr, _ = pair2
    mask, near = orc.canny(fr[0], 640, 480)
    ref, _ = numpy_canny(fr[0]["rgba"], 640, 480)
    assert near == 0
    assert np.array_equal(mask > 0, ref)
    e, idx = orc.extract_edges(fr[0], 640, 480)
    assert np.array_equal(idx, np.flatnonzero(mask.ravel()))
    assert np.array_equal(e, fr[0][idx])
    assert 3000 < len(e) < 40000


def test_crop35():
    w, h = 640, 480
    c = np.zeros(w * h, orc.POINT)
    c["rgba"] = np.arange(w * h, dtype=np.uint32)
    out, ow, oh = orc.crop35(c, w, h)
    assert (ow, oh) == (384, 288)
    exp = np.arange(w * h, dtype=np.uint32).reshape(h, w)[96:384, 128:512].ravel()
    assert np.array_equal(out["rgba"], exp)


# ------------------------------------------------------------------ ApproximateVoxelGrid vs a dict restatement
def py_approx_voxel(pts, leaf=0.01):
    inv = F(1.0) / F(leaf)
    hist = {}
    out = []

    def flush(e):
        cnt = F(e["n"])
        s = e["s"]
        xyz = [F(s[i]) / cnt for i in range(3)]
        rgb = [int(F(s[i]) / cnt) for i in range(3, 6)]
        out.append((xyz[0], xyz[1], xyz[2], (rgb[0] << 16) | (rgb[1] << 8) | rgb[2]))

    for p in pts:
        ijk = tuple(int(np.floor(F(p[a]) * inv)) for a in ("x", "y", "z"))
        slot = (ijk[0] * 7171 + ijk[1] * 3079 + ijk[2] * 4231) & 511
        e = hist.get(slot)
        if e is not None and e["k"] != ijk:
            flush(e)
            e = None
        if e is None:
            e = {"k": ijk, "n": 0, "s": [F(0)] * 6}
            hist[slot] = e
        c = int(p["rgba"])
        vals = [F(p["x"]), F(p["y"]), F(p["z"]), F((c >> 16) & 255), F((c >> 8) & 255), F(c & 255)]
        e["s"] = [F(a + b) for a, b in zip(e["s"], vals)]
        e["n"] += 1
    for slot in sorted(hist):
        flush(hist[slot])
    o = np.zeros(len(out), orc.POINT)
    for i, (x, y, z, c) in enumerate(out):
        o[i] = (x, y, z, c)
    return o


def test_approx_voxel_matches_dict_restatement():
    rng = np.random.default_rng(7)
    p = rand_cloud(rng, 3000, 0.08)  # dense: many points per 1 cm voxel and many slot collisions
    # make runs: sort some stretches so consecutive points share voxels
    p[:1500] = np.sort(p[:1500], order=["z", "y", "x"])
    got = orc.approx_voxel(p)
    exp = py_approx_voxel(p)
    assert len(got) == len(exp) and np.array_equal(got, exp)
    assert len(got) < len(p)
    assert (got["rgba"] >> 24 == 0).all()  # alpha byte is dropped by PCL's int packing


def test_approx_voxel_edge_cases():
    assert len(orc.approx_voxel(np.zeros(0, orc.POINT))) == 0
    one = np.zeros(1, orc.POINT)
    one[0] = (0.5, -0.25, 2.0, 0xFF102030)
    out = orc.approx_voxel(one)
    assert len(out) == 1 and out[0]["x"] == F(0.5) and out[0]["rgba"] == 0x102030
    # a voxel revisited after eviction is emitted twice (order-dependent behaviour)
    a = (0.005, 0.005, 0.005)
    ijk_b = None
    for k in range(1, 2000):  # find a different voxel hashing to the same slot
        if (k * 7171) & 511 == 0:
            ijk_b = k
            break
    b = (ijk_b * 0.01 + 0.005, 0.005, 0.005)
    seq = np.zeros(3, orc.POINT)
    seq[0] = a + (0,)
    seq[1] = b + (0,)
    seq[2] = a + (0,)
    out = orc.approx_voxel(seq)
    assert len(out) == 3
    big = np.zeros(5, orc.POINT)
    big["x"] = 0.123
    assert len(orc.approx_voxel(big, (1.0, 1.0, 1.0))) == 1


# ------------------------------------------------------------------ transform / NN
def test_transform_matches_float32_formula():
    rng = np.random.default_rng(3)
    p = rand_cloud(rng, 1000, 3.0)
    T = rigid(rng).astype(F)
    out = orc.transform(p, T)
    for a, row in zip(("x", "y", "z"), range(3)):
        e = ((T[row, 0] * p["x"] + T[row, 1] * p["y"]).astype(F) + T[row, 2] * p["z"]).astype(F) + T[row, 3]
        assert np.array_equal(out[a], e.astype(F))
    assert np.array_equal(out["rgba"], p["rgba"])


def test_nn_kdtree_equals_brute_and_scipy():
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(5)
    tgt = rand_cloud(rng, 4000)
    q = rand_cloud(rng, 1500)
    tgt[100:120] = tgt[50]  # duplicates: lowest index must win
    q[:20] = tgt[100:120]
    ib, db = orc.nn(tgt, q, "brute")
    ik, dk = orc.nn(tgt, q, "kdtree")
    assert np.array_equal(ib, ik) and np.array_equal(db, dk)
    assert (ik[:20] == 50).all() and (dk[:20] == 0).all()
    xyz = lambda c: np.stack([c["x"], c["y"], c["z"]], 1).astype(np.float64)
    dd, ii = cKDTree(xyz(tgt)).query(xyz(q))
    assert np.allclose(np.sqrt(dk.astype(np.float64)), dd, rtol=1e-5, atol=1e-7)
    same = ii == ik
    assert same.mean() > 0.98  # the rest are float ties / duplicates


def test_nn_grid_lattice_ties_lowest_index():
    g = np.arange(6, dtype=F) * F(0.5)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    tgt = np.zeros(X.size, orc.POINT)
    tgt["x"], tgt["y"], tgt["z"] = X.ravel(), Y.ravel(), Z.ravel()
    q = np.zeros(1, orc.POINT)
    q[0] = (0.25, 0.25, 0.25, 0)  # equidistant from 8 lattice points
    ik, _ = orc.nn(tgt, q, "kdtree")
    ib, _ = orc.nn(tgt, q, "brute")
    assert ik[0] == ib[0] == 0


# ------------------------------------------------------------------ Umeyama / ICP
def kabsch(src, tgt):
    ms, mt = src.mean(0), tgt.mean(0)
    S = (tgt - mt).T @ (src - ms) / len(src)
    U, d, Vt = np.linalg.svd(S)
    D = np.diag([1, 1, np.sign(np.linalg.det(U) * np.linalg.det(Vt))])
    R = U @ D @ Vt
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = mt - R @ ms
    return T


@pytest.mark.parametrize("use_float", [False, True])
def test_umeyama_vs_numpy(use_float):
    rng = np.random.default_rng(11)
    for trial in range(20):
        n = int(rng.integers(3, 200))
        src = rng.uniform(-2, 2, (n, 3)).astype(F)
        if trial % 5 == 0:
            src[:, 2] = 0.3  # planar: exercises the det(U)det(V) sign branch
        T = rigid(rng, 1.0, 0.5)
        tgt = (src.astype(np.float64) @ T[:3, :3].T + T[:3, 3] + rng.normal(0, 1e-3, (n, 3))).astype(F)
        got = orc.umeyama(src, tgt, use_float)
        exp = kabsch(src.astype(np.float64), tgt.astype(np.float64))
        tol = 2e-4 if use_float else 2e-6
        assert np.allclose(got, exp, atol=tol), (trial, np.abs(got - exp).max())
        assert abs(np.linalg.det(got[:3, :3].astype(np.float64)) - 1) < 1e-4


def test_icp_recovers_known_transform_and_reference_epsilon_stops_after_one_iteration():
    rng = np.random.default_rng(13)
    tgt = rand_cloud(rng, 3000, 0.5)
    Tgt = rigid(rng, 0.002, 0.001)
    src = orc.transform(tgt, np.linalg.inv(Tgt))
    # literal reference settings (icp:42-45): converges after the first iteration via the transform criterion
    r = orc.icp_align(src, tgt, orc.icp_params(), want_first_corr=True)
    assert r["converged"] and r["iterations"] == 1 and r["state"] == 2
    ib, db = orc.nn(tgt, src, "brute")
    exp = np.where(db.astype(np.float64) > 0.01 * 0.01, -1, ib)
    assert np.array_equal(r["first_corr"], exp)
    # forced iterations: converges to the true transform
    prm = orc.icp_params(max_iterations=30, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300,
                         mse_threshold_absolute=-1.0)
    r = orc.icp_align(src, tgt, prm)
    assert r["iterations"] == 30 and r["state"] == 1
    assert np.allclose(r["T"], Tgt, atol=2e-5)
    assert np.array_equal(r["aligned"], orc.transform(src, r["T"]))
    assert orc.fitness(r["aligned"], tgt) < 1e-9


def test_icp_no_correspondences():
    rng = np.random.default_rng(17)
    tgt = rand_cloud(rng, 100, 0.5)
    src = tgt.copy()
    src["x"] += 10.0
    g = rigid(rng)
    r = orc.icp_align(src, tgt, orc.icp_params(), guess=g)
    assert not r["converged"] and r["state"] == 5 and r["iterations"] == 0
    assert np.allclose(r["T"], g.astype(F))


def test_fitness_definition():
    rng = np.random.default_rng(19)
    tgt = rand_cloud(rng, 500)
    src = rand_cloud(rng, 200)
    _, d2 = orc.nn(tgt, src, "brute")
    assert np.isclose(orc.fitness(src, tgt), d2.astype(np.float64).mean(), rtol=1e-12)
    thr = float(np.median(d2))
    sel = d2.astype(np.float64) <= thr
    assert np.isclose(orc.fitness(src, tgt, thr), d2[sel].astype(np.float64).mean(), rtol=1e-12)
    assert orc.fitness(src, tgt, -1.0) == np.finfo(np.float64).max


# ------------------------------------------------------------------ NDT
def ndt_target(rng, n=6000):
    # points on three noisy planes inside a 3 m box: well-conditioned voxel Gaussians at 1 m resolution
    p = rand_cloud(rng, n, 1.5)
    third = n // 3
    p["x"][:third] = (1.2 + rng.normal(0, 0.02, third)).astype(F)
    p["y"][third:2 * third] = (-0.9 + rng.normal(0, 0.02, third)).astype(F)
    p["z"][2 * third:] = (1.4 + rng.normal(0, 0.02, n - 2 * third)).astype(F)
    return p


def test_ndt_voxel_statistics_vs_numpy():
    rng = np.random.default_rng(23)
    tgt = ndt_target(rng)
    grid = orc.NdtGrid(tgt, orc.ndt_params())
    vox = grid.voxels()
    assert len(vox) > 10
    xyz = np.stack([tgt["x"], tgt["y"], tgt["z"]], 1)
    cell = np.floor(xyz.astype(F) * F(1.0)).astype(np.int64)
    order = np.lexsort((vox["ijk"][:, 0], vox["ijk"][:, 1], vox["ijk"][:, 2]))
    assert np.array_equal(order, np.arange(len(vox)))  # std::map leaf-index order
    for v in vox:
        sel = (cell == v["ijk"]).all(1)
        P = xyz[sel].astype(np.float64)
        n = len(P)
        assert n >= 6 and (v["npts"] == n or v["npts"] == -1)
        mean = P.mean(0)
        assert np.allclose(v["mean"], mean, atol=1e-12)
        assert np.allclose(v["centroid"], mean, atol=1e-5)
        cov = ((P - mean).T @ (P - mean) / n) * ((n - 1.0) / n)  # PCL 1.9.x normalisation
        w, E = np.linalg.eigh(cov)
        assert np.allclose(v["evals"], w, rtol=1e-6, atol=1e-12)
        if v["npts"] > 0:
            lo = 0.01 * w[2]
            if w[0] < lo:
                w2 = np.maximum(w, [lo, lo, 0])
                w2[1] = max(w[1], lo)
                cov = E @ np.diag(w2) @ E.T
            assert np.allclose(v["cov"].reshape(3, 3), cov, rtol=1e-6, atol=1e-12)
            assert np.allclose(v["icov"].reshape(3, 3) @ cov, np.eye(3), atol=1e-6)


def test_ndt_min_points_and_degenerate_voxels():
    p = np.zeros(5, orc.POINT)
    p["x"] = [0.1, 0.2, 0.3, 0.4, 0.5]
    assert len(orc.NdtGrid(p, orc.ndt_params()).voxels()) == 0  # < 6 points: no Gaussian
    q = np.zeros(8, orc.POINT)
    q["x"] = 0.5
    q["y"] = 0.5
    q["z"] = 0.5
    v = orc.NdtGrid(q, orc.ndt_params()).voxels()
    assert len(v) == 1 and v[0]["npts"] == -1  # zero covariance: largest eigenvalue <= 0 -> rejected


def test_ndt_gradient_and_hessian_vs_finite_differences():
    # two anisotropic blobs in neighbouring 1 m voxels; every source point stays within the 1 m radius of both voxel
    # centroids under the finite-difference perturbations, so the (discontinuous) neighbourhood membership is fixed and
    # the analytic derivatives (Magnusson eq. 6.12/6.13) must match central differences of the score / gradient
    rng = np.random.default_rng(29)
    tgt = np.zeros(1000, orc.POINT)
    for b, cx in enumerate((0.5, 1.5)):
        sl = slice(500 * b, 500 * (b + 1))
        tgt["x"][sl] = (cx + rng.normal(0, 0.10, 500).clip(-0.4, 0.4)).astype(F)
        tgt["y"][sl] = (0.5 + rng.normal(0, 0.05, 500).clip(-0.4, 0.4)).astype(F)
        tgt["z"][sl] = (0.5 + rng.normal(0, 0.08, 500).clip(-0.4, 0.4)).astype(F)
    src = np.zeros(300, orc.POINT)
    src["x"] = (1.0 + rng.uniform(-0.25, 0.25, 300)).astype(F)
    src["y"] = (0.5 + rng.uniform(-0.2, 0.2, 300)).astype(F)
    src["z"] = (0.5 + rng.uniform(-0.2, 0.2, 300)).astype(F)
    grid = orc.NdtGrid(tgt, orc.ndt_params())
    assert len(grid.voxels()) == 2
    p0 = np.array([0.02, -0.015, 0.01, 0.05, -0.08, 0.06])
    s0, g0, H0, npairs = grid.derivatives(src, p0)
    assert npairs == 2 * len(src) and s0 > 0  # score is the positive likelihood sum (-d1 * exp)
    assert np.allclose(H0, H0.T, rtol=1e-9, atol=1e-9)
    h = 1e-3
    g_fd = np.zeros(6)
    H_fd = np.zeros((6, 6))
    for i in range(6):
        e = np.zeros(6)
        e[i] = h
        sp, gp, _, n1 = grid.derivatives(src, p0 + e, False)
        sm, gm, _, n2 = grid.derivatives(src, p0 - e, False)
        assert n1 == n2 == npairs
        g_fd[i] = (sp - sm) / (2 * h)
        H_fd[i] = (gp - gm) / (2 * h)
    assert np.allclose(g0, g_fd, rtol=1e-3, atol=1e-3 * np.abs(g0).max())
    # H[4,4] is excluded: PCL's h_ang_d1_ = (-cy*cz, cy*sz, +sy) carries the sign of the literal upstream source
    # (d^2/dry^2 of +sy is -sy); the oracle restates PCL, not the calculus, so that one entry differs from FD.
    msk = np.ones((6, 6), bool)
    msk[4, 4] = False
    assert np.allclose(H0[msk], H_fd[msk], rtol=2e-3, atol=1e-4 * np.abs(H0).max())
    assert abs(H0[4, 4] - H_fd[4, 4]) > 1.0
    # small-angle shortcut (|angle| < 10e-5 -> cos=1, sin=0) is taken at p = 0
    s1, g1, H1, _ = grid.derivatives(src, np.zeros(6))
    s2, g2, H2, _ = grid.derivatives(src, np.array([0, 0, 0, 2e-4, 2e-4, 2e-4]))
    assert np.allclose(g1, g2, rtol=5e-2, atol=1e-2 * np.abs(g1).max())


def test_pose_matrix_round_trip():
    rng = np.random.default_rng(31)
    for _ in range(50):
        p = np.concatenate([rng.uniform(-1, 1, 3), rng.uniform(-1.2, 1.2, 3)])
        T = orc.pose_to_matrix(p)
        assert np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-6)
        q = orc.matrix_to_pose(T)
        assert np.allclose(orc.pose_to_matrix(q), T, atol=2e-6)
    # the reference's accumulated guesses R_y(k * -0.523599) (ndt:86-89) keep rx = rz = 0
    for k in range(1, 8):
        T = np.eye(4)
        T[:3, :3] = gen_scene.rot_y(k * -0.523599)
        q = orc.matrix_to_pose(T)
        assert np.allclose(orc.pose_to_matrix(q), T, atol=2e-6)


def test_ndt_align_recovers_small_offset():
    rng = np.random.default_rng(37)
    tgt = ndt_target(rng, 9000)
    Tgt = np.eye(4)
    Tgt[:3, :3] = gen_scene.rot_y(0.02)
    Tgt[:3, 3] = [0.05, -0.03, 0.04]
    src = orc.transform(tgt[::2], np.linalg.inv(Tgt))
    r = orc.ndt_align(src, tgt, orc.ndt_params())
    assert r["converged"] and 1 <= r["iterations"] <= 52
    assert np.abs(r["T"][:3, 3] - Tgt[:3, 3]).max() < 0.02
    assert np.allclose(r["T"][:3, :3], Tgt[:3, :3], atol=0.01)
    assert np.array_equal(r["aligned"], orc.transform(src, r["T"]))
    assert r["n_derivative_evals"] >= r["iterations"] + 1


# ------------------------------------------------------------------ scheme drivers on BASELINE config 1 / 2 inputs
def pose_err(A, B):
    D = np.linalg.inv(np.asarray(A, np.float64)) @ np.asarray(B, np.float64)
    # atan2 of the skew part: well conditioned for tiny angles (acos(1 - 1e-8) would already read 1.4e-4 rad)
    sk = np.linalg.norm([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]]) / 2
    return np.arctan2(sk, (np.trace(D[:3, :3]) - 1) / 2), np.linalg.norm(D[:3, 3])


def test_register_pair_config2_ground_truth(pair2):
    fr, T = pair2
    gt = gen_scene.pairwise_gt(T, 1)
    guess = np.eye(4)
    guess[:3, :3] = gen_scene.rot_y(-0.523599)
    prm = orc.icp_params(max_iterations=50, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300,
                         mse_threshold_absolute=-1.0)
    r = orc.register_pair(fr[0], fr[1], 640, 480, "icp", icp=prm, guess=guess, want_full=True)
    assert r["converged"] and r["coarse_icp"].iterations == 50 and r["fine"].iterations == 50
    ang, tr = pose_err(r["T"], gt)
    # point-to-point ICP on 1 cm voxel centroids of pixel-quantised edges: a few mm / sub-mrad is what the method gives
    assert ang < 2e-3 and tr < 5e-3
    full = orc.transform(orc.transform(fr[1], r["T_coarse"]), r["T_fine"])
    assert np.array_equal(r["transformed_full"], full)


def test_scheme_edge_ndt_config1(sweep3):
    fr, T = sweep3
    out = orc.scheme_edge(fr.reshape(-1), 640, 480, "ndt")
    assert out["accepted"].tolist() == [1, 1, 1]
    assert len(out["global"]) == 3 * 640 * 480
    assert np.array_equal(out["global"][:640 * 480], fr[0])
    for k in (1, 2):
        ang, tr = pose_err(out["T"][k], T[k])
        assert ang < 0.03 and tr < 0.05, (k, ang, tr)


def test_scheme_edge_icp_imu_guess_matches_fixed_guess(sweep3):
    fr, T = sweep3
    n = 3
    fixed = orc.scheme_edge(fr.reshape(-1), 640, 480, "icp", rads=-0.523599)
    # IMU thetas whose y component reproduces R_y(k * rads): guess uses AngleAxis(-theta.y, Y) (icp:87)
    th = np.zeros((n, 3), F)
    th[:, 1] = [0.1 + 0.523599 * k for k in range(n)]  # absolute angles; the scheme subtracts thetas[0]
    imu = orc.scheme_edge(fr.reshape(-1), 640, 480, "icp", thetas=th)
    assert np.allclose(imu["thetas"][1:, 1], [0.523599, 2 * 0.523599], atol=1e-6)
    assert imu["accepted"].tolist() == fixed["accepted"].tolist() == [1, 1, 1]
    assert np.allclose(imu["T"], fixed["T"], atol=1e-3)


def test_scheme_incremental_small():
    rng = np.random.default_rng(41)
    base = rand_cloud(rng, 4000, 0.4)
    frames = [base]
    for k in range(2):
        frames.append(orc.transform(base, np.linalg.inv(rigid(rng, 0.001, 0.0008))))
    out = orc.scheme_incremental(np.concatenate(frames), len(base), leaf=(0.01, 0.01, 0.01))
    assert out["accepted"].tolist() == [1, 1, 1] and len(out["target"]) == 3 * len(base)


def test_pcd_round_trip(tmp_path, pair2):
    fr, _ = pair2
    path = str(tmp_path / "synth-0.pcd")
    gen_scene.write_pcd(path, fr[0], 640, 480)
    back, w, h = gen_scene.read_pcd(path)
    assert (w, h) == (640, 480) and np.array_equal(back, fr[0])


def test_depth_edge_labels_hand_cases():
    """OrganizedEdgeBase restatement (orc_depth_edge_labels) on cases small enough to label by hand: a raised patch is
    OCCLUDED (farther) and its rim OCCLUDING, a NaN hole with valid depth across it leaves no label when the depths agree,
    NaNs that run into the image border make their finite neighbours NAN_BOUNDARY."""
    W_, H_ = 12, 9
    p = np.zeros(W_ * H_, orc.POINT)
    p["z"] = 1.0
    z = p["z"].reshape(H_, W_)
    z[5:7, 8:10] = 2.0
    z[3, 5] = np.nan
    lab = orc.depth_edge_labels(p, W_, H_)
    assert (lab[5:7, 8:10] == 4).all()                        # farther than its neighbours: occluded
    assert lab[4, 7] == 2 and lab[7, 10] == 2 and lab[5, 7] == 2  # the rim in front of it: occluding
    assert (lab[2:5, 4:7] == 0).all()                          # across the 1-pixel hole the depth is the same
    assert (lab[0] == 0).all() and (lab[:, 0] == 0).all()      # border pixels are never labelled
    q = p.copy()
    q["z"].reshape(H_, W_)[:, :3] = np.nan                     # a NaN band touching the left border
    lab = orc.depth_edge_labels(q, W_, H_)
    assert (lab[1:-1, 3] == 1).all()                           # nothing finite across it: NaN boundary
    # threshold is relative to the depth (edge_extractor.hpp:19 passes 0.2): a 15 % step is not an edge, a 25 % step is
    s = np.zeros(W_ * H_, orc.POINT)
    s["z"] = 2.0
    s["z"].reshape(H_, W_)[:, 6:] = 2.3
    assert (orc.depth_edge_labels(s, W_, H_) == 0).all()
    s["z"].reshape(H_, W_)[:, 6:] = 2.6
    lab = orc.depth_edge_labels(s, W_, H_)
    assert (lab[1:-1, 5] == 2).all() and (lab[1:-1, 6] == 4).all()
