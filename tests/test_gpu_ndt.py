"""GPU parity tests for the NDT path (voxel Gaussians, fused derivatives, Newton + More-Thuente) against the oracle."""
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
    p["rgba"] = rng.integers(0, 2**24, n).astype(np.uint32)
    return p


def ndt_target(rng, n=6000):
    p = rand_cloud(rng, n, 1.5)
    third = n // 3
    p["x"][:third] = (1.2 + rng.normal(0, 0.02, third)).astype(F)
    p["y"][third:2 * third] = (-0.9 + rng.normal(0, 0.02, third)).astype(F)
    p["z"][2 * third:] = (1.4 + rng.normal(0, 0.02, n - 2 * third)).astype(F)
    return p


def pose_err(A, B):
    D = np.linalg.inv(np.asarray(A, np.float64)) @ np.asarray(B, np.float64)
    # atan2 of the skew part: well conditioned for tiny angles (acos(1 - 1e-8) would already read 1.4e-4 rad)
    sk = np.linalg.norm([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]]) / 2
    return np.arctan2(sk, (np.trace(D[:3, :3]) - 1) / 2), np.linalg.norm(D[:3, 3])


def edge_clouds(fr):
    return [orc.approx_voxel(orc.extract_edges(f, W, H)[0]) for f in fr]


@pytest.mark.parametrize("res", [1.0, 0.25])
def test_voxel_gaussians_match_oracle(ctx, res):
    rng = np.random.default_rng(3)
    tg = [ndt_target(rng, 6000), ndt_target(rng, 2500)]
    tg[1]["x"][:40] = np.nan  # non-finite points are skipped
    prm_o, prm_g = orc.ndt_params(resolution=res), R.ndt_params(resolution=res)
    got = R.ndt_voxels(ctx, ctx.upload(tg), prm_g)
    for k in range(2):
        exp = orc.NdtGrid(tg[k], prm_o).voxels()
        g = got[k]
        assert len(g) == len(exp) > 5
        assert np.array_equal(g["ijk"], exp["ijk"]) and np.array_equal(g["npts"], exp["npts"])
        assert np.array_equal(g["centroid"], exp["centroid"])
        assert np.array_equal(g["mean"], exp["mean"])          # sequential fp64 sums in input order: bit-identical
        ok = exp["npts"] > 0
        assert np.allclose(g["cov"], exp["cov"], rtol=1e-12, atol=1e-18)
        assert np.allclose(g["evals"], exp["evals"], rtol=1e-9, atol=1e-15)
        assert np.allclose(g["icov"][ok], exp["icov"][ok], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("n,res", [(5, 1.0), (700, 0.25), (6000, 1.0), (6000, 0.2), (16384, 0.2), (16385, 0.25), (20000, 0.15),
                                   (32768, 0.25), (40000, 0.25)])
def test_single_target_voxel_build_all_size_classes(ctx, n, res):
    """One target per call takes the one-CTA build (k_ndt_build_one): keys / sort in shared memory up to 16,384 points, in
    global memory up to 32,768, the multi-kernel radix-sort path above that.  Same voxel records as the oracle in every class,
    with non-finite points, a far outlier (wide cell box -> more key bits) and duplicates in the input."""
    rng = np.random.default_rng(100 + n)
    t = ndt_target(rng, n)
    if n > 100:
        t["x"][7:19] = np.nan
        t["z"][n // 2] = np.inf
        t["x"][3], t["y"][3], t["z"][3] = F(41.5), F(-77.25), F(13.0)  # outlier far from the room
        t[40:60] = t[20:40]                                            # exact duplicates
    prm_o, prm_g = orc.ndt_params(resolution=res), R.ndt_params(resolution=res)
    g = R.ndt_voxels(ctx, ctx.upload([t]), prm_g)[0]
    exp = orc.NdtGrid(t, prm_o).voxels()
    assert len(g) == len(exp)
    if n > 100:
        assert len(exp) > 3
    assert np.array_equal(g["ijk"], exp["ijk"]) and np.array_equal(g["npts"], exp["npts"])
    assert np.array_equal(g["centroid"], exp["centroid"])
    assert np.array_equal(g["mean"], exp["mean"])
    assert np.array_equal(g["cov"], exp["cov"]) and np.array_equal(g["evals"], exp["evals"])  # same Jacobi, same bits
    ok = exp["npts"] > 0
    assert np.allclose(g["icov"][ok], exp["icov"][ok], rtol=1e-9, atol=1e-9)


def test_single_target_voxel_build_degenerate_inputs(ctx):
    prm_o, prm_g = orc.ndt_params(resolution=0.5), R.ndt_params(resolution=0.5)
    rng = np.random.default_rng(9)
    allnan = ndt_target(rng, 300)
    allnan["y"][:] = np.nan
    assert len(R.ndt_voxels(ctx, ctx.upload([allnan]), prm_g)[0]) == 0 == len(orc.NdtGrid(allnan, prm_o).voxels())
    one_cell = rand_cloud(rng, 900, 0.05)  # every point in one or two cells around the origin
    g, exp = R.ndt_voxels(ctx, ctx.upload([one_cell]), prm_g)[0], orc.NdtGrid(one_cell, prm_o).voxels()
    assert len(g) == len(exp) >= 1 and np.array_equal(g["mean"], exp["mean"]) and np.array_equal(g["npts"], exp["npts"])
    same = np.repeat(rand_cloud(rng, 1, 1.0), 64)  # 64 copies of one point: zero covariance
    g, exp = R.ndt_voxels(ctx, ctx.upload([same]), prm_g)[0], orc.NdtGrid(same, prm_o).voxels()
    assert len(g) == len(exp) == 1 and np.array_equal(g["npts"], exp["npts"]) and np.array_equal(g["mean"], exp["mean"])


def test_voxel_gaussians_on_edge_clouds(ctx, sweep3):
    fr, _ = sweep3
    ec = edge_clouds(fr)
    for res in (1.0, 0.05):
        got = R.ndt_voxels(ctx, ctx.upload(ec), R.ndt_params(resolution=res))
        for k, e in enumerate(ec):
            exp = orc.NdtGrid(e, orc.ndt_params(resolution=res)).voxels()
            assert len(got[k]) == len(exp)
            assert np.array_equal(got[k]["ijk"], exp["ijk"])
            assert np.array_equal(got[k]["mean"], exp["mean"])
            # degenerate (collinear) voxels sit on the eigenvalue sign test: compare the decision where it is well posed
            well = np.abs(exp["evals"][:, 0]) > 1e-9 * np.abs(exp["evals"][:, 2])
            assert np.array_equal(got[k]["npts"][well], exp["npts"][well])


def test_derivatives_match_oracle(ctx):
    rng = np.random.default_rng(5)
    tg = [ndt_target(rng, 6000), ndt_target(rng, 4000)]
    sr = [t[::3].copy() for t in tg]
    poses = np.array([[0.02, -0.015, 0.01, 0.01, -0.02, 0.015], [0, 0, 0, 0, 0, 0]])
    s, g, Hm = R.ndt_derivatives(ctx, ctx.upload(sr), ctx.upload(tg), poses)
    for k in range(2):
        so, go, Ho, npairs = orc.NdtGrid(tg[k], orc.ndt_params()).derivatives(sr[k], poses[k])
        assert npairs > len(sr[k])
        assert abs(s[k] - so) <= 1e-7 * abs(so)
        assert np.allclose(g[k], go, rtol=1e-6, atol=1e-6 * np.abs(go).max())
        assert np.allclose(Hm[k], Ho, rtol=1e-6, atol=1e-6 * np.abs(Ho).max())


def test_ndt_align_matches_oracle_synthetic(ctx):
    rng = np.random.default_rng(7)
    tgs, srs, gts = [], [], []
    for k in range(4):
        t = ndt_target(rng, 9000)
        T = np.eye(4)
        T[:3, :3] = gen_scene.rot_y(0.02 * (k + 1))
        T[:3, 3] = [0.05, -0.03 * k, 0.04]
        tgs.append(t)
        srs.append(orc.transform(t[::2], np.linalg.inv(T)))
        gts.append(T)
    res, aligned = R.ndt_align(ctx, ctx.upload(srs), ctx.upload(tgs))
    al = aligned.download()
    for k in range(4):
        o = orc.ndt_align(srs[k], tgs[k], orc.ndt_params())
        r = res[k]
        assert r["converged"] == o["converged"] and r["iterations"] == o["iterations"], (k, r["iterations"], o["iterations"])
        assert r["n_derivative_evals"] == o["n_derivative_evals"] and r["n_hessian_evals"] == o["n_hessian_evals"]
        ang, tr = pose_err(r["T"], o["T"])
        assert ang < 1e-4 and tr < 1e-4, (k, ang, tr)
        assert abs(r["score"] - o["score"]) <= 1e-5 * abs(o["score"])
        assert abs(r["trans_probability"] - o["trans_probability"]) <= 1e-5 * abs(o["trans_probability"])
        assert np.array_equal(al[k].view(np.uint32), orc.transform(srs[k], r["T"]).view(np.uint32))
        assert np.abs(r["T"][:3, 3] - gts[k][:3, 3]).max() < 0.03


def test_ndt_align_edge_clouds_config1_guess(ctx, sweep3):
    fr, _ = sweep3
    ec = edge_clouds(fr)
    guess = np.eye(4)
    guess[:3, :3] = gen_scene.rot_y(-0.523599)
    res, _ = R.ndt_align(ctx, ctx.upload([ec[1], ec[2]]), ctx.upload([ec[0], ec[1]]), guess=guess, want_aligned=False)
    for k, (s, t) in enumerate(((1, 0), (2, 1))):
        o = orc.ndt_align(ec[s], ec[t], orc.ndt_params(), guess=guess)
        assert res[k]["iterations"] == o["iterations"] and res[k]["converged"] == o["converged"]
        ang, tr = pose_err(res[k]["T"], o["T"])
        assert ang < 1e-4 and tr < 1e-4, (k, ang, tr)


def test_register_pairs_ndt_coarse_matches_oracle(ctx, sweep3):
    fr, _ = sweep3
    frames = ctx.upload(list(fr), W, H)
    guess = np.eye(4)
    guess[:3, :3] = gen_scene.rot_y(-0.523599)
    res = R.register_pairs(ctx, frames, [1, 2], [0, 1], R.COARSE_NDT, guess=guess)
    for i, (s, t) in enumerate(((1, 0), (2, 1))):
        o = orc.register_pair(fr[t], fr[s], W, H, "ndt", guess=guess)
        assert bool(res[i].converged) == o["converged"]
        assert res[i].coarse_iterations == o["coarse_ndt"].iterations
        for key in ("T_coarse", "T_fine"):
            ang, tr = pose_err(R.c_to_mat(getattr(res[i], key)), o[key])
            assert ang < 1e-4 and tr < 1e-4, (i, key, ang, tr)


def test_config3_1280x720_ndt_005_voxels(ctx):
    """BASELINE configs[2]: edge-based NDT on a 1280x720 pair (921,600 points per frame), 0.05 m voxels."""
    W2, H2 = 1280, 720
    fr, Tgt = gen_scene.make_sweep(3, 2, W2, H2, noise_scale=0.2)
    frames = ctx.upload(list(fr), W2, H2)
    edges, mask = R.edge_extract(ctx, frames, want_mask=True)
    ec = []
    for k in range(2):
        m, near = orc.canny(fr[k], W2, H2)
        assert near == 0 and np.array_equal(mask[k], m)
        ec.append(orc.approx_voxel(orc.extract_edges(fr[k], W2, H2)[0]))
    vox = R.voxel_approx(ctx, edges).download()
    for k in range(2):
        assert np.array_equal(vox[k].view(np.uint32), ec[k].view(np.uint32))
    guess = np.eye(4)
    guess[:3, :3] = gen_scene.rot_y(-0.523599)
    kw = dict(resolution=0.05)
    go = orc.NdtGrid(ec[0], orc.ndt_params(**kw)).voxels()
    gg = R.ndt_voxels(ctx, ctx.upload([ec[0]]), R.ndt_params(**kw))[0]
    assert len(gg) == len(go) > 100 and np.array_equal(gg["ijk"], go["ijk"]) and np.array_equal(gg["mean"], go["mean"])
    p0 = orc.matrix_to_pose(guess)
    s, g, Hm = R.ndt_derivatives(ctx, ctx.upload([ec[1]]), ctx.upload([ec[0]]), p0, R.ndt_params(**kw))
    so, gr, Ho, _ = orc.NdtGrid(ec[0], orc.ndt_params(**kw)).derivatives(ec[1], p0)
    assert abs(s[0] - so) <= 1e-6 * max(abs(so), 1.0)
    assert np.allclose(g[0], gr, rtol=1e-5, atol=1e-6 * max(np.abs(gr).max(), 1.0))
    assert np.allclose(Hm[0], Ho, rtol=1e-5, atol=1e-6 * max(np.abs(Ho).max(), 1.0))
    res, _ = R.ndt_align(ctx, ctx.upload([ec[1]]), ctx.upload([ec[0]]), R.ndt_params(**kw), guess=guess, want_aligned=False)
    o = orc.ndt_align(ec[1], ec[0], orc.ndt_params(**kw), guess=guess)
    assert res[0]["converged"] == o["converged"] and res[0]["iterations"] == o["iterations"]
    ang, tr = pose_err(res[0]["T"], o["T"])
    assert ang < 1e-4 and tr < 1e-4, (ang, tr)
