"""Golden fixtures (tests/golden/rspcl_golden.npz, made by tests/golden/make_golden.py from the oracle on seed 7).
CPU: the oracle still reproduces them bit for bit.  GPU: the CUDA path hits the same vectors without the oracle."""
import os

import numpy as np
import pytest

import gen_scene
import orc

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rspcl_golden.npz"))
W, H = 640, 480
CW, CH, R0, C0 = [int(v) for v in G["crop"]]


def crop(frame):
    return frame.reshape(H, W)[R0:R0 + CH, C0:C0 + CW].reshape(-1).copy()


@pytest.fixture(scope="module")
def frames7():
    return gen_scene.make_sweep(int(G["seed"]), 2)


def guess():
    g = np.eye(4)
    g[:3, :3] = gen_scene.rot_y(-0.523599)
    return g


def test_oracle_reproduces_golden(frames7):
    fr, T = frames7
    small = [crop(fr[0]), crop(fr[1])]
    for k in range(2):
        m, _ = orc.canny(small[k], CW, CH)
        assert np.array_equal(np.packbits(m > 0), G["edge_mask_bits"][k])
    e0, idx0 = orc.extract_edges(small[0], CW, CH)
    assert np.array_equal(idx0, G["edge_idx0"])
    assert np.array_equal(orc.approx_voxel(e0).view(np.uint32).reshape(-1, 4), G["voxel0"])
    ef = [orc.approx_voxel(orc.extract_edges(f, W, H)[0]) for f in fr]
    assert [len(e) for e in ef] == G["n_voxel_full"].tolist()
    r = orc.icp_align(ef[1], ef[0], orc.icp_params(), guess=guess(), want_first_corr=True)
    assert np.array_equal(r["first_corr"], G["icp_ref_first_corr"])
    assert np.array_equal(r["T"], G["icp_ref_T"])
    n = orc.ndt_align(ef[1], ef[0], orc.ndt_params(), guess=guess())
    assert [n["iterations"], n["n_derivative_evals"], n["n_hessian_evals"]] == G["ndt_meta"].tolist()
    assert np.allclose(n["T"], G["ndt_T"], atol=1e-6)


@pytest.mark.gpu
def test_gpu_hits_golden_vectors(frames7):
    import rspcl_b200 as R
    fr, T = frames7
    ctx = R.Context(0)
    small = [crop(fr[0]), crop(fr[1])]
    edges, mask = R.edge_extract(ctx, ctx.upload(small, CW, CH), want_mask=True)
    for k in range(2):
        assert np.array_equal(np.packbits(mask[k] > 0), G["edge_mask_bits"][k])
    e0 = edges.download()[0]
    assert np.array_equal(e0, small[0][G["edge_idx0"]])
    v0 = R.voxel_approx(ctx, ctx.upload([e0])).download()[0]
    assert np.array_equal(v0.view(np.uint32).reshape(-1, 4), G["voxel0"])
    frames = ctx.upload(list(fr), W, H)
    ef = R.voxel_approx(ctx, R.edge_extract(ctx, frames), in_place=True)
    assert ef.counts().tolist() == G["n_voxel_full"].tolist()
    full = ef.download()
    src, tgt = ctx.upload([full[1]]), ctx.upload([full[0]])
    res, _, fc = R.icp_align(ctx, src, tgt, R.icp_params(), guess=guess(), want_first_corr=True, want_aligned=False)
    assert np.array_equal(fc, G["icp_ref_first_corr"])
    assert [res[0]["n_corr"], res[0]["iterations"], res[0]["state"]] == G["icp_ref_ncorr"].tolist()
    assert np.abs(res[0]["T"] - G["icp_ref_T"]).max() < 1e-5
    kw = dict(max_iterations=10, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300, mse_threshold_absolute=-1.0)
    r10, _, _ = R.icp_align(ctx, src, tgt, R.icp_params(**kw), guess=guess(), want_aligned=False)
    assert np.abs(r10[0]["T"] - G["icp_10_T"]).max() < 1e-5
    n, _ = R.ndt_align(ctx, src, tgt, guess=guess(), want_aligned=False)
    assert [n[0]["iterations"], n[0]["n_derivative_evals"], n[0]["n_hessian_evals"]] == G["ndt_meta"].tolist()
    assert np.abs(n[0]["T"] - G["ndt_T"]).max() < 1e-5
    assert abs(n[0]["score"] - float(G["ndt_score"][0])) <= 1e-5 * abs(float(G["ndt_score"][0]))
    vox = R.ndt_voxels(ctx, tgt)[0]
    assert np.array_equal(vox["ijk"], G["ndt_vox_ijk"]) and np.array_equal(vox["npts"], G["ndt_vox_npts"])
    assert np.array_equal(vox["mean"], G["ndt_vox_mean"])
    ctx.close()
