"""The Umeyama rotation of the ICP solve (icp.cu: polar_rotation_fast) restated in numpy: Newton-on-SO(3) correction
passes from the identity or from an fp32 scaled-Newton seed must reproduce U S V^T of the SVD (reflection fix included)
for well-conditioned, near-planar and reflected covariances.  This pins the accuracy claim of DESIGN.md section 4.1/5; the
device code itself is covered by the GPU parity tests (transforms and n_corr against the oracle)."""
import numpy as np


def skew(k):
    return np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])


def newton_seed_fp32(A):
    X = (A / np.sqrt((A * A).sum() / 3)).astype(np.float32)
    scaling = True
    for _ in range(20):
        Xd = X.astype(np.float64)
        det = np.float32(np.linalg.det(Xd))
        if not det > 0:
            return None
        c = (np.linalg.inv(Xd).T * np.linalg.det(Xd)).astype(np.float32)
        idet = np.float32(1) / det
        a, bq = np.float32(0.5), np.float32(0.5) * idet
        if scaling:
            g2 = np.float32(np.sqrt((c * c).sum() * idet * idet / (X * X).sum()))
            if abs(g2 - 1) < 2e-2:
                scaling = False
            a, bq = np.float32(0.5) * np.sqrt(g2), np.float32(0.5) * idet / np.sqrt(g2)
        Xn = (a * X + bq * c).astype(np.float32)
        diff = float(((Xn - X) ** 2).sum())
        X = Xn
        if not scaling and diff <= 1e-10:
            return X.astype(np.float64)
    return None


def polar_fast(A):
    """Mirror of the device control flow: returns (R, passes) or (None, reason)."""
    scale = np.sqrt((A * A).sum() / 3)
    newton_ok = np.linalg.det(A) > 1e-7 * scale ** 3
    Y, seeded = np.eye(3), False
    for p in range(7):
        if p > 0:
            Y = 0.5 * (Y + np.linalg.inv(Y).T)
        M = Y.T @ A
        H = (M + M.T) / 2
        G = np.trace(H) * np.eye(3) - H
        good = G[0, 0] > 0 and G[0, 0] * G[1, 1] - G[0, 1] ** 2 > 0 and np.linalg.det(G) > 0
        if good:
            k = np.linalg.solve(G, np.array([M[2, 1] - M[1, 2], M[0, 2] - M[2, 0], M[1, 0] - M[0, 1]]))
            kk = k @ k
            good = kk < 1e-2
        if not good:
            if seeded or not newton_ok:
                return None, "exact path"
            Y = newton_seed_fp32(A)
            if Y is None:
                return None, "seed failed"
            seeded = True
            continue
        K = skew(k)
        Y = Y @ (np.eye(3) + K + 0.5 * K @ K)
        if kk < 1e-11:
            return Y, p + 1
    return None, "passes exhausted"


def umeyama_rotation(A):
    U, s, Vt = np.linalg.svd(A)
    S = np.diag([1, 1, np.sign(np.linalg.det(U) * np.linalg.det(Vt))])
    return U @ S @ Vt


def rot(rng, ang):
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    K = skew(ax)
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def test_identity_seed_small_increments_incl_planar_and_reflected():
    rng = np.random.default_rng(1)
    worst, n_fast = 0.0, 0
    for _ in range(3000):
        V, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        sv = np.array([1, 10 ** rng.uniform(-1.5, 0), 10 ** rng.uniform(-9, -2) * rng.choice([-1, 1])])
        A = rot(rng, 10 ** rng.uniform(-7, -1.1)) @ (V @ np.diag(sv) @ V.T) * 10 ** rng.uniform(-2, 2)
        R, passes = polar_fast(A)
        assert R is not None, passes
        n_fast += passes <= 2
        worst = max(worst, np.abs(R - umeyama_rotation(A)).max())
    assert worst < 1e-12, worst
    assert n_fast > 2000  # an ICP increment usually needs one or two passes


def test_fp32_seed_for_large_rotations():
    rng = np.random.default_rng(2)
    worst, solved = 0.0, 0
    for _ in range(2000):
        V, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        sv = 10 ** rng.uniform(-2.0, 0.5, 3)
        A = rot(rng, rng.uniform(0.2, 3.0)) @ (V @ np.diag(sv) @ V.T)
        R, passes = polar_fast(A)
        if R is None:  # ill-conditioned + far from identity: the device falls back to the fp64 iteration / SVD
            continue
        solved += 1
        worst = max(worst, np.abs(R - umeyama_rotation(A)).max())
    assert solved > 1900
    assert worst < 1e-12, worst


def test_degenerate_inputs_are_refused_not_mis_solved():
    # a 180-degree turn about an axis makes A symmetric but not positive definite: G is not positive definite and the
    # identity seed must not be accepted
    H = np.diag([3.0, 2.0, 1.0])
    A = np.diag([1.0, -1.0, -1.0]) @ H
    R, why = polar_fast(A)
    if R is not None:
        assert np.abs(R - umeyama_rotation(A)).max() < 1e-12
    # rank one: no unique rotation, nothing to certify
    A1 = np.outer([1.0, 2.0, 3.0], [0.5, -1.0, 2.0])
    R, why = polar_fast(A1)
    assert R is None or np.abs(R - umeyama_rotation(A1)).max() < 1e-9
