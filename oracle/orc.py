"""ctypes binding of the CPU oracle (oracle/orc.h).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product (realsense-pointcloud_b200/) never does.  PARITY UNPINNED: see oracle/orc.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborc.so")

POINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgba", "<u4")])


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


class IcpParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("max_corr_dist", C.c_double), ("transformation_epsilon", C.c_double),
                ("euclidean_fitness_epsilon", C.c_double), ("mse_threshold_absolute", C.c_double),
                ("min_correspondences", C.c_int), ("umeyama_float", C.c_int)]


class IcpResult(C.Structure):
    _fields_ = [("T", C.c_float * 16), ("converged", C.c_int), ("state", C.c_int), ("iterations", C.c_int),
                ("n_corr", C.c_int), ("mse", C.c_double), ("prev_mse", C.c_double)]


class NdtParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("transformation_epsilon", C.c_double), ("step_size", C.c_double),
                ("resolution", C.c_float), ("outlier_ratio", C.c_double), ("min_points_per_voxel", C.c_int),
                ("min_covar_eigvalue_mult", C.c_double)]


class NdtResult(C.Structure):
    _fields_ = [("T", C.c_float * 16), ("converged", C.c_int), ("iterations", C.c_int),
                ("trans_probability", C.c_double), ("score", C.c_double), ("p", C.c_double * 6),
                ("n_derivative_evals", C.c_int), ("n_hessian_evals", C.c_int)]


class SchemeStats(C.Structure):
    _fields_ = [("n_frames", C.c_int), ("n_accepted", C.c_int), ("t_edges", C.c_double), ("t_voxel", C.c_double),
                ("t_coarse", C.c_double), ("t_fine", C.c_double), ("t_transform", C.c_double), ("t_total", C.c_double)]


NDT_VOXEL = np.dtype([("ijk", "<i4", 3), ("npts", "<i4"), ("centroid", "<f4", 3), ("_pad", "<i4"), ("mean", "<f8", 3),
                      ("cov", "<f8", 9), ("icov", "<f8", 9), ("evals", "<f8", 3)])

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_fitness.restype = C.c_double
        _lib.orc_ndt_derivatives.restype = C.c_double
        _lib.orc_ndt_grid_build.restype = C.c_void_p
        assert NDT_VOXEL.itemsize == 224
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def pts(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == POINT, a.dtype
    return a


def mat_to_c(T):
    """4x4 (row, col) numpy -> column-major float[16]."""
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).T).ravel().copy()


def c_to_mat(t):
    return np.array(t, dtype=np.float32).reshape(4, 4).T.copy()


def gaussian_kernel3():
    k = np.zeros(9, np.float32)
    lib().orc_gaussian_kernel3(_p(k))
    return k.reshape(3, 3)


def canny(cloud, w, h, t_low=40.0, t_high=100.0, debug=False):
    cloud = pts(cloud)
    n = w * h
    mask = np.zeros(n, np.uint8)
    near = C.c_int(0)
    if debug:
        d = {k: np.zeros(n, np.float32) for k in ("blur", "gx", "gy", "mag", "maxima")}
        d["dir"] = np.zeros(n, np.uint8)
        cnt = lib().orc_canny(_p(cloud), w, h, C.c_float(t_low), C.c_float(t_high), _p(mask), _p(d["blur"]), _p(d["gx"]),
                              _p(d["gy"]), _p(d["mag"]), _p(d["dir"]), _p(d["maxima"]), C.byref(near))
        d["near_bin_edge"] = near.value
        d["count"] = cnt
        return mask.reshape(h, w), d
    lib().orc_canny(_p(cloud), w, h, C.c_float(t_low), C.c_float(t_high), _p(mask), None, None, None, None, None, None,
                    C.byref(near))
    return mask.reshape(h, w), near.value


def extract_edges(cloud, w, h, t_low=40.0, t_high=100.0):
    cloud = pts(cloud)
    out = np.zeros(w * h, POINT)
    idx = np.zeros(w * h, np.int32)
    n = lib().orc_extract_edges(_p(cloud), w, h, C.c_float(t_low), C.c_float(t_high), _p(out), _p(idx))
    return out[:n].copy(), idx[:n].copy()


def depth_edge_labels(cloud, w, h, th_depth_discon=0.2, max_search_neighbors=50):
    """OrganizedEdgeBase labels (1 NaN boundary, 2 occluding, 4 occluded) with edge_extractor.hpp:19-20's settings."""
    cloud = pts(cloud)
    lab = np.zeros(w * h, np.uint8)
    lib().orc_depth_edge_labels(_p(cloud), w, h, C.c_float(th_depth_discon), int(max_search_neighbors), _p(lab))
    return lab.reshape(h, w)


def crop35(cloud, w, h):
    cloud = pts(cloud)
    out = np.zeros((w * 3 // 5) * (h * 3 // 5), POINT)
    ow, oh = C.c_int(), C.c_int()
    lib().orc_crop35(_p(cloud), w, h, _p(out), C.byref(ow), C.byref(oh))
    return out, ow.value, oh.value


def approx_voxel(cloud, leaf=(0.01, 0.01, 0.01)):
    cloud = pts(cloud)
    out = np.zeros(max(len(cloud), 1), POINT)
    lf = np.asarray(leaf, np.float32)
    n = lib().orc_approx_voxel(_p(cloud), len(cloud), _p(lf), _p(out))
    return out[:n].copy()


def voxel_keys(cloud, leaf=(0.01, 0.01, 0.01)):
    cloud = pts(cloud)
    ijk = np.zeros((len(cloud), 3), np.int32)
    slot = np.zeros(len(cloud), np.int32)
    lf = np.asarray(leaf, np.float32)
    lib().orc_voxel_keys(_p(cloud), len(cloud), _p(lf), _p(ijk), _p(slot))
    return ijk, slot


def transform(cloud, T):
    cloud = pts(cloud)
    out = np.zeros(len(cloud), POINT)
    t = mat_to_c(T)
    lib().orc_transform(_p(cloud), len(cloud), _p(t), _p(out))
    return out


def nn(tgt, q, method="kdtree"):
    tgt, q = pts(tgt), pts(q)
    idx = np.zeros(len(q), np.int32)
    d2 = np.zeros(len(q), np.float32)
    f = lib().orc_nn_kdtree if method == "kdtree" else lib().orc_nn_brute
    f(_p(tgt), len(tgt), _p(q), len(q), _p(idx), _p(d2))
    return idx, d2


def icp_params(reference=True, **kw):
    p = IcpParams()
    (lib().orc_icp_reference_params if reference else lib().orc_icp_default_params)(C.byref(p))
    for k, v in kw.items():
        assert hasattr(p, k), k
        setattr(p, k, v)
    return p


def ndt_params(**kw):
    p = NdtParams()
    lib().orc_ndt_reference_params(C.byref(p))
    for k, v in kw.items():
        assert hasattr(p, k), k
        setattr(p, k, v)
    return p


def icp_align(src, tgt, prm, guess=None, prev_mse=None, want_aligned=True, want_first_corr=False):
    src, tgt = pts(src), pts(tgt)
    res = IcpResult()
    res.prev_mse = np.finfo(np.float64).max if prev_mse is None else prev_mse
    g = mat_to_c(guess) if guess is not None else None
    aligned = np.zeros(max(len(src), 1), POINT) if want_aligned else None
    fc = np.zeros(max(len(src), 1), np.int32) if want_first_corr else None
    lib().orc_icp_align(_p(src), len(src), _p(tgt), len(tgt), C.byref(prm), _p(g), C.byref(res), _p(aligned), _p(fc))
    out = {"T": c_to_mat(res.T), "converged": bool(res.converged), "state": res.state, "iterations": res.iterations,
           "n_corr": res.n_corr, "mse": res.mse, "prev_mse": res.prev_mse}
    if want_aligned:
        out["aligned"] = aligned[:len(src)]
    if want_first_corr:
        out["first_corr"] = fc[:len(src)]
    return out


def icp_align_dump(src, tgt, prm, guess=None, n_iters=1):
    """Correspondences (target index or -1) of the first n_iters iterations: result dict, [n_iters, len(src)] int32."""
    src, tgt = pts(src), pts(tgt)
    res = IcpResult()
    res.prev_mse = np.finfo(np.float64).max
    g = mat_to_c(guess) if guess is not None else None
    corr = np.zeros((n_iters, max(len(src), 1)), np.int32)
    lib().orc_icp_align_dump(_p(src), len(src), _p(tgt), len(tgt), C.byref(prm), _p(g), C.byref(res), int(n_iters), _p(corr))
    out = {"T": c_to_mat(res.T), "converged": bool(res.converged), "state": res.state, "iterations": res.iterations,
           "n_corr": res.n_corr, "mse": res.mse, "prev_mse": res.prev_mse}
    return out, corr[:, :len(src)]


def umeyama(src_xyz, tgt_xyz, use_float=False):
    s = np.ascontiguousarray(src_xyz, np.float32)
    t = np.ascontiguousarray(tgt_xyz, np.float32)
    T = np.zeros(16, np.float32)
    lib().orc_umeyama(_p(s), _p(t), len(s), int(use_float), _p(T))
    return c_to_mat(T)


def fitness(src_transformed, tgt, max_range=np.finfo(np.float64).max):
    s, t = pts(src_transformed), pts(tgt)
    return lib().orc_fitness(_p(s), len(s), _p(t), len(t), C.c_double(max_range))


def ndt_gauss_constants(resolution, outlier_ratio=0.55):
    d1, d2 = C.c_double(), C.c_double()
    lib().orc_ndt_gauss_constants(C.c_float(resolution), C.c_double(outlier_ratio), C.byref(d1), C.byref(d2))
    return d1.value, d2.value


class NdtGrid:
    def __init__(self, tgt, prm):
        tgt = pts(tgt)
        self.h = C.c_void_p(lib().orc_ndt_grid_build(_p(tgt), len(tgt), C.byref(prm)))
        self.prm = prm

    def voxels(self):
        n = lib().orc_ndt_grid_size(self.h)
        out = np.zeros(n, NDT_VOXEL)
        if n:
            lib().orc_ndt_grid_get(self.h, _p(out))
        return out

    def derivatives(self, src, p, compute_hessian=True):
        src = pts(src)
        pp = np.asarray(p, np.float64).copy()
        g = np.zeros(6)
        H = np.zeros(36)
        npairs = C.c_longlong(0)
        s = lib().orc_ndt_derivatives(self.h, _p(src), len(src), C.byref(self.prm), _p(pp), _p(g), _p(H),
                                      int(compute_hessian), C.byref(npairs))
        return s, g, H.reshape(6, 6), npairs.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ndt_grid_free(self.h)
            self.h = None


def ndt_align(src, tgt, prm, guess=None):
    src, tgt = pts(src), pts(tgt)
    res = NdtResult()
    g = mat_to_c(guess) if guess is not None else None
    aligned = np.zeros(max(len(src), 1), POINT)
    lib().orc_ndt_align(_p(src), len(src), _p(tgt), len(tgt), C.byref(prm), _p(g), C.byref(res), _p(aligned))
    return {"T": c_to_mat(res.T), "converged": bool(res.converged), "iterations": res.iterations,
            "trans_probability": res.trans_probability, "score": res.score, "p": np.array(res.p),
            "n_derivative_evals": res.n_derivative_evals, "n_hessian_evals": res.n_hessian_evals,
            "aligned": aligned[:len(src)]}


def pose_to_matrix(p):
    T = np.zeros(16, np.float32)
    pp = np.asarray(p, np.float64).copy()
    lib().orc_pose_to_matrix(_p(pp), _p(T))
    return c_to_mat(T)


def matrix_to_pose(T):
    p = np.zeros(6)
    t = mat_to_c(T)
    lib().orc_matrix_to_pose(_p(t), _p(p))
    return p


def register_pair(frame_tgt, frame_src, w, h, coarse="icp", icp=None, ndt=None, leaf=(0.01, 0.01, 0.01), guess=None,
                  want_full=False):
    ft, fs = pts(frame_tgt), pts(frame_src)
    icp = icp or icp_params()
    ndt = ndt or ndt_params()
    lf = np.asarray(leaf, np.float32)
    g = mat_to_c(guess if guess is not None else np.eye(4))
    Tc, Tf = np.zeros(16, np.float32), np.zeros(16, np.float32)
    cr, nr, fr, st = IcpResult(), NdtResult(), IcpResult(), SchemeStats()
    full = np.zeros(w * h, POINT) if want_full else None
    ok = lib().orc_register_pair(_p(ft), _p(fs), w, h, 1 if coarse == "ndt" else 0, C.byref(icp), C.byref(ndt), _p(lf),
                                 _p(g), _p(Tc), _p(Tf), C.byref(cr), C.byref(nr), C.byref(fr), _p(full), C.byref(st))
    out = {"converged": bool(ok), "T_coarse": c_to_mat(Tc), "T_fine": c_to_mat(Tf), "fine": fr, "stats": st,
           "coarse_icp": cr, "coarse_ndt": nr}
    out["T"] = (out["T_fine"].astype(np.float64) @ out["T_coarse"].astype(np.float64)).astype(np.float32)
    if want_full:
        out["transformed_full"] = full
    return out


def scheme_edge(frames, w, h, coarse="ndt", rads=-0.523599, thetas=None, icp=None, ndt=None, leaf=(0.01, 0.01, 0.01)):
    frames = pts(frames)
    n = len(frames) // (w * h)
    icp = icp or icp_params()
    ndt = ndt or ndt_params()
    lf = np.asarray(leaf, np.float32)
    out = np.zeros(n * w * h, POINT)
    T = np.zeros((n, 16), np.float32)
    acc = np.zeros(n, np.int32)
    st = SchemeStats()
    th = np.ascontiguousarray(thetas, np.float32).copy() if thetas is not None else None
    m = lib().orc_scheme_edge(_p(frames), n, w, h, 1 if coarse == "ndt" else 0, int(thetas is not None),
                              C.c_float(rads), _p(th), C.byref(icp), C.byref(ndt), _p(lf), _p(out), _p(T), _p(acc),
                              C.byref(st))
    return {"global": out[:m], "T": np.stack([c_to_mat(t) for t in T]), "accepted": acc, "stats": st, "thetas": th}


def scheme_incremental(frames, npts, icp=None, leaf=(1.0, 1.0, 1.0)):
    frames = pts(frames)
    n = len(frames) // npts
    icp = icp or icp_params()
    lf = np.asarray(leaf, np.float32)
    out = np.zeros(n * npts, POINT)
    T = np.zeros((n, 16), np.float32)
    acc = np.zeros(n, np.int32)
    m = lib().orc_scheme_incremental(_p(frames), n, npts, C.byref(icp), _p(lf), _p(out), _p(T), _p(acc))
    return {"target": out[:m], "T": np.stack([c_to_mat(t) for t in T]), "accepted": acc}
