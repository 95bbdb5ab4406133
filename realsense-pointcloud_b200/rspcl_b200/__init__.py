"""ctypes binding of librspcl_b200.so (include/rspcl.h) used by the parity tests and bench.py.

The product is the C-ABI library (hand-written sm_100a CUDA) plus the C++ host facade in ../host/; this module only
marshals numpy buffers into that ABI.  There is no CPU fallback: if the shared library or a CUDA device is missing
every entry point raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
SO_PATH = os.environ.get("RSPCL_LIB") or os.path.join(_PKG, "librspcl_b200.so")  # RSPCL_LIB: A/B-test another build

POINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgba", "<u4")])           # RSPCL_LAYOUT_PCD16
PCL32 = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("w", "<f4"), ("rgba", "<u4"),
                  ("pad", "<u4", 3)])                                                    # RSPCL_LAYOUT_PCL32
LAYOUT_PCD16, LAYOUT_PCL32 = 0, 1
COARSE_ICP, COARSE_NDT = 0, 1
DBL_MAX = float(np.finfo(np.float64).max)


class RspclError(RuntimeError):
    pass


class IcpParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("min_correspondences", C.c_int32), ("max_corr_dist", C.c_double),
                ("transformation_epsilon", C.c_double), ("euclidean_fitness_epsilon", C.c_double),
                ("mse_threshold_absolute", C.c_double)]


class IcpResult(C.Structure):
    _fields_ = [("T", C.c_float * 16), ("converged", C.c_int32), ("state", C.c_int32), ("iterations", C.c_int32),
                ("n_corr", C.c_int32), ("mse", C.c_double), ("prev_mse", C.c_double)]


class NdtParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("min_points_per_voxel", C.c_int32),
                ("transformation_epsilon", C.c_double), ("step_size", C.c_double), ("outlier_ratio", C.c_double),
                ("min_covar_eigvalue_mult", C.c_double), ("resolution", C.c_float), ("_pad", C.c_int32)]


class NdtResult(C.Structure):
    _fields_ = [("T", C.c_float * 16), ("converged", C.c_int32), ("iterations", C.c_int32),
                ("n_derivative_evals", C.c_int32), ("n_hessian_evals", C.c_int32), ("trans_probability", C.c_double),
                ("score", C.c_double), ("p", C.c_double * 6)]


class PairResult(C.Structure):
    _fields_ = [("T_coarse", C.c_float * 16), ("T_fine", C.c_float * 16), ("converged", C.c_int32),
                ("coarse_iterations", C.c_int32), ("fine_iterations", C.c_int32), ("n_corr", C.c_int32),
                ("n_src", C.c_int32), ("n_tgt", C.c_int32), ("mse", C.c_double)]


NDT_VOXEL = np.dtype([("ijk", "<i4", 3), ("npts", "<i4"), ("centroid", "<f4", 3), ("_pad", "<i4"), ("mean", "<f8", 3),
                      ("cov", "<f8", 9), ("icov", "<f8", 9), ("evals", "<f8", 3)])

EXPORTS = [
    "rspcl_ctx_create", "rspcl_ctx_destroy", "rspcl_last_error", "rspcl_ctx_sync", "rspcl_timer_start",
    "rspcl_timer_stop", "rspcl_timer_mark", "rspcl_timer_span", "rspcl_launch_count", "rspcl_profile_enable", "rspcl_profile_reset", "rspcl_profile_get", "rspcl_host_alloc", "rspcl_host_free", "rspcl_cloud_create",
    "rspcl_cloud_destroy", "rspcl_cloud_n_seg", "rspcl_cloud_stride", "rspcl_cloud_dims", "rspcl_cloud_upload",
    "rspcl_cloud_counts", "rspcl_cloud_download", "rspcl_cloud_invalidate_gray", "rspcl_cloud_download_xyz_pcl32", "rspcl_crop35", "rspcl_edge_extract", "rspcl_edge_labels", "rspcl_voxel_approx",
    "rspcl_voxel_keys", "rspcl_transform", "rspcl_concat", "rspcl_cloud_copy_segment", "rspcl_icp_reference_params",
    "rspcl_icp_align", "rspcl_icp_align_dump", "rspcl_debug_plan_clusters", "rspcl_fitness", "rspcl_nearest", "rspcl_ndt_reference_params", "rspcl_ndt_align",
    "rspcl_ndt_voxels", "rspcl_ndt_derivatives", "rspcl_register_pairs", "rspcl_register_sequence", "rspcl_comm_unique_id", "rspcl_comm_init",
    "rspcl_comm_destroy", "rspcl_icp_align_sharded", "rspcl_ndt_align_sharded",
]

_lib = None


def build(force=False):
    """Compile the CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", _PKG, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", _PKG, "-j", "8"], stdout=subprocess.DEVNULL)
    return SO_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RspclError("librspcl_b200.so is missing: run `make -C realsense-pointcloud_b200` "
                             "(__graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(SO_PATH)
        L.rspcl_last_error.restype = C.c_char_p
        L.rspcl_last_error.argtypes = [C.c_void_p]
        L.rspcl_launch_count.restype = C.c_longlong
        L.rspcl_launch_count.argtypes = [C.c_void_p]
        for name in EXPORTS:
            f = getattr(L, name)
            if name not in ("rspcl_last_error", "rspcl_launch_count", "rspcl_ctx_destroy", "rspcl_cloud_destroy",
                            "rspcl_icp_reference_params", "rspcl_ndt_reference_params"):
                f.restype = C.c_int
        L.rspcl_ctx_destroy.restype = None
        L.rspcl_cloud_destroy.restype = None
        _lib = L
    return _lib


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return a


def mat_to_c(T):
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).T).ravel().copy()


def c_to_mat(t):
    return np.array(t, dtype=np.float32).reshape(4, 4).T.copy()


def mats_to_c(Ts, n):
    Ts = np.asarray(Ts, np.float32)
    if Ts.ndim == 2:
        Ts = np.broadcast_to(Ts, (n, 4, 4))
    return np.ascontiguousarray(np.transpose(Ts, (0, 2, 1))).reshape(n, 16).copy()


def to_pcl32(pts):
    out = np.zeros(len(pts), PCL32)
    out["x"], out["y"], out["z"], out["w"], out["rgba"] = pts["x"], pts["y"], pts["z"], 1.0, pts["rgba"]
    return out


def from_pcl32(p):
    out = np.zeros(len(p), POINT)
    out["x"], out["y"], out["z"], out["rgba"] = p["x"], p["y"], p["z"], p["rgba"]
    return out


def timer_span(ctxs):
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    ms = C.c_float()
    ctxs[0].check(lib().rspcl_timer_span(arr, len(ctxs), C.byref(ms)))
    return ms.value


def icp_params(**kw):
    p = IcpParams()
    lib().rspcl_icp_reference_params(C.byref(p))
    for k, v in kw.items():
        assert hasattr(p, k), k
        setattr(p, k, v)
    return p


def ndt_params(**kw):
    p = NdtParams()
    lib().rspcl_ndt_reference_params(C.byref(p))
    for k, v in kw.items():
        assert hasattr(p, k), k
        setattr(p, k, v)
    return p


class Context:
    def __init__(self, device=0):
        self.h = C.c_void_p()
        rc = lib().rspcl_ctx_create(int(device), C.byref(self.h))
        if rc != 0:
            raise RspclError("rspcl_ctx_create(device=%d) failed with status %d: no CUDA device / no CPU fallback"
                             % (device, rc))
        self.device = device

    def check(self, rc):
        if rc != 0:
            raise RspclError("rspcl status %d: %s" % (rc, lib().rspcl_last_error(self.h).decode()))

    def sync(self):
        self.check(lib().rspcl_ctx_sync(self.h))

    def timer_start(self):
        self.check(lib().rspcl_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self.check(lib().rspcl_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def timer_mark(self):
        self.check(lib().rspcl_timer_mark(self.h))

    def launches(self):
        return lib().rspcl_launch_count(self.h)

    def profile(self, on):
        self.check(lib().rspcl_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self.check(lib().rspcl_profile_reset(self.h))

    def profile_get(self, kernel):
        ms, n, u = C.c_double(), C.c_longlong(), C.c_double()
        self.check(lib().rspcl_profile_get(self.h, kernel.encode(), C.byref(ms), C.byref(n), C.byref(u)))
        return {"ms": ms.value, "launches": n.value, "units": u.value}

    def pinned(self, nbytes):
        """numpy uint8 view of a pinned host allocation (kept alive by the returned array's base)."""
        ptr = C.c_void_p()
        self.check(lib().rspcl_host_alloc(self.h, C.c_size_t(nbytes), C.byref(ptr)))
        buf = (C.c_uint8 * nbytes).from_address(ptr.value)
        arr = np.frombuffer(buf, np.uint8)
        return arr

    def close(self):
        if self.h:
            lib().rspcl_ctx_destroy(self.h)
            self.h = C.c_void_p()

    # ---------------------------------------------------------------- clouds
    def cloud(self, n_seg, stride):
        return Cloud(self, n_seg, stride)

    def upload(self, segments, width=0, height=0, layout=LAYOUT_PCD16, stride=None):
        """segments: list of POINT (or PCL32) arrays -> Cloud batch."""
        counts = np.array([len(s) for s in segments], np.int32)
        st = int(stride if stride is not None else max(1, counts.max() if len(counts) else 1))
        c = Cloud(self, len(segments), st)
        c.upload(segments, width, height, layout)
        return c


class Cloud:
    def __init__(self, ctx, n_seg, stride):
        self.ctx = ctx
        self.h = C.c_void_p()
        ctx.check(lib().rspcl_cloud_create(ctx.h, int(n_seg), int(stride), C.byref(self.h)))
        self.n_seg, self.stride = int(n_seg), int(stride)

    def upload(self, segments, width=0, height=0, layout=LAYOUT_PCD16):
        dt = POINT if layout == LAYOUT_PCD16 else PCL32
        counts = np.array([len(s) for s in segments], np.int32)
        host = np.concatenate([np.ascontiguousarray(s, dt) for s in segments]) if len(segments) else np.zeros(0, dt)
        if len(host) == 0:
            host = np.zeros(1, dt)
        self.ctx.check(lib().rspcl_cloud_upload(self.ctx.h, self.h, _p(host), layout, _p(counts), len(counts),
                                                int(width), int(height)))
        self.ctx.sync()
        return self

    def upload_raw(self, host_ptr, counts, width, height, layout):
        """Asynchronous upload from a (pinned) host buffer address."""
        self.ctx.check(lib().rspcl_cloud_upload(self.ctx.h, self.h, host_ptr, layout, _p(counts), len(counts),
                                                int(width), int(height)))

    def counts(self):
        c = np.zeros(self.n_seg, np.int32)
        self.ctx.check(lib().rspcl_cloud_counts(self.ctx.h, self.h, _p(c)))
        return c

    def download(self, layout=LAYOUT_PCD16):
        cnt = self.counts()
        dt = POINT if layout == LAYOUT_PCD16 else PCL32
        total = int(cnt.sum())
        host = np.zeros(max(total, 1), dt)
        self.ctx.check(lib().rspcl_cloud_download(self.ctx.h, self.h, _p(host), layout, C.c_longlong(max(total, 1)),
                                                  _p(cnt)))
        out, o = [], 0
        for n in cnt:
            out.append(host[o:o + n].copy())
            o += n
        return out

    def invalidate_gray(self):
        self.ctx.check(lib().rspcl_cloud_invalidate_gray(self.ctx.h, self.h))

    def dims(self):
        w, h = C.c_int(), C.c_int()
        lib().rspcl_cloud_dims(self.h, C.byref(w), C.byref(h))
        return w.value, h.value

    def free(self):
        if self.h:
            lib().rspcl_cloud_destroy(self.ctx.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.free()
        except Exception:
            pass


# -------------------------------------------------------------------- operators (mirror the reference call sites)
def crop35(ctx, frames):
    w, h = frames.dims()
    out = Cloud(ctx, frames.n_seg, max(1, (w * 3 // 5) * (h * 3 // 5)))
    ctx.check(lib().rspcl_crop35(ctx.h, frames.h, out.h))
    return out


def edge_extract(ctx, frames, t_low=40.0, t_high=100.0, want_mask=False, out_stride=None):
    w, h = frames.dims()
    out = Cloud(ctx, frames.n_seg, int(out_stride or w * h))
    mask = np.zeros((frames.n_seg, h, w), np.uint8) if want_mask else None
    ctx.check(lib().rspcl_edge_extract(ctx.h, frames.h, C.c_float(t_low), C.c_float(t_high), out.h, _p(mask)))
    return (out, mask) if want_mask else out


def edge_labels(ctx, frames, th_depth_discon=0.2, max_search_neighbors=50, t_low=40.0, t_high=100.0):
    """Per-pixel edge labels (1 NaN boundary, 2 occluding, 4 occluded, 16 RGB Canny): [n_seg, h, w] uint8."""
    w, h = frames.dims()
    lab = np.zeros((frames.n_seg, h, w), np.uint8)
    ctx.check(lib().rspcl_edge_labels(ctx.h, frames.h, C.c_float(th_depth_discon), int(max_search_neighbors),
                                      C.c_float(t_low), C.c_float(t_high), _p(lab)))
    return lab


def voxel_approx(ctx, cloud, leaf=(0.01, 0.01, 0.01), in_place=False):
    lf = np.asarray(leaf, np.float32)
    out = cloud if in_place else Cloud(ctx, cloud.n_seg, cloud.stride)
    ctx.check(lib().rspcl_voxel_approx(ctx.h, cloud.h, _p(lf), out.h))
    return out


def voxel_keys(ctx, cloud, leaf=(0.01, 0.01, 0.01)):
    lf = np.asarray(leaf, np.float32)
    total = int(cloud.counts().sum())
    ijk = np.zeros((max(total, 1), 3), np.int32)
    slot = np.zeros(max(total, 1), np.int32)
    ctx.check(lib().rspcl_voxel_keys(ctx.h, cloud.h, _p(lf), _p(ijk), _p(slot)))
    return ijk[:total], slot[:total]


def transform(ctx, cloud, T, out=None):
    """T: one 4x4 (broadcast) or [n_seg,4,4]."""
    T = np.asarray(T, np.float32)
    broadcast = T.ndim == 2
    t = mat_to_c(T) if broadcast else mats_to_c(T, cloud.n_seg)
    out = out if out is not None else Cloud(ctx, cloud.n_seg, cloud.stride)
    ctx.check(lib().rspcl_transform(ctx.h, cloud.h, _p(t), int(broadcast), out.h))
    return out


def concat(ctx, a, b, stride=None):
    out = Cloud(ctx, a.n_seg, int(stride or (a.stride + b.stride)))
    ctx.check(lib().rspcl_concat(ctx.h, a.h, b.h, out.h))
    return out


def nearest(ctx, query, tgt):
    total = int(query.counts().sum())
    idx = np.zeros(max(total, 1), np.int32)
    d2 = np.zeros(max(total, 1), np.float32)
    ctx.check(lib().rspcl_nearest(ctx.h, query.h, tgt.h, _p(idx), _p(d2)))
    return idx[:total], d2[:total]


def fitness(ctx, src_transformed, tgt, max_range=DBL_MAX):
    out = np.zeros(src_transformed.n_seg, np.float64)
    ctx.check(lib().rspcl_fitness(ctx.h, src_transformed.h, tgt.h, C.c_double(max_range), _p(out)))
    return out


def icp_align(ctx, src, tgt, prm=None, guess=None, prev_mse=None, want_aligned=True, want_first_corr=False):
    prm = prm or icp_params()
    S = src.n_seg
    res = (IcpResult * S)()
    for s in range(S):
        res[s].prev_mse = DBL_MAX if prev_mse is None else float(np.broadcast_to(prev_mse, (S,))[s])
    g = mats_to_c(guess, S) if guess is not None else None
    aligned = Cloud(ctx, S, src.stride) if want_aligned else None
    fc = np.zeros(max(int(src.counts().sum()), 1), np.int32) if want_first_corr else None
    ctx.check(lib().rspcl_icp_align(ctx.h, src.h, tgt.h, C.byref(prm), _p(g), res, aligned.h if aligned else None,
                                    _p(fc)))
    out = [{"T": c_to_mat(r.T), "converged": bool(r.converged), "state": r.state, "iterations": r.iterations,
            "n_corr": r.n_corr, "mse": r.mse, "prev_mse": r.prev_mse} for r in res]
    return out, aligned, fc


def icp_align_dump(ctx, src, tgt, prm=None, guess=None, n_iters=1):
    """Per-iteration correspondences (match index or -1) of the first n_iters iterations: [n_iters, total points]."""
    prm = prm or icp_params()
    S = src.n_seg
    res = (IcpResult * S)()
    for s in range(S):
        res[s].prev_mse = DBL_MAX
    g = mats_to_c(guess, S) if guess is not None else None
    total = max(int(src.counts().sum()), 1)
    corr = np.zeros((n_iters, total), np.int32)
    ctx.check(lib().rspcl_icp_align_dump(ctx.h, src.h, tgt.h, C.byref(prm), _p(g), res, int(n_iters), _p(corr)))
    out = [{"T": c_to_mat(r.T), "converged": bool(r.converged), "state": r.state, "iterations": r.iterations,
            "n_corr": r.n_corr, "mse": r.mse, "prev_mse": r.prev_mse} for r in res]
    return out, corr


def ndt_align(ctx, src, tgt, prm=None, guess=None, want_aligned=True):
    prm = prm or ndt_params()
    S = src.n_seg
    res = (NdtResult * S)()
    g = mats_to_c(guess, S) if guess is not None else None
    aligned = Cloud(ctx, S, src.stride) if want_aligned else None
    ctx.check(lib().rspcl_ndt_align(ctx.h, src.h, tgt.h, C.byref(prm), _p(g), res, aligned.h if aligned else None))
    out = [{"T": c_to_mat(r.T), "converged": bool(r.converged), "iterations": r.iterations,
            "n_derivative_evals": r.n_derivative_evals, "n_hessian_evals": r.n_hessian_evals,
            "trans_probability": r.trans_probability, "score": r.score, "p": np.array(r.p)} for r in res]
    return out, aligned


def ndt_voxels(ctx, tgt, prm=None, capacity=1 << 20):
    prm = prm or ndt_params()
    rec = np.zeros(capacity, NDT_VOXEL)
    nv = np.zeros(tgt.n_seg, np.int32)
    ctx.check(lib().rspcl_ndt_voxels(ctx.h, tgt.h, C.byref(prm), _p(rec), C.c_longlong(capacity), _p(nv)))
    out, o = [], 0
    for n in nv:
        out.append(rec[o:o + n].copy())
        o += n
    return out


def ndt_derivatives(ctx, src, tgt, p, prm=None):
    prm = prm or ndt_params()
    S = src.n_seg
    pp = np.ascontiguousarray(np.broadcast_to(np.asarray(p, np.float64), (S, 6))).copy()
    score = np.zeros(S)
    g = np.zeros((S, 6))
    H = np.zeros((S, 36))
    ctx.check(lib().rspcl_ndt_derivatives(ctx.h, src.h, tgt.h, C.byref(prm), _p(pp), _p(score), _p(g), _p(H)))
    return score, g, H.reshape(S, 6, 6)


def register_pairs(ctx, frames, src_idx, tgt_idx, coarse=COARSE_ICP, icp=None, ndt=None, leaf=(0.01, 0.01, 0.01),
                   t_low=40.0, t_high=100.0, guess=None, out_transformed=None):
    icp = icp or icp_params()
    ndt = ndt or ndt_params()
    si = np.ascontiguousarray(src_idx, np.int32)
    ti = np.ascontiguousarray(tgt_idx, np.int32)
    n = len(si)
    lf = np.asarray(leaf, np.float32)
    g = mats_to_c(guess, n) if guess is not None else None
    res = (PairResult * n)()
    ctx.check(lib().rspcl_register_pairs(ctx.h, frames.h, _p(si), _p(ti), n, int(coarse), C.byref(icp), C.byref(ndt),
                                         _p(lf), C.c_float(t_low), C.c_float(t_high), _p(g), res,
                                         out_transformed.h if out_transformed is not None else None))
    return res


def register_sequence(ctx, frames, guesses, coarse=COARSE_NDT, icp=None, ndt=None, leaf=(0.01, 0.01, 0.01), t_low=40.0,
                      t_high=100.0, out_global=None, want_target=False, target_capacity=None):
    """The reference's accumulating-target loop for one sweep, device-resident (rspcl_register_sequence).
    guesses: [n_frames, 4, 4] (entry 0 unused).  Returns (results, merged cloud, final edge target or None)."""
    icp = icp or icp_params()
    ndt = ndt or ndt_params()
    n = frames.n_seg
    w, h = frames.dims()
    lf = np.asarray(leaf, np.float32)
    g = mats_to_c(guesses, n)
    res = (PairResult * n)()
    out = out_global if out_global is not None else Cloud(ctx, 1, n * w * h)
    tgt = Cloud(ctx, 1, int(target_capacity or n * w * h // 4)) if want_target else None
    ctx.check(lib().rspcl_register_sequence(ctx.h, frames.h, int(coarse), C.byref(icp), C.byref(ndt), _p(lf), C.c_float(t_low),
                                            C.c_float(t_high), _p(g), res, out.h, tgt.h if tgt is not None else None))
    return res, out, tgt


# -------------------------------------------------------------------- point-sharded multi-GPU mode
def comm_unique_id():
    buf = np.zeros(128, np.uint8)
    rc = lib().rspcl_comm_unique_id(_p(buf))
    if rc != 0:
        raise RspclError("rspcl_comm_unique_id failed (libnccl.so.2 not loadable?)")
    return buf


def comm_init(ctx, nranks, rank, unique_id):
    uid = np.ascontiguousarray(unique_id, np.uint8)
    ctx.check(lib().rspcl_comm_init(ctx.h, int(nranks), int(rank), _p(uid)))


def comm_destroy(ctx):
    ctx.check(lib().rspcl_comm_destroy(ctx.h))


def icp_align_sharded(ctx, src_shard, tgt, prm=None, guess=None, want_aligned=False):
    prm = prm or icp_params()
    S = src_shard.n_seg
    res = (IcpResult * S)()
    for s in range(S):
        res[s].prev_mse = DBL_MAX
    g = mats_to_c(guess, S) if guess is not None else None
    aligned = Cloud(ctx, S, src_shard.stride) if want_aligned else None
    ctx.check(lib().rspcl_icp_align_sharded(ctx.h, src_shard.h, tgt.h, C.byref(prm), _p(g), res, aligned.h if aligned else None))
    out = [{"T": c_to_mat(r.T), "converged": bool(r.converged), "state": r.state, "iterations": r.iterations,
            "n_corr": r.n_corr, "mse": r.mse} for r in res]
    return out, aligned


def ndt_align_sharded(ctx, src_shard, tgt, prm=None, guess=None, want_aligned=False):
    prm = prm or ndt_params()
    S = src_shard.n_seg
    res = (NdtResult * S)()
    g = mats_to_c(guess, S) if guess is not None else None
    aligned = Cloud(ctx, S, src_shard.stride) if want_aligned else None
    ctx.check(lib().rspcl_ndt_align_sharded(ctx.h, src_shard.h, tgt.h, C.byref(prm), _p(g), res, aligned.h if aligned else None))
    out = [{"T": c_to_mat(r.T), "converged": bool(r.converged), "iterations": r.iterations, "score": r.score,
            "n_derivative_evals": r.n_derivative_evals, "trans_probability": r.trans_probability} for r in res]
    return out, aligned
