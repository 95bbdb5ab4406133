#!/usr/bin/env python
"""Seeded synthetic D435i-style RGB-D frames of a procedural indoor scene with ground-truth rigid transforms.

Stands in for librealsense capture (out of scope, BASELINE.json north_star).  Frames are organized row-major
clouds of 16-byte points {x,y,z,rgba} in the camera frame (x right, y down, z forward) and can be written as
binary .pcd files named dataset/<prefix>-<k>.pcd, the convention of the reference's --registration mode
(/root/reference/src/main.cpp:79-83).

Ground truth: X_0 = T_k X_k with T_k = [R_y(k*rads) * dR_k | t_k]; rads defaults to the reference's fixed guess
(-0.523599 rad per frame, icp_edge_based_registration.hpp:135,98-100) and (dR_k, t_k) is a small seeded
perturbation the registration has to recover.
"""
import argparse
import os

import numpy as np

POINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgba", "<u4")])

INTRINSICS = {  # D435-like pinhole models
    (640, 480): (385.0, 385.0, 320.0, 240.0),
    (1280, 720): (640.0, 640.0, 640.0, 360.0),
}


def rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float64)


def rot_axis(axis, a):
    axis = np.asarray(axis, np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * (K @ K)


class Scene:
    """Closed 5 x 3 x 4 m room (x width, y height, z depth) with textured walls and a few boxes."""

    def __init__(self, seed):
        rng = np.random.default_rng(seed)
        self.lo = np.array([-2.5, -1.5, -2.0])
        self.hi = np.array([2.5, 1.5, 2.0])
        # six faces: (axis, side). Each face gets a base gray and a list of texture patches in its (u,v) frame.
        self.faces = []
        for axis in range(3):
            for side in range(2):
                base = rng.integers(90, 170, 3)
                patches = []
                n_patch = int(rng.integers(10, 16))
                for _ in range(n_patch):
                    cu, cv = rng.uniform(-2.2, 2.2), rng.uniform(-1.8, 1.8)
                    su, sv = rng.uniform(0.15, 0.7), rng.uniform(0.15, 0.7)
                    kind = int(rng.integers(0, 3))  # 0 solid poster, 1 checker, 2 stripes
                    strong = rng.random() < 0.7
                    delta = rng.integers(110, 160) if strong else rng.integers(45, 95)
                    sign = 1 if rng.random() < 0.5 else -1
                    col = np.clip(base + sign * delta + rng.integers(-15, 15, 3), 0, 255)
                    col2 = np.clip(col - sign * rng.integers(120, 200), 0, 255)
                    pitch = rng.uniform(0.06, 0.18)
                    patches.append((cu, cv, su, sv, kind, col, col2, pitch))
                self.faces.append((axis, side, base, patches))
        # boxes (AABBs) standing in the room, uniform colour per face with a contrasting stripe
        self.boxes = []
        for _ in range(5):
            c = np.array([rng.uniform(-2.0, 2.0), 0.0, rng.uniform(-1.6, 1.6)])
            if np.hypot(c[0], c[2]) < 0.9:
                c[[0, 2]] *= 0.9 / max(np.hypot(c[0], c[2]), 1e-3) * 1.3
            hs = np.array([rng.uniform(0.12, 0.3), rng.uniform(0.2, 0.6), rng.uniform(0.12, 0.3)])
            c[1] = 1.5 - hs[1]  # y down: floor at y=+1.5
            col = rng.integers(20, 235, 3)
            self.boxes.append((c - hs, c + hs, col))

    def _face_color(self, face, u, v):
        axis, side, base, patches = face
        col = np.broadcast_to(base.astype(np.float64), u.shape + (3,)).copy()
        # gentle large-scale shading so walls are not flat (weak gradients, below the Canny low threshold)
        col += (8.0 * np.sin(1.3 * u + 0.7 * v))[..., None]
        for (cu, cv, su, sv, kind, c1, c2, pitch) in patches:
            inside = (np.abs(u - cu) < su) & (np.abs(v - cv) < sv)
            if not inside.any():
                continue
            if kind == 0:
                col[inside] = c1
            elif kind == 1:
                chk = ((np.floor((u - cu) / pitch) + np.floor((v - cv) / pitch)) % 2) == 0
                col[inside & chk] = c1
                col[inside & ~chk] = c2
            else:
                st = (np.floor((u - cu) / pitch) % 2) == 0
                col[inside & st] = c1
                col[inside & ~st] = c2
        return np.clip(col, 0, 255)

    def render(self, R_wc, t_wc, w, h, noise_scale=0.0, rng=None):
        """R_wc, t_wc: camera-to-world pose. Returns organized POINT array of w*h points (camera frame)."""
        fx, fy, cx, cy = INTRINSICS[(w, h)]
        uu, vv = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
        d_cam = np.stack([(uu - cx) / fx, (vv - cy) / fy, np.ones_like(uu)], -1).reshape(-1, 3)
        d = d_cam @ R_wc.T
        o = np.asarray(t_wc, np.float64)
        n = d.shape[0]
        t_best = np.full(n, np.inf)
        color = np.zeros((n, 3))
        with np.errstate(divide="ignore", invalid="ignore"):
            # room faces (seen from inside)
            for face in self.faces:
                axis, side = face[0], face[1]
                plane = self.hi[axis] if side else self.lo[axis]
                t = (plane - o[axis]) / d[:, axis]
                ok = (t > 1e-6) & (t < t_best)
                p = o + t[:, None] * d
                oa = [a for a in range(3) if a != axis]
                ok &= (p[:, oa[0]] >= self.lo[oa[0]] - 1e-9) & (p[:, oa[0]] <= self.hi[oa[0]] + 1e-9)
                ok &= (p[:, oa[1]] >= self.lo[oa[1]] - 1e-9) & (p[:, oa[1]] <= self.hi[oa[1]] + 1e-9)
                if ok.any():
                    col = self._face_color(face, p[ok, oa[0]], p[ok, oa[1]])
                    t_best[ok] = t[ok]
                    color[ok] = col
            # boxes (slab test)
            for (blo, bhi, bcol) in self.boxes:
                t1 = (blo - o) / d
                t2 = (bhi - o) / d
                tn = np.minimum(t1, t2)
                tf = np.maximum(t1, t2)
                tnear = tn.max(1)
                tfar = tf.min(1)
                ok = (tnear < tfar) & (tnear > 1e-6) & (tnear < t_best)
                if ok.any():
                    ax = tn[ok].argmax(1)
                    p = o + tnear[ok, None] * d[ok]
                    shade = np.array([1.0, 0.75, 0.55])[ax]
                    col = bcol[None, :] * shade[:, None]
                    # a contrasting band around the box
                    band = (np.abs(p[:, 1] - (blo[1] + bhi[1]) * 0.5) < 0.04)
                    col[band] = 255 - col[band]
                    t_best[ok] = tnear[ok]
                    color[ok] = col
        z = t_best  # d_cam has z = 1, so camera-frame depth equals the ray parameter
        if noise_scale > 0:
            rng = rng or np.random.default_rng(0)
            sigma = noise_scale * 0.08 * z * z / (fx * 0.05)
            z = z + rng.normal(0.0, 1.0, n) * sigma
        pts = np.zeros(n, POINT)
        pts["x"] = (d_cam[:, 0] * z).astype(np.float32)
        pts["y"] = (d_cam[:, 1] * z).astype(np.float32)
        pts["z"] = z.astype(np.float32)
        c = np.rint(color).astype(np.uint32)
        pts["rgba"] = (np.uint32(255) << 24) | (c[:, 0] << 16) | (c[:, 1] << 8) | c[:, 2]
        return pts


def sample_room_surface(seed, n, chunk=4_000_000):
    """n points uniformly sampled on the inner surface of the 5 x 3 x 4 m room (area-weighted over the six faces),
    as a POINT array -- the dense single-cloud input of the point-sharded configuration (BASELINE configs[4])."""
    rng = np.random.default_rng(seed)
    lo = np.array([-2.5, -1.5, -2.0])
    hi = np.array([2.5, 1.5, 2.0])
    ext = hi - lo
    areas = np.array([ext[1] * ext[2]] * 2 + [ext[0] * ext[2]] * 2 + [ext[0] * ext[1]] * 2)
    out = np.zeros(n, POINT)
    done = 0
    while done < n:
        m = min(chunk, n - done)
        face = rng.choice(6, size=m, p=areas / areas.sum())
        p = lo + rng.random((m, 3)) * ext
        axis = face // 2
        side = face % 2
        p[np.arange(m), axis] = np.where(side == 1, hi[axis], lo[axis])
        out["x"][done:done + m] = p[:, 0]
        out["y"][done:done + m] = p[:, 1]
        out["z"][done:done + m] = p[:, 2]
        out["rgba"][done:done + m] = (np.uint32(255) << 24) | (face.astype(np.uint32) * np.uint32(0x202020) + np.uint32(0x303030))
        done += m
    return out


def make_sweep(seed, n_frames, w=640, h=480, rads=-0.523599, max_rot_deg=0.08, max_trans=0.003, noise_scale=0.0):
    """Returns (frames [n, w*h] POINT, T_gt [n,4,4] float64 with X_0 = T_k X_k)."""
    rng = np.random.default_rng(seed)
    scene = Scene(seed)
    frames = np.zeros((n_frames, w * h), POINT)
    T = np.zeros((n_frames, 4, 4))
    for k in range(n_frames):
        if k == 0:
            dR, t = np.eye(3), np.zeros(3)
        else:
            axis = rng.normal(size=3)
            dR = rot_axis(axis, np.deg2rad(rng.uniform(-max_rot_deg, max_rot_deg)))
            t = rng.uniform(-max_trans, max_trans, 3)
        R = rot_y(k * rads) @ dR
        T[k] = np.eye(4)
        T[k, :3, :3] = R
        T[k, :3, 3] = t
        frames[k] = scene.render(R, t, w, h, noise_scale, rng)
    return frames, T


def pairwise_gt(T, k):
    """Ground-truth transform taking frame k into frame k-1."""
    return np.linalg.inv(T[k - 1]) @ T[k]


# ---------------------------------------------------------------- PCD binary I/O (x y z rgb, 16 B/pt)
def write_pcd(path, pts, w, h):
    pts = np.ascontiguousarray(pts, POINT)
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F F\n"
           "COUNT 1 1 1 1\nWIDTH %d\nHEIGHT %d\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (w, h, w * h))
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(pts.tobytes())


def read_pcd(path):
    with open(path, "rb") as f:
        meta = {}
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("#"):
                continue
            k, _, v = line.partition(" ")
            meta[k] = v
            if k == "DATA":
                break
        w, h, n = int(meta["WIDTH"]), int(meta["HEIGHT"]), int(meta["POINTS"])
        fields = meta["FIELDS"].split()
        if meta["DATA"] == "binary" and fields in (["x", "y", "z", "rgb"], ["x", "y", "z", "rgba"]):
            pts = np.frombuffer(f.read(n * 16), POINT).copy()
        elif meta["DATA"] == "ascii":
            rows = np.loadtxt(f, dtype=np.float64, ndmin=2)
            pts = np.zeros(n, POINT)
            pts["x"], pts["y"], pts["z"] = rows[:, 0], rows[:, 1], rows[:, 2]
            if rows.shape[1] > 3:
                pts["rgba"] = rows[:, 3].astype(np.float32).view(np.uint32)
        else:
            raise ValueError("unsupported PCD layout: %s %s" % (meta["DATA"], fields))
    return pts, w, h


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--prefix", default="synth")
    ap.add_argument("--out", default="dataset")
    ap.add_argument("--noise", type=float, default=0.0)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    fr, T = make_sweep(a.seed, a.frames, a.width, a.height, noise_scale=a.noise)
    for k in range(a.frames):
        write_pcd(os.path.join(a.out, "%s-%d.pcd" % (a.prefix, k)), fr[k], a.width, a.height)
    np.save(os.path.join(a.out, "%s-gt.npy" % a.prefix), T)
    print("wrote %d frames to %s" % (a.frames, a.out))
