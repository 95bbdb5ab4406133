"""World-size-2 gloo test (CPU) of the pair-sharded sweep logic: block partition, result gather, trajectory composition.
The per-pair registration itself is done by the oracle here (tests may use it); on the GPU box the same host logic
drives rspcl_register_pairs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rspcl_b200.sweep as sweep


def test_shard_pairs_partition():
    for n_frames in (1, 2, 3, 10, 64, 65):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = sweep.shard_pairs(n_frames, world, r)
                seen += list(range(lo, hi))
                flo, fhi = sweep.frames_needed(n_frames, world, r)
                if hi > lo:
                    assert flo == lo - 1 and fhi == hi
            assert seen == list(range(1, n_frames))
            sizes = [sweep.shard_pairs(n_frames, world, r) for r in range(world)]
            assert max(h - l for l, h in sizes) - min(h - l for l, h in sizes) <= 1


def test_compose_trajectory():
    rng = np.random.default_rng(0)
    import gen_scene
    T = [np.eye(4)]
    for _ in range(5):
        M = np.eye(4)
        M[:3, :3] = gen_scene.rot_axis(rng.normal(size=3), rng.uniform(-0.5, 0.5))
        M[:3, 3] = rng.normal(size=3)
        T.append(M)
    traj = sweep.compose_trajectory(T)
    assert np.allclose(traj[3], T[1] @ T[2] @ T[3])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for sub in ("oracle", "tools", "realsense-pointcloud_b200"):
        sys.path.insert(0, os.path.join(root, sub))
    import gen_scene
    import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    base = np.zeros(1500, orc.POINT)
    base["x"], base["y"], base["z"] = rng.uniform(-0.4, 0.4, (3, 1500)).astype(np.float32)
    gt = [np.eye(4)]
    frames = [base]
    for k in range(1, n_frames):  # frame k = frame k-1 moved by a small rigid motion
        M = np.eye(4)
        M[:3, :3] = gen_scene.rot_axis(rng.normal(size=3), 0.002)
        M[:3, 3] = rng.uniform(-0.001, 0.001, 3)
        gt.append(M)
        frames.append(orc.transform(frames[-1], np.linalg.inv(M)))
    lo, hi = sweep.shard_pairs(n_frames, world, rank)
    prm = orc.icp_params(max_iterations=20, transformation_epsilon=-1.0, euclidean_fitness_epsilon=-1e300,
                         mse_threshold_absolute=-1.0)
    local = np.stack([orc.icp_align(frames[k], frames[k - 1], prm, want_aligned=False)["T"] for k in range(lo, hi)]) \
        if hi > lo else np.zeros((0, 4, 4))
    full = sweep.gather_pairwise(local, n_frames, world, rank, dist)
    traj = sweep.compose_trajectory(full)
    if rank == 0:
        np.save(os.path.join(out_dir, "traj.npy"), traj)
        np.save(os.path.join(out_dir, "gt.npy"), sweep.compose_trajectory(gt))
    dist.barrier()
    dist.destroy_process_group()


def test_pair_sharded_sweep_world2_gloo(tmp_path):
    n_frames = 6
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_frames, str(tmp_path)), nprocs=2, join=True)
    traj = np.load(tmp_path / "traj.npy")
    gt = np.load(tmp_path / "gt.npy")
    assert traj.shape == (n_frames, 4, 4)
    assert np.abs(traj - gt).max() < 5e-4
