"""Host-side logic of the pair-sharded multi-GPU sweep (BASELINE configs[3]: a frame sweep whose pairwise registrations
k -> k-1 are independent).  No data-path collective: each rank registers a contiguous block of pairs;
only the 4x4 results are gathered and the trajectory is a prefix product on the host."""
import numpy as np


def shard_pairs(n_frames, world, rank):
    """Pairs are (src=k, tgt=k-1) for k = 1..n_frames-1.  Returns (first_k, last_k_exclusive) of this rank's contiguous
    block; blocks differ in size by at most one pair.  A rank needs frames [first_k - 1, last_k)."""
    n_pairs = max(n_frames - 1, 0)
    base, rem = divmod(n_pairs, world)
    lo = 1 + rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def frames_needed(n_frames, world, rank):
    lo, hi = shard_pairs(n_frames, world, rank)
    if hi <= lo:
        return 0, 0
    return lo - 1, hi


def compose_trajectory(pairwise):
    """pairwise[k] maps frame k into frame k-1 (pairwise[0] ignored).  Returns T[k] mapping frame k into frame 0."""
    n = len(pairwise)
    out = np.zeros((n, 4, 4), np.float64)
    out[0] = np.eye(4)
    for k in range(1, n):
        out[k] = out[k - 1] @ np.asarray(pairwise[k], np.float64)
    return out


def gather_pairwise(local, n_frames, world, rank, dist=None):
    """local: [hi-lo,4,4] transforms of this rank's block.  Returns the full [n_frames,4,4] array on every rank."""
    full = np.zeros((n_frames, 4, 4), np.float64)
    full[0] = np.eye(4)
    if dist is None or world == 1:
        lo, hi = shard_pairs(n_frames, 1, 0)
        full[lo:hi] = local
        return full
    import torch
    sizes = [shard_pairs(n_frames, world, r) for r in range(world)]
    mx = max(h - l for l, h in sizes)
    buf = torch.zeros((mx, 4, 4), dtype=torch.float64)
    lo, hi = sizes[rank]
    if hi > lo:
        buf[:hi - lo] = torch.from_numpy(np.asarray(local, np.float64))
    if dist.get_backend() == "nccl":
        buf = buf.cuda()
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    for r, (l, h) in enumerate(sizes):
        if h > l:
            full[l:h] = parts[r][:h - l].cpu().numpy()
    return full
