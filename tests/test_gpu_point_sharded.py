"""Point-sharded mode on 2 GPUs: the source split over two ranks, 17 fp64 partial sums per iteration exchanged
through NVLink peer memory (or ncclAllReduce), identical solve on both ranks.  Needs two devices: skipped on a 1-GPU box
(run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_point_sharded.py -m gpu`)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("peer", ["1", "0"])
def test_two_rank_icp_equals_single_gpu(peer):
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, RSPCL_PEER_XCHG=peer)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "point_sharded.py"), "--points", "2000000", "--iters", "12",
           "--check-single"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert res["n_gpus"] == 2 and res["iterations"] == 12
    assert res["max_T_spread_over_ranks"] == 0.0            # every rank sums the partials in rank order: identical bits
    assert res["max_abs_T_diff_vs_single_gpu"] < 1e-6
    assert abs(res["n_corr"] - res["n_corr_single_gpu"]) <= 3
    assert res["max_abs_T_error_vs_ground_truth"] < 1e-4
