"""The torch.distributed (one process per GPU, NCCL) form of the base-sharded kNN on real GPUs: launched exactly as bench.py
is (python -m torch.distributed.run), compared inside the workers with the single-GPU kNN over the whole base
(compute_groundtruth.cpp:396-444: per-part results merged by distance).  Needs >= 2 GPUs; the gloo tests cover the host logic."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_sharded_knn_over_nccl_equals_single_gpu(world, metric):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29610 + world + (0 if metric == "ip" else 10)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "_nccl_knn_worker.py"), metric]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "share_tau=True resync=True full=1 slice=1 untouched=1" in r.stdout
    assert "share_tau=True resync=False full=1 slice=1 untouched=1" in r.stdout
    assert "share_tau=False resync=False full=1 slice=1 untouched=1" in r.stdout
