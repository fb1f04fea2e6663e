"""bench.py's launch contract, the parts that need no GPU: the reference arm runs on rank 0 alone (other ranks exit 0 without
work or output), and without a CUDA device the product arm refuses to run instead of falling back to anything."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT, has_cuda

BENCH = os.path.join(ROOT, "bench.py")


def _run(args, **env):
    e = {**os.environ, **{k: str(v) for k, v in env.items()}}
    return subprocess.run([sys.executable, BENCH, *args], capture_output=True, text=True, env=e, timeout=300)


def test_reference_arm_other_ranks_exit_quietly():
    p = _run(["--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"], RANK=1, LOCAL_RANK=1, WORLD_SIZE=2,
             MASTER_ADDR="127.0.0.1", MASTER_PORT=29999)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.skipif(has_cuda(), reason="only meaningful on a machine without a GPU")
def test_product_arm_fails_loudly_without_gpu():
    p = _run(["--steps", "1", "--warmup", "1", "--nb", "1000"])
    assert p.returncode != 0 and "CUDA device" in (p.stderr + p.stdout)
    assert '"metric"' not in p.stdout  # no bench line from a CPU


def test_help_lists_the_contract_flags():
    p = _run(["--help"])
    assert p.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--shard"):
        assert flag in p.stdout
