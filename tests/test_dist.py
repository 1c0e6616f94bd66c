"""Multi-rank tests: the host-side plan over gloo on CPU (world_size 2 and 3), the NVLink engine over NCCL on GPUs."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def launch(mode, nproc, timeout=600):
    port = 29500 + (os.getpid() % 500) + nproc
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"), mode]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.parametrize("world", [2, 3])
def test_plan_consistent_across_ranks_gloo(built, world):
    r = launch("plan", world)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for k in range(world):
        assert "rank %d plan ok" % k in r.stdout


@pytest.mark.gpu
def test_two_gpu_engine_matches_single_gpu(built):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    r = launch("gpu", 2, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("dist ok") == 9
