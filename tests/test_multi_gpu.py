"""NCCL merge of partial aggregates on >= 2 GPUs of one box (skipped on a 1-GPU box; the CPU twin of this test is
tests/test_sharding_gloo.py)."""
import os
import subprocess
import sys

import pytest

from tests import common as T

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_merge_over_nccl_equals_whole_table(world):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(T.ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:]
