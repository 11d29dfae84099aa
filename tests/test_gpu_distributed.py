"""N > 1 on real GPUs: the collated-batch path over NCCL (tactile_gym_b200/distributed.py:CollatedBatch), world size 2 under
torchrun.  Needs two visible GPUs (`gpurun --gpus 2`); skipped on a one-GPU box.  The host-side logic of the same code is covered
on the CPU with gloo (tests/test_distributed_gloo.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("task", ["edge", "push"])
def test_collated_batch_over_nccl_matches_single_gpu(task):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "_dist_gather_child.py"), task]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.count("GATHER-OK") == 2, (out.stdout[-2000:], out.stderr[-4000:])
