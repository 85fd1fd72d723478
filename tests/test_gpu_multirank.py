"""N NCCL ranks give the same gathered score table as one rank (needs >= 2 GPUs; one rank per GPU, torchrun)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_scores_equal_single_rank():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs (NCCL does not place two ranks on one device); the CPU twin of this test is "
                    "tests/test_parallel_gloo.py, and tools/multirank_check.py is run under gpurun --gpus 2")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "multirank_check.py"), "64"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    line = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["world"] == world and line["gathered_equals_single_rank"] is True
