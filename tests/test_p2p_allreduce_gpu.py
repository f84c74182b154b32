"""The NVLink peer-memory scalar all-reduce (csrc/p2p.cu) needs >= 2 GPUs of one box: runs tools/check_p2p_allreduce.py under
torchrun when they are there (gpurun --gpus 2 ...), skips on a single-GPU box.  Last run: profiles/r02_scaling.md."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_p2p_scalar_allreduce_under_torchrun():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(torch.cuda.device_count(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", "29547",
                        os.path.join(ROOT, "tools", "check_p2p_allreduce.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    j = json.loads(line)
    assert j["world"] == n
    assert j["values_and_bit_identity"] and j["2000_skewed_calls"] and j["autograd_wrapper"] and j["graph_replay"], j
