"""2-GPU NCCL parity: gradients of the captured multi-rank step == mean of the ranks' single-GPU
gradients; spawned with torchrun when the box has >= 2 GPUs (skipped on the 1-GPU test tier; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl.py -m gpu`, log committed under profiles/)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_captured_nccl_gradients_equal_mean_of_rank_gradients(precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = os.path.join(ROOT, "gpurun_out", f"ddp_nccl_check_{precision}.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_ddp_nccl.py"), "--precision", precision, "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = json.load(open(out))
    assert line["ok"] and line["params_bit_identical_across_ranks"] and line["world"] == 2
