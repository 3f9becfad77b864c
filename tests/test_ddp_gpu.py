"""GPU, >= 2 devices: N-rank NCCL training step against the oracle run on the concatenated batch with per-shard
BatchNorm (semantics of new_betavaegan.py:99-193 under DataParallel).  Spawns torchrun over
tools/check_dp_oracle.py (eager and CUDA-graph mode, all three loops).  Skipped on a one-GPU box; the round's own
2-GPU run of it is committed under profiles/."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_step_matches_oracle_on_concatenated_batch():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_dp_oracle.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(r.stdout[-4000:])
    assert r.returncode == 0 and "DP ORACLE PARITY OK" in r.stdout, r.stdout[-4000:]
