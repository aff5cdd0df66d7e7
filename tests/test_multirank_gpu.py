"""Multi-process NCCL parity on real GPUs (skipped with fewer than 2): the frame-sharded denoise -- both endpoint frames
on rank 0, one broadcast per self-attention layer on a side stream, cross-attention endpoints projected locally, CUDA
graphs with the captured collectives -- against the single-GPU run of the same sequence, for the text processors
(outer / inner) and the three IP-Adapter processors, at every world size the box offers (2 / 4 / 8)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worlds():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return [w for w in (2, 4, 8) if w <= n]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_denoise_equals_single_gpu(world):
    if world not in _worlds():
        pytest.skip(f"needs {world} GPUs")
    port = 23000 + (os.getpid() + world) % 4000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "check_sharded.py"), "--frames", str(max(9, world + 1))]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=850, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    lines = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    cases = {(l["case"], l["graphs"]) for l in lines}
    assert len(cases) == 10, cases                      # 5 processor sets x {eager, graphs}
    assert all(l["ok"] for l in lines), lines
