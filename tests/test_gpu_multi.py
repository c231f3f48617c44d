"""Multi-GPU slab decomposition on real devices: needs >= 2 GPUs (skipped otherwise)."""
import os
import subprocess
import sys

import pytest

from lattice_qft_b200 import engine as E

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("ws", [2, 4, 8])
def test_slab_decomposition_matches_single_gpu(ws):
    if E.device_count() < ws:
        pytest.skip(f"needs {ws} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ws}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + ws), os.path.join(ROOT, "tests", "_multi_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("ok (") >= 8, res.stdout
