"""Multi-GPU parity (sharded state, NCCL qubit swaps) -- runs when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_sharded_vs_oracle(nproc):
    import torch

    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    only = os.environ.get("QCB_DIST_NPROC")
    if only and int(only) != nproc:
        pytest.skip(f"QCB_DIST_NPROC={only}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + nproc), os.path.join(ROOT, "tests", "dist_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(res.stdout[-4000:])
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "FAIL" not in res.stdout and res.stdout.count("PASS") >= 15
