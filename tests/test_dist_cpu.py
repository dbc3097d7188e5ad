"""N>1 path on the CPU: world_size-2 (and 4) gloo processes, one shard each, real all-to-all qubit swaps: gate application
(forward circuit == oracle, canonical layout restored) and adjoint gradients (pair swaps, all-reduced environments == the
oracle's restatement of the reference's shift sweep)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc needed to build the CPU replay library")


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_sharded_protocol(world):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emu_helper
    import oracle

    emu_helper.build()
    oracle.build()
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29610 + world), os.path.join(ROOT, "tests", "gloo_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "PASS" in res.stdout and "gradients rc=0" in res.stdout
