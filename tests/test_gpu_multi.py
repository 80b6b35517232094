"""N-GPU parity (needs >= 2 visible GPUs, otherwise skipped): launches tests/multi_gpu_check.py
under torchrun; see its docstring."""
import os
import subprocess
import sys

import pytest

import helpers

pytestmark = pytest.mark.gpu


def test_z_slab_ranks_match_undecomposed_run():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(helpers.ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTI_GPU_CHECK_OK" in r.stdout
